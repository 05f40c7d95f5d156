"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Allsteps-v0 batched MDP step.

Nothing under ``oracle/`` is product code. Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the checker
(or as the timed CPU baseline). The product package ``allsteps_isaaclab_b200`` never imports it.

Modules
-------
``allsteps_oracle``   CPU torch restatement ("port") of the reference algorithm; travels to the GPU box.
``philox``            numpy twin of the kernels' Philox4x32-10 stream (so uniforms can be injected).
``ref_loader``        imports the UNMODIFIED reference from ``/root/reference`` behind stub modules
                      (only works in the build container; used to pin the port and to make
                      ``tests/golden/*.npz``).
``ref_fake_env``      fake ``self`` hosting the unmodified reference methods on synthetic tensors.

Parity pin: the port is checked (tests/test_oracle_vs_reference.py, when /root/reference exists) and
the committed golden vectors were produced by executing the reference's own methods
(tests/golden/make_golden.py).  The reference's own test-suite holds no fixture for this path
(SURVEY.md section 4), so "reference executed here" is the pin.
"""
