"""Quick device-time probe of the fused step at several env counts (development tool, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from allsteps_isaaclab_b200.config import AllstepsCfg
from allsteps_isaaclab_b200 import synthetic as syn
from allsteps_isaaclab_b200.mdp import AllstepsMDP, PhysicsViews, StepBuffers

B_ALG = 652


def probe(N, steps=50, sets=4, warmup=10):
    from allsteps_isaaclab_b200.workload import ChainedWorkload

    cfg = AllstepsCfg()
    dev = torch.device("cuda:0")
    origins = syn.env_origins_grid(N, cfg.env_spacing).to(dev)
    c5 = "--c5" in sys.argv  # BASELINE config 5: ~half of the envs reset per step, stones regenerated
    grid = "--grid" in sys.argv  # BASELINE config 3: 11 x 11 pitch x yaw grid curriculum (extension)
    mdp = AllstepsMDP(N, device=dev, seed=1, skip_pass2=("--skip-pass2" in sys.argv), intended_regen=c5,
                      grid_bins=11 if grid else 0)
    mdp.generate_stones(origins)
    st0 = syn.random_mdp_state(cfg, N, torch.Generator().manual_seed(1))
    mdp.import_state({k: st0[k] for k in ("curr_target_index", "swing_leg", "target_reach_count",
                                          "episode_length_buf", "potentials")})
    isaac = "--isaac-views" in sys.argv  # slices of root_state_w (N,13) and body_state_w (N,17,13), as Isaac Lab hands out
    out = StepBuffers(N, dev, reset_rows=("--no-rows" not in sys.argv))
    # every input set is generated from the MDP state it meets; the state is rewound when the cycle restarts
    wl = ChainedWorkload(mdp, origins, out, sets, 1234, cfg, layout="isaac" if isaac else "dense",
                         fall_fraction=0.3 if c5 else 0.02, stones_change=c5 or grid)
    pool = wl.sets
    l0 = mdp.launch_count
    if "--three-call" in sys.argv:
        ep_len = torch.zeros(N, dtype=torch.int64, device=dev)

        def one():
            k = wl.j % wl.period
            if k == 0 and wl.j > 0:
                wl.rewind()
            v, d = wl.sets[k]
            ep_len.add_(1)
            mdp.pass1(v, d["actions"], out, episode_length=ep_len)
            if "--device-list" in sys.argv:
                mdp.reset(origins, None, out, episode_length=ep_len)
                mdp.pass2(v, out)
                wl.j += 1
                return
            ids = out.dones.nonzero(as_tuple=False).squeeze(-1)
            if len(ids):
                mdp.reset(origins, ids, out, episode_length=ep_len)
                mdp.pass2(v, out)
            else:
                mdp.no_reset()
            wl.j += 1
    else:
        one = wl.step
    for i in range(warmup):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if "--timing" in sys.argv:
        import ctypes
        buf = (ctypes.c_uint64 * 16)()
        mdp.lib.as_debug_timing(mdp.handle, buf, 1, mdp._stream())
        for i in range(warmup):
            one()
        mdp.lib.as_debug_timing(mdp.handle, buf, 0, mdp._stream())
        n = max(buf[15], 1)
        names = ["M: start->state/window regs", "M: wait root tiles", "M: pass1+reset+pass2+stores", "M: wait at CTA barrier",
                 "M: post-barrier work", "M: fence+barrier+store issue", "M: wait bulk store read", "M: whole CTA",
                 "J: start->joint tiles", "J: joint loop", "J: wait at CTA barrier", "J: post-barrier (obs cols, reset rows)",
                 "J: fence+barrier"]
        for i, nm in enumerate(names):
            print(f"   {nm:42s} {buf[i] / n / 1965.0:7.2f} us")
    stats = mdp.read_stats()
    print(f"N={N:>8}  {ms*1e3:9.1f} us/step  {N/ms/1e6:9.3f} G env-steps/s  {N*B_ALG/ms/1e6:8.1f} GB/s alg  "
          f"reset {100*wl.reset_rate:.1f}%  advance {100*wl.advance_rate:.1f}%  "
          f"launches/step={(mdp.launch_count-l0)/(steps+warmup):.1f}", flush=True)


def set_l2_fetch_granularity(nbytes):
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    torch.cuda.init(); torch.zeros(1, device="cuda")
    rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(nbytes))  # cudaLimitMaxL2FetchGranularity
    val = ctypes.c_size_t(0)
    rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
    print("cudaLimitMaxL2FetchGranularity ->", val.value, "rc", rc, flush=True)


if __name__ == "__main__":
    for a in sys.argv[1:]:
        if a.startswith("--l2gran="):
            set_l2_fetch_granularity(int(a.split("=")[1]))
    ns = [int(a) for a in sys.argv[1:] if a.isdigit()] or [4096, 65536, 262144, 1048576]
    steps = 12 if "--short" in sys.argv else 50
    for n in ns:
        sets = [int(a.split('=')[1]) for a in sys.argv if a.startswith('--sets=')]
        probe(n, steps=steps, warmup=4 if "--short" in sys.argv else 10, sets=sets[0] if sets else 4)
