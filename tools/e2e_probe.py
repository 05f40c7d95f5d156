"""Where the end-to-end (host buffers) step time goes: PCIe legs timed alone and together (development tool).

  A  fused step, everything resident in HBM
  B  fused step, contact matrices in pinned host memory (k_contact_gather reads them through PCIe), nothing else
     on the link
  C  bulk H2D of the other inputs alone
  D  B and C concurrently on two streams
  E  D2H of the step's results alone
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from allsteps_isaaclab_b200 import synthetic as syn
from allsteps_isaaclab_b200.config import AllstepsCfg
from allsteps_isaaclab_b200.mdp import AllstepsMDP, PhysicsViews, StepBuffers


def timed(fn, reps, streams=()):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record()
    for s in streams:
        s.wait_stream(main)
    for i in range(reps):
        fn(i)
    for s in streams:
        main.wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    reps = 8
    cfg = AllstepsCfg()
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(1234)
    origins = syn.env_origins_grid(N, cfg.env_spacing).to(dev)
    mdp = AllstepsMDP(N, device=dev, seed=1)
    mdp.generate_stones(origins)
    st = mdp.export_state()
    d = syn.random_physics_state(cfg, st["steps_pos"], st["curr_target_index"], st["swing_leg"], gen)
    out = StepBuffers(N, dev)
    v_dev = PhysicsViews.from_dict(d, origins)
    host = {k: t.cpu().pin_memory() for k, t in d.items() if torch.is_tensor(t)}
    v_zc = PhysicsViews.from_dict({**d, "force_matrix_right": host["force_matrix_right"],
                                   "force_matrix_left": host["force_matrix_left"]}, origins)
    bulk_keys = ["root_pos_w", "root_quat_w", "root_lin_vel_w", "body_pos_w", "joint_pos", "joint_vel", "actions"]
    dst = {k: torch.empty_like(d[k]) for k in bulk_keys}
    bulk_bytes = sum(host[k].numel() * 4 for k in bulk_keys)
    host_out = {"obs": torch.empty(N, 59).pin_memory(), "reward": torch.empty(N).pin_memory(),
                "terminated": torch.empty(N, dtype=torch.bool).pin_memory(),
                "time_out": torch.empty(N, dtype=torch.bool).pin_memory()}
    out_bytes = sum(t.numel() * t.element_size() for t in host_out.values())
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def step_dev(i):
        mdp.step(v_dev, d["actions"], out)

    def step_zc(i):
        mdp.step(v_zc, d["actions"], out)

    def bulk(i):
        with torch.cuda.stream(s_in):
            for k in bulk_keys:
                dst[k].copy_(host[k], non_blocking=True)

    def d2h(i):
        with torch.cuda.stream(s_out):
            for k, t in host_out.items():
                t.copy_(getattr(out, k), non_blocking=True)

    def both(i):
        bulk(i)
        step_zc(i)

    def all3(i):
        bulk(i)
        d2h(i)
        step_zc(i)

    for f in (step_dev, step_zc, bulk, d2h):
        f(0)
    a = timed(step_dev, reps)
    b = timed(step_zc, reps)
    c = timed(bulk, reps, (s_in,))
    dd = timed(both, reps, (s_in,))
    e = timed(d2h, reps, (s_out,))
    f = timed(all3, reps, (s_in, s_out))
    print(f"N={N}")
    print(f"A step, HBM inputs                     {a:8.3f} ms")
    print(f"B step, contact matrices over PCIe     {b:8.3f} ms")
    print(f"C bulk H2D {bulk_bytes/1e6:7.1f} MB                {c:8.3f} ms  {bulk_bytes/c/1e6:6.1f} GB/s")
    print(f"D B + C concurrently                   {dd:8.3f} ms")
    print(f"E D2H {out_bytes/1e6:7.1f} MB                     {e:8.3f} ms  {out_bytes/e/1e6:6.1f} GB/s")
    print(f"F B + C + E concurrently               {f:8.3f} ms")


if __name__ == "__main__":
    main()
