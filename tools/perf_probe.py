"""Quick device-time probe of the fused step at several env counts (development tool, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from allsteps_isaaclab_b200.config import AllstepsCfg
from allsteps_isaaclab_b200 import synthetic as syn
from allsteps_isaaclab_b200.mdp import AllstepsMDP, PhysicsViews, StepBuffers

B_ALG = 652


def probe(N, steps=50, sets=4, warmup=10):
    cfg = AllstepsCfg()
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(1234)
    origins = syn.env_origins_grid(N, cfg.env_spacing).to(dev)
    mdp = AllstepsMDP(N, device=dev, seed=1, skip_pass2=("--skip-pass2" in sys.argv))
    mdp.generate_stones(origins)
    st0 = syn.random_mdp_state(cfg, N, torch.Generator().manual_seed(1))
    mdp.import_state({k: st0[k] for k in ("curr_target_index", "swing_leg", "target_reach_count",
                                          "episode_length_buf", "potentials")})
    st = mdp.export_state()
    pool = []
    isaac = "--isaac-views" in sys.argv  # slices of root_state_w (N,13) and body_state_w (N,17,13), as Isaac Lab hands out
    for s in range(sets):
        d = syn.random_physics_state(cfg, st["steps_pos"], st["curr_target_index"], st["swing_leg"], gen)
        rows = (0, 1, 2)
        if isaac:
            root_state = torch.zeros(N, 13, device=dev)
            root_state[:, 0:3], root_state[:, 3:7], root_state[:, 7:10] = d["root_pos_w"], d["root_quat_w"], d["root_lin_vel_w"]
            rows = (16, 13, 0)  # right_foot, left_foot, torso among 17 bodies
            body_state = torch.zeros(N, 17, 13, device=dev)
            for k, r in enumerate(rows):
                body_state[:, r, 0:3] = d["body_pos_w"][:, k]
            d["root_pos_w"], d["root_quat_w"], d["root_lin_vel_w"] = root_state[:, 0:3], root_state[:, 3:7], root_state[:, 7:10]
            d["body_pos_w"] = body_state[..., 0:3]
            d["_keep"] = (root_state, body_state)
        pool.append((PhysicsViews.from_dict(d, origins, rows), d))
    out = StepBuffers(N, dev, reset_rows=("--rows" in sys.argv))
    for i in range(warmup):
        v, d = pool[i % sets]
        mdp.step(v, d["actions"], out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        v, d = pool[i % sets]
        mdp.step(v, d["actions"], out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if "--timing" in sys.argv:
        import ctypes
        buf = (ctypes.c_uint64 * 16)()
        mdp.lib.as_debug_timing(mdp.handle, buf, 1, mdp._stream())
        for i in range(warmup):
            v, d = pool[i % sets]; mdp.step(v, d["actions"], out)
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
          f"resets/step={stats['n_reset']}  launches/step={mdp.launch_count/(steps+warmup):.1f}", flush=True)


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
