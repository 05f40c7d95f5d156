"""Small cases of every kernel of the library through the public API, for compute-sanitizer:

    compute-sanitizer --tool memcheck  --error-exitcode 1 python tools/sanitize_cases.py
    compute-sanitizer --tool synccheck --error-exitcode 1 python tools/sanitize_cases.py
    compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize_cases.py --small

Logs of a run on B200 are kept under profiles/ (r02_sanitizer_*.txt).  Results are not checked here (the parity tests
do that); the point is that every launch of every kernel runs under the tool."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from allsteps_isaaclab_b200 import synthetic as syn
from allsteps_isaaclab_b200 import symmetry
from allsteps_isaaclab_b200.config import AllstepsCfg
from allsteps_isaaclab_b200.mdp import AllstepsMDP, PhysicsViews, StepBuffers
from allsteps_isaaclab_b200.workload import to_isaac_layout

SMALL = "--small" in sys.argv
dev = torch.device("cuda:0")
cfg = AllstepsCfg()


def make(N, seed=3, **kw):
    origins = syn.env_origins_grid(N, cfg.env_spacing).to(dev)
    mdp = AllstepsMDP(N, device=dev, seed=seed, **kw)
    mdp.generate_stones(origins)
    st0 = syn.random_mdp_state(cfg, N, torch.Generator().manual_seed(seed))
    st0["curr_target_index"] = torch.randint(9, 20, (N,), generator=torch.Generator().manual_seed(seed))
    mdp.import_state({k: st0[k] for k in ("curr_target_index", "swing_leg", "target_reach_count",
                                          "episode_length_buf", "potentials")})
    return mdp, origins


def physics(mdp, gen, origins, layout="dense", fall=0.05):
    st = mdp.export_state()
    d = syn.random_physics_state(cfg, st["steps_pos"], st["curr_target_index"], st["swing_leg"], gen, fall_fraction=fall)
    rows = (0, 1, 2)
    if layout == "isaac":
        d, rows = to_isaac_layout(d)
    return PhysicsViews.from_dict(d, origins, rows), d


def fused(N, steps, layout="dense", **kw):
    buf_kw = {k: kw.pop(k) for k in ("reward_terms", "obs_clip") if k in kw}
    mdp, origins = make(N, **kw)
    out = StepBuffers(N, dev, **buf_kw)
    gen = torch.Generator(device=dev).manual_seed(5)
    for _ in range(steps):
        v, d = physics(mdp, gen, origins, layout)
        mdp.step(v, d["actions"], out)
    torch.cuda.synchronize()
    return mdp, origins, out


def three_call(N, steps, device_list, fall=0.05):
    mdp, origins = make(N)
    out = StepBuffers(N, dev)
    gen = torch.Generator(device=dev).manual_seed(6)
    ep = torch.zeros(N, dtype=torch.int64, device=dev)
    for _ in range(steps):
        v, d = physics(mdp, gen, origins, fall=fall)
        ep += 1
        mdp.pass1(v, d["actions"], out, episode_length=ep)
        if device_list:
            mdp.reset(origins, None, out, episode_length=ep)
            mdp.pass2(v, out)
        else:
            ids = out.dones.nonzero().squeeze(-1)
            if len(ids):
                mdp.reset(origins, ids, out, episode_length=ep)
                mdp.pass2(v, out)
            else:
                mdp.no_reset()
    torch.cuda.synchronize()


big = 4096 + 77 if SMALL else (1 << 17) + 77
print("fused, ragged tile, per-term rewards, observation clamp", flush=True)
mdp, origins, out = fused(333, 4, reward_terms=True, obs_clip=5.0)
print("fused, no env resets (fix-up path)", flush=True)
m2, o2 = make(200)
b2 = StepBuffers(200, dev)
g2 = torch.Generator(device=dev).manual_seed(9)
for _ in range(3):
    v, d = physics(m2, g2, o2, fall=0.0)
    d["root_lin_vel_w"].mul_(0.1)
    m2.step(v, d["actions"], b2)
print(f"fused, {big} envs (k_prepare* + prepared instantiation + ragged tail)", flush=True)
fused(big, 3)
print(f"fused, {big} envs, Isaac Lab view layout (packed root tile, body rows gathered)", flush=True)
fused(big, 2, layout="isaac")
print("fused with stone regeneration / grid curriculum / missed-step termination", flush=True)
fused(1500, 3, intended_regen=True)
fused(1500, 3, grid_bins=7)
fused(700, 2, missed_step=True)
print("3-call path: host id list, device-side list, quiet steps", flush=True)
three_call(900, 4, False)
three_call(900, 4, True)
three_call(60, 4, False, fall=0.0)
three_call(60, 4, True, fall=0.0)
three_call(big, 2, True)
print("action path, mirror rows, state exchange, snapshot / restore, stone poses, peer self-exchange", flush=True)
mdp.apply_action(torch.randn(333, 21, device=dev))
symmetry.mirror_batch(mdp, torch.randn(1000, 59, device=dev), torch.randn(1000, 21, device=dev),
                      torch.randn(1000, 21, device=dev))
mdp.mirror_rows(torch.randn(77, 59, device=dev), "obs")
snap = mdp.snapshot(include_stones=True)
mdp.restore(snap, include_stones=True)
mdp.import_state(mdp.export_state())
mdp.export_stone_poses(torch.arange(0, 333, 3, device=dev))
mdp.export_stone_poses()
ms, os_ = make(500)
ms.connect_self()
bs = StepBuffers(500, dev)
gs = torch.Generator(device=dev).manual_seed(11)
for _ in range(3):
    v, d = physics(ms, gs, os_)
    ms.step(v, d["actions"], bs)
assert ms.peer_status()["timeouts"] == 0
torch.cuda.synchronize()
print("sanitize_cases done", flush=True)
