"""Determinism / coverage stress of the fused step (development tool).

Two AllstepsMDP instances with the same seed and state are stepped on the same inputs, one eagerly and one through a
captured CUDA graph; every output buffer is pre-filled with a NaN sentinel before each step.  Any element that
differs between the two, or that still holds the sentinel, is reported.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

from allsteps_isaaclab_b200.mdp import AllstepsMDP, PhysicsViews, StepBuffers
from scenario import Scenario

NAMES = ("obs", "reward", "terminated", "time_out", "dones")


def run(N, steps, seed, use_graph, pdl_vs_plain=False):
    """pdl_vs_plain: the first instance is created with ALLSTEPS_PDL=0 (plain stream order between the step's kernels),
    the second with the default programmatic dependent launches -- any read that runs ahead of the kernel it depends
    on shows up as a difference."""
    sc = Scenario(N, seed=seed)
    st0 = sc.initial_mdp_state()
    origins = sc.env_origins.cuda()
    junk = torch.full((64 << 20,), float("nan"), device="cuda")  # poison the caching allocator's free blocks
    del junk
    mdps = []
    for i in range(2):
        if pdl_vs_plain and i == 0:
            os.environ["ALLSTEPS_PDL"] = "0"
        mdps.append(AllstepsMDP(N, device="cuda:0", seed=seed))
        os.environ.pop("ALLSTEPS_PDL", None)
    for m in mdps:
        m.generate_stones(origins)
        m.import_state({k: st0[k] for k in ("curr_target_index", "swing_leg", "target_reach_count",
                                            "episode_length_buf", "potentials")})
    stones = mdps[0].export_state()["steps_pos"].cpu()
    phys = sc.physics(stones, st0["curr_target_index"], st0["swing_leg"])
    static = {k: v.cuda() for k, v in phys.items()}
    views = PhysicsViews.from_dict(static, origins, sc.body_indices)
    outs = [StepBuffers(N, "cuda:0"), StepBuffers(N, "cuda:0")]
    step = mdps[0].capture_step(views, static["actions"], outs[0]) if use_graph else None
    bad_steps = 0
    for i in range(steps):
        st = mdps[1].export_state()
        phys = sc.physics(st["steps_pos"].cpu(), st["curr_target_index"].cpu(), st["swing_leg"].cpu())
        for k, v in phys.items():
            static[k].copy_(v)
        for o in outs:
            o.obs.fill_(float("nan"))
            o.reward.fill_(float("nan"))
        if step is not None:
            step.replay()
        else:
            mdps[0].step(views, static["actions"], outs[0])
        mdps[1].step(views, static["actions"], outs[1])
        torch.cuda.synchronize()
        msgs = []
        for name in NAMES:
            g, e = getattr(outs[0], name), getattr(outs[1], name)
            if g.dtype.is_floating_point:
                for tag, t in (("first", g), ("second", e)):
                    n_nan = int(torch.isnan(t).sum())
                    if n_nan:
                        where = torch.isnan(t).nonzero()[:6].tolist()
                        msgs.append(f"{name}[{tag}] holds {n_nan} unwritten/NaN elements, e.g. {where}")
            if not torch.equal(g, e):
                bad = (g != e).nonzero()
                vals = [(g[tuple(b)].item(), e[tuple(b)].item()) for b in bad[:6]]
                msgs.append(f"{name} differs at {len(bad)} places, first {bad[:6].tolist()}: {vals}")
        a, b = mdps[0].export_state(), mdps[1].export_state()
        for k in a:
            if not torch.equal(a[k], b[k]):
                bad = (a[k] != b[k]).nonzero()
                msgs.append(f"state {k} differs at {len(bad)} places, first {bad[:4].tolist()}")
        if msgs:
            bad_steps += 1
            flags = {n: getattr(outs[1], n) for n in ("terminated", "time_out")}
            print(f"N={N} graph={use_graph} step {i}: n_reset {int(outs[0].n_reset)}/{int(outs[1].n_reset)}")
            for m_ in msgs[:8]:
                print("   ", m_)
            env = None
            for name in ("obs",):
                g, e = getattr(outs[0], name), getattr(outs[1], name)
                bad = (g != e).nonzero()
                if len(bad):
                    env = int(bad[0][0])
            if env is not None:
                print(f"    env {env}: terminated={bool(flags['terminated'][env])} time_out={bool(flags['time_out'][env])} "
                      f"idx={int(b['curr_target_index'][env])} ep={int(b['episode_length_buf'][env])}")
    print(f"N={N} graph={use_graph} steps={steps}: {bad_steps} bad steps", flush=True)
    return bad_steps


if __name__ == "__main__":
    total = 0
    for N in (4096, 5000, 300, 20000):
        for g in (True, False):
            total += run(N, 40, seed=41, use_graph=g)
    # the sizes that use the separate gather kernel and the programmatic launch chain gather -> step -> finish
    for N in ((1 << 17) + 37, 1 << 20):
        total += run(N, 12, seed=43, use_graph=False, pdl_vs_plain=True)
        total += run(N, 12, seed=43, use_graph=True, pdl_vs_plain=True)
    sys.exit(1 if total else 0)
