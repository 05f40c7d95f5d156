"""GPU: the CUDA path against the committed golden vectors (outputs of the unmodified reference)."""
from __future__ import annotations

import pytest
import torch

import golden_util as gu

pytestmark = pytest.mark.gpu


def _close(a, b, what, rtol=1e-5):
    a = a.detach().cpu().double()
    b = gu.t(b).double()
    assert a.shape == b.shape, what
    err = ((a - b).abs() / b.abs().clamp(min=1.0)).max().item() if a.numel() else 0.0
    assert err <= rtol, f"{what}: {err:.3e}"


def _exact(a, b, what):
    a = a.detach().cpu()
    b = gu.t(b)
    assert torch.equal(a.to(b.dtype), b), f"{what}: mismatch at {(a.to(b.dtype) != b).nonzero()[:3].tolist()}"


@pytest.mark.parametrize("name", gu.REPLAYS)
def test_cuda_reproduces_reference_replay(name):
    from allsteps_isaaclab_b200.mdp import AllstepsMDP, PhysicsViews, StepBuffers

    d = gu.load(name)
    N = int(d["num_envs"])
    body_rows = tuple(int(x) for x in d["body_indices"])
    # the fixture's uniforms were produced by the Philox twin with seed = fixture seed, step = step index
    mdp = AllstepsMDP(N, device="cuda:0", seed=int(d["seed"]))
    origins = gu.t(d["env_origins"]).cuda()
    mdp.generate_stones(origins)
    st = mdp.export_state()
    _close(st["steps_pos"], d["init_steps_pos"], "initial steps_pos")
    _close(st["steps_dphi"], d["init_steps_dphi"], "initial steps_dphi")
    init = gu.initial_state(d)
    mdp.import_state({"curr_target_index": init["curr_target_index"], "swing_leg": init["swing_leg"],
                      "target_reach_count": init["target_reach_count"],
                      "episode_length_buf": init["episode_length_buf"], "curriculum": init["curriculum"],
                      "potentials": init["potentials"], "steps_pos": gu.t(d["init_steps_pos"]),
                      "steps_dphi": gu.t(d["init_steps_dphi"])})
    out = StepBuffers(N, "cuda:0")
    for step in range(int(d["steps"])):
        if f"s{step}_forced_index" in d:
            mdp.import_state({"curr_target_index": gu.t(d[f"s{step}_forced_index"])})
        phys = {k: v.cuda() for k, v in gu.step_inputs(d, step).items()}
        mdp.step(PhysicsViews.from_dict(phys, origins, body_rows), phys["actions"], out)
        torch.cuda.synchronize()
        _exact(out.terminated, d[f"s{step}_terminated"], f"step {step} terminated")
        _exact(out.time_out, d[f"s{step}_time_out"], f"step {step} time_out")
        _close(out.reward, d[f"s{step}_reward"], f"step {step} reward")
        _close(out.obs, d[f"s{step}_obs"], f"step {step} obs")
        st = mdp.export_state()
        for k in ("curr_target_index", "prev_target_index", "next_target_index", "swing_leg",
                  "target_reach_count", "episode_length_buf", "curriculum"):
            _exact(st[k], d[f"s{step}_{k}"], f"step {step} {k}")
        _close(st["potentials"], d[f"s{step}_potentials"], f"step {step} potentials")
        ids = gu.t(d[f"s{step}_reset_ids"])
        n = int(out.n_reset.item())
        assert n == len(ids)
        if n:
            got = out.reset_ids[:n].long().sort().values
            _exact(got, ids, f"step {step} reset ids")
            root = torch.cat((gu.t(d[f"s{step}_w_root_pose"]), gu.t(d[f"s{step}_w_root_velocity"])), -1)
            _close(out.reset_root_state[got], root, f"step {step} root rows")
            _close(out.reset_joint_pos[got], d[f"s{step}_w_joint_pos"], f"step {step} joint_pos rows")
            _close(out.reset_joint_vel[got], d[f"s{step}_w_joint_vel"], f"step {step} joint_vel rows")


def test_cuda_stones_match_reference_at_every_level():
    from allsteps_isaaclab_b200.mdp import AllstepsMDP

    d = gu.load("stones_levels.npz")
    N = d["levels"].shape[0]
    mdp = AllstepsMDP(N, device="cuda:0", seed=0)
    mdp.import_state({"curriculum": gu.t(d["levels"])})
    mdp.generate_stones(torch.zeros(N, 3, device="cuda"), uniforms=gu.t(d["uniforms"]).cuda())
    st = mdp.export_state()
    _close(st["steps_pos"], d["pos_local"], "stone positions")
    _close(st["steps_dphi"], d["dphi"], "cumulative yaw")


def test_stone_pose_egress_equals_the_reference_method():
    """`as_export_stone_poses` against what the reference's own
    `RigidObjectCollection.write_object_pose_to_sim` (rigid_object_collection.py:271-301, executed unmodified by
    tests/golden/make_golden.py through oracle/ref_rigid_collection.py) hands to `root_physx_view.set_transforms` for
    the stone write of ENV:119-120: the same index list and, at those indices, the same x,y,z,w pose rows."""
    from allsteps_isaaclab_b200.mdp import AllstepsMDP

    d = gu.load("stone_poses_view.npz")
    steps_pos = gu.t(d["steps_pos"]).cuda()
    N = steps_pos.shape[0]
    mdp = AllstepsMDP(N, device="cuda:0", seed=1)
    mdp.import_state({"steps_pos": steps_pos, "steps_dphi": torch.zeros(N, 20)})
    env_ids = gu.t(d["env_ids"]).cuda()
    poses, view_ids = mdp.export_stone_poses(env_ids)
    torch.cuda.synchronize()
    assert torch.equal(view_ids.cpu().long(), gu.t(d["view_ids"]))
    assert torch.equal(poses[view_ids.long()].cpu(), gu.t(d["rows"]))
    poses, view_ids = mdp.export_stone_poses()
    torch.cuda.synchronize()
    assert torch.equal(view_ids.cpu().long(), gu.t(d["view_ids_all"]))
    assert torch.equal(poses.cpu(), gu.t(d["poses_all"]))
