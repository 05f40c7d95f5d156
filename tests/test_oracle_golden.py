"""CPU: the oracle port against the golden vectors produced by executing the unmodified reference
(tests/golden/make_golden.py).  Everything must be bit-identical: same torch ops in the same order."""
from __future__ import annotations

import numpy as np
import pytest
import torch

import golden_util as gu
from allsteps_isaaclab_b200.config import AllstepsCfg
from oracle import allsteps_oracle as ao
from scenario import install_mdp_state


def same(a: torch.Tensor, b, what):
    b = gu.t(b)
    assert a.shape == b.shape, f"{what}: {tuple(a.shape)} vs {tuple(b.shape)}"
    assert torch.equal(a, b), f"{what}: differs at {(a != b).nonzero()[:3].tolist()}"


@pytest.mark.parametrize("name", gu.REPLAYS)
def test_port_reproduces_reference_replay(name):
    d = gu.load(name)
    cfg = AllstepsCfg()
    N = int(d["num_envs"])
    orc = ao.AllstepsOracle(cfg, N, gu.t(d["env_origins"]), gu.t(d["joint_limits"]), tuple(d["body_indices"]),
                            gu.t(d["init_stone_uniforms"]))
    same(orc.steps_pos, d["init_steps_pos"], "initial steps_pos")
    same(orc.steps_dphi, d["init_steps_dphi"], "initial steps_dphi")
    install_mdp_state(orc, gu.initial_state(d))
    n_reset = n_quiet = promoted = 0
    for step in range(int(d["steps"])):
        if f"s{step}_forced_index" in d:
            orc.curr_target_index[:] = gu.t(d[f"s{step}_forced_index"])
            orc.prev_target_index = torch.clamp(orc.curr_target_index - 1, 0, 19)
            orc.next_target_index = torch.clamp(orc.curr_target_index + 1, 0, 19)
        phys = gu.step_inputs(d, step)
        level0 = int(orc.curriculum[0])
        obs, rew, term, to, ids = orc.step(phys, phys["actions"], gu.t(d[f"s{step}_mirror_u"]),
                                           gu.t(d[f"s{step}_noise_u"]), None)
        same(obs, d[f"s{step}_obs"], f"step {step} obs")
        same(rew, d[f"s{step}_reward"], f"step {step} reward")
        same(term, d[f"s{step}_terminated"], f"step {step} terminated")
        same(to, d[f"s{step}_time_out"], f"step {step} time_out")
        same(ids, d[f"s{step}_reset_ids"], f"step {step} reset ids")
        for k in gu.STATE_KEYS + ["old_potentials"]:
            same(getattr(orc, k), d[f"s{step}_{k}"], f"step {step} {k}")
        if len(ids):
            n_reset += 1
            w = orc.reset_writes
            same(w["root_pose"], d[f"s{step}_w_root_pose"], f"step {step} root pose write")
            same(w["root_velocity"], d[f"s{step}_w_root_velocity"], f"step {step} root velocity write")
            same(w["joint_pos"], d[f"s{step}_w_joint_pos"], f"step {step} joint pos write")
            same(w["joint_vel"], d[f"s{step}_w_joint_vel"], f"step {step} joint vel write")
        else:
            n_quiet += 1
        promoted += int(int(orc.curriculum[0]) != level0)
    if name.endswith("n64.npz"):
        assert n_reset > 0 and promoted > 0
    else:
        assert n_quiet > 0


def test_stone_generation_all_levels():
    d = gu.load("stones_levels.npz")
    pos, dphi = ao.generate_stones(AllstepsCfg(), gu.t(d["levels"]), gu.t(d["uniforms"]))
    same(pos, d["pos_local"], "stone positions")
    same(dphi, d["dphi"], "stone cumulative yaw")
    # level 0 is a straight flat line 0.75 m apart and the first three stones never depend on the level
    lvl0 = (gu.t(d["levels"]) == 0).nonzero().flatten()
    assert torch.allclose(pos[lvl0, :, 0], 0.75 * torch.arange(20.0).expand(len(lvl0), 20), atol=1e-5)
    assert torch.equal(pos[:, :3], pos[0:1, :3].expand(pos.shape[0], 3, 3))


def test_math_helpers_match_reference():
    d = gu.load("math_helpers.npz")
    q, v, p, t = (gu.t(d[k]) for k in ("q", "v", "p", "t"))
    roll, pitch, yaw = ao.euler_xyz_wrapped(q)
    same(roll, d["roll"], "roll")
    same(pitch, d["pitch"], "pitch")
    same(yaw, d["yaw"], "yaw")
    assert float(roll.min()) >= 0.0 and float(roll.max()) <= 2 * np.pi + 1e-6  # wrapped to [0, 2*pi)
    same(ao.rotate_by_inverse(q, v), d["rotate_inverse"], "quat_rotate_inverse")
    same(ao.point_in_frame(p, q, t), d["frame_point"], "subtract_frame_transforms")
    same(ao.scale_to_unit(gu.t(d["x"]), gu.t(d["lo"]), gu.t(d["hi"])), d["scaled"], "scale_transform")
    same(ao.unscale_from_unit(gu.t(d["x"]), gu.t(d["lo"]), gu.t(d["hi"])), d["unscaled"], "unscale_transform")


def test_quat_rotate_inverse_known_answers():
    """The reference's own unit test pins quat_rotate_inverse against the older bmm formulation
    (source/isaaclab/test/utils/test_math.py:278-424); restated here as the closed form R(q)^T v."""
    g = torch.Generator().manual_seed(0)
    q = torch.randn(1024, 4, generator=g, dtype=torch.float64)
    q = q / q.norm(dim=-1, keepdim=True)
    v = torch.randn(1024, 3, generator=g, dtype=torch.float64)
    w, x, y, z = q.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                     2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                     2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1).view(-1, 3, 3)
    expect = torch.einsum("nji,nj->ni", R, v)
    got = ao.rotate_by_inverse(q.float(), v.float()).double()
    assert torch.allclose(got, expect, atol=1e-5)
    # round trip through the frame transform (test_math.py:447-469 combine/subtract round trip)
    p = torch.randn(1024, 3, generator=g)
    local = ao.point_in_frame(p, q.float(), p + torch.einsum("nij,nj->ni", R.float(), v.float()))
    assert torch.allclose(local, v.float(), atol=1e-4)


def test_symmetric_states_port_against_the_reference_fixture():
    """tests/golden/mirror_symmetry.npz = outputs of the reference's own get_symmetric_states_* (ENV:570-660)."""
    from allsteps_isaaclab_b200.config import AllstepsCfg
    from oracle import allsteps_oracle as ao

    d = gu.load("mirror_symmetry.npz")
    cfg = AllstepsCfg()
    tabs = (cfg.right_joint_indices, cfg.left_joint_indices, cfg.negation_joint_indices)
    f = lambda k: torch.from_numpy(d[k].view(np.float32).copy())  # noqa: E731
    bits = lambda t: t.numpy().view(np.uint32)  # noqa: E731
    assert np.array_equal(bits(ao.symmetric_states(f("obs"), *tabs, "obs")), d["rl_games_obs"])
    assert np.array_equal(bits(ao.symmetric_states(f("actions"), *tabs, "actions")), d["rl_games_actions"])
    assert np.array_equal(bits(ao.symmetric_states(f("mus"), *tabs, "actions")), d["rl_games_mus"])
    assert np.array_equal(d["rsl_rl_obs"], d["rl_games_obs"]) and np.array_equal(d["rsl_rl_actions"], d["rl_games_actions"])
    assert ao.symmetric_states(None, *tabs, "obs") is None
