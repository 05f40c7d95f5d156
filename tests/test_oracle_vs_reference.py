"""CPU, build container only: the oracle port against the UNMODIFIED reference executed live (skipped where
/root/reference is not mounted, e.g. on the GPU box -- the golden fixtures cover that case)."""
from __future__ import annotations

import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference checkout not mounted")


@pytest.mark.parametrize("num_envs,steps,high", [(64, 25, False), (1024, 10, True)])
def test_port_is_bit_identical_to_live_reference(num_envs, steps, high):
    from allsteps_isaaclab_b200.config import BODY_NAMES, JOINT_NAMES
    from oracle import allsteps_oracle as ao
    from oracle import ref_fake_env as rf
    from scenario import Scenario, install_mdp_state

    sc = Scenario(num_envs, seed=31 + num_envs, full_bodies=True)
    su = sc.stone_uniforms(0)
    st0 = sc.initial_mdp_state()
    if high:
        st0["curr_target_index"] = torch.randint(12, 20, (num_envs,), generator=sc.gen)
    orc = ao.AllstepsOracle(sc.cfg, num_envs, sc.env_origins, sc.joint_limits, sc.body_indices, su)
    phys = sc.physics(orc.steps_pos, st0["curr_target_index"], st0["swing_leg"])
    world = dict(phys)
    world["env_origins"] = sc.env_origins
    world["joint_pos_limits"] = sc.joint_limits.unsqueeze(0).repeat(num_envs, 1, 1)
    ref = rf.make_reference_env(world, sc.cfg, BODY_NAMES, JOINT_NAMES, su)
    assert torch.equal(ref.steps_pos, orc.steps_pos) and torch.equal(ref.steps_dphi, orc.steps_dphi)
    install_mdp_state(ref, st0)
    install_mdp_state(orc, st0)
    for step in range(steps):
        phys = sc.physics(orc.steps_pos, orc.curr_target_index, orc.swing_leg)
        m, n = sc.reset_uniforms(step)
        rf.load_physics(ref, phys)
        r = rf.step_mdp(ref, phys["actions"], rf.UniformTables(m, n, sc.stone_uniforms(step)))
        o = orc.step(phys, phys["actions"], m, n, sc.stone_uniforms(step))
        for a, b, what in zip(r, o, ("obs", "reward", "terminated", "time_out", "reset ids")):
            assert torch.equal(a, b), f"step {step}: {what}"
        for k in ("curr_target_index", "prev_target_index", "next_target_index", "swing_leg", "target_reach_count",
                  "episode_length_buf", "curriculum", "potentials", "old_potentials", "targets_w", "targets_b",
                  "foot_contact"):
            assert torch.equal(getattr(ref, k), getattr(orc, k)), f"step {step}: {k}"
        if len(r[4]):
            c = ref.robot.rec.calls
            assert torch.equal(c["root_pose"][0], orc.reset_writes["root_pose"])
            assert torch.equal(c["joint_state"][0], orc.reset_writes["joint_pos"])
        assert torch.equal(ref.applied_gain_curriculum[ref.curriculum].unsqueeze(-1) * ref.joint_gears.unsqueeze(0)
                           * ref.actions, orc.joint_efforts())


def test_port_propagates_nan_inputs_like_the_live_reference():
    """NaN in actions / joint state / foot height / root quaternion / root velocity: the port must show NaN exactly
    where the executed reference does (torch.clamp / torch.minimum hand NaN through, comparisons with NaN are false).
    This pins the oracle that tests/test_gpu_parity.py::test_nan_inputs_propagate_like_torch holds the kernels to."""
    from allsteps_isaaclab_b200.config import BODY_NAMES, JOINT_NAMES
    from oracle import allsteps_oracle as ao
    from oracle import ref_fake_env as rf
    from scenario import Scenario, install_mdp_state

    num_envs = 256
    sc = Scenario(num_envs, seed=77, full_bodies=True, fall_fraction=0.0)
    su = sc.stone_uniforms(0)
    st0 = sc.initial_mdp_state()
    orc = ao.AllstepsOracle(sc.cfg, num_envs, sc.env_origins, sc.joint_limits, sc.body_indices, su)
    phys = sc.physics(orc.steps_pos, st0["curr_target_index"], st0["swing_leg"])
    world = dict(phys)
    world["env_origins"] = sc.env_origins
    world["joint_pos_limits"] = sc.joint_limits.unsqueeze(0).repeat(num_envs, 1, 1)
    ref = rf.make_reference_env(world, sc.cfg, BODY_NAMES, JOINT_NAMES, su)
    install_mdp_state(ref, st0)
    install_mdp_state(orc, st0)
    nan = float("nan")
    saw_nan = False
    for step in range(3):
        phys = sc.physics(orc.steps_pos, orc.curr_target_index, orc.swing_leg)
        phys["actions"][3, 2] = nan
        phys["joint_vel"][7, 5] = nan
        phys["joint_pos"][9, 0] = nan
        phys["body_pos_w"][11, sc.body_indices[0], 2] = nan
        phys["body_pos_w"][12, sc.body_indices[1], 2] = nan
        phys["root_quat_w"][13, 1] = nan
        phys["root_lin_vel_w"][14, 0] = nan
        m, n = sc.reset_uniforms(step)
        rf.load_physics(ref, phys)
        r = rf.step_mdp(ref, phys["actions"], rf.UniformTables(m, n, sc.stone_uniforms(step)))
        o = orc.step(phys, phys["actions"], m, n, sc.stone_uniforms(step))
        for a, b, what in zip(r, o, ("obs", "reward", "terminated", "time_out", "reset ids")):
            assert torch.equal(torch.isnan(a), torch.isnan(b)) if a.dtype.is_floating_point else True, what
            if a.dtype.is_floating_point:
                saw_nan |= bool(torch.isnan(a).any())
                ok = ~torch.isnan(a)
                assert torch.equal(a[ok], b[ok]), f"step {step}: {what}"
            else:
                assert torch.equal(a, b), f"step {step}: {what}"
        for k in ("curr_target_index", "swing_leg", "target_reach_count", "episode_length_buf"):
            assert torch.equal(getattr(ref, k), getattr(orc, k)), f"step {step}: {k}"
    assert saw_nan


def test_reference_cfg_constants_match_product_config():
    """The product's constants (allsteps_isaaclab_b200/config.py) against the reference's own cfg class / env."""
    from allsteps_isaaclab_b200.config import AllstepsCfg

    ref = ref_loader.load_reference()
    R, C = ref.AllstepsEnvCfg, AllstepsCfg()
    assert R.num_steps == C.num_steps and R.step_radius == C.step_radius
    assert list(R.joint_gears) == list(C.joint_gears)
    assert tuple(R.right_body_names) == C.right_body_names and tuple(R.left_body_names) == C.left_body_names
    assert tuple(R.negation_body_names) == C.negation_body_names
    for k in ("energy_cost_scale", "actions_cost_scale", "alive_reward_scale", "dof_vel_scale",
              "joint_at_limit_cost_scale", "death_cost", "termination_height_absolute", "episode_length_s",
              "decimation"):
        assert getattr(R, k) == getattr(C, k), k
    assert tuple(R.initial_joint_angle_range) == C.initial_joint_angle_range
    assert tuple(R.initial_joint_angle_clip_range) == C.initial_joint_angle_clip_range
    assert ref.env_module.EPSILON == C.contact_epsilon


def test_symmetric_states_port_equals_the_live_reference_functions():
    """`get_symmetric_states_rl_games` / `_rsl_rl` (ENV:570-660) executed from the reference checkout on
    a fake wrapped env (three index tensors + two batched spaces) against the oracle's restatement, bit for bit."""
    import numpy as np

    from allsteps_isaaclab_b200.config import AllstepsCfg
    from oracle import allsteps_oracle as ao

    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as mg

    ref = ref_loader.load_reference()
    cfg = AllstepsCfg()
    env = mg.fake_wrapped_env(cfg)
    g = torch.Generator().manual_seed(77)
    obs, actions, mus = torch.randn(333, 59, generator=g), torch.randn(333, 21, generator=g), torch.randn(333, 21, generator=g)
    obs[0, :] = float("nan")
    obs[1, :] = -0.0
    tabs = (cfg.right_joint_indices, cfg.left_joint_indices, cfg.negation_joint_indices)
    bits = lambda t: t.numpy().view(np.uint32)  # noqa: E731
    o, a, m = ref.env_module.get_symmetric_states_rl_games(obs, actions, env, False, mus)
    assert np.array_equal(bits(o), bits(ao.symmetric_states(obs, *tabs, "obs")))
    assert np.array_equal(bits(a), bits(ao.symmetric_states(actions, *tabs, "actions")))
    assert np.array_equal(bits(m), bits(ao.symmetric_states(mus, *tabs, "actions")))
    o, a = ref.env_module.get_symmetric_states_rsl_rl(obs, actions, env)
    assert np.array_equal(bits(o), bits(ao.symmetric_states(obs, *tabs, "obs")))
    assert np.array_equal(bits(a), bits(ao.symmetric_states(actions, *tabs, "actions")))
    # the env's index tensors are what config.py derives from the names (CFG:217-219)
    assert env.unwrapped.right_body_indices.tolist() == [2, 3, 4, 9, 11, 12, 13, 17, 19]
    assert env.unwrapped.left_body_indices.tolist() == [5, 6, 7, 10, 14, 15, 16, 18, 20]
    assert env.unwrapped.negation_body_indices.tolist() == [0, 8]
    # and the committed fixture is what the live reference produces now
    import golden_util as gu

    d = gu.load("mirror_symmetry.npz")
    fresh = mg.mirror_symmetry()
    for k in fresh:
        assert np.array_equal(fresh[k], d[k]), k


def test_stone_pose_fixture_is_what_the_live_reference_writes():
    """tests/golden/stone_poses_view.npz against a fresh execution of the reference's
    RigidObjectCollection.write_object_pose_to_sim (oracle/ref_rigid_collection.py), plus the closed form the CUDA
    kernel implements: view row s*N + e = (x, y, z, 0, 0, 0, 1), index list object-major."""
    import numpy as np

    import golden_util as gu
    from oracle import ref_rigid_collection as rc

    d = gu.load("stone_poses_view.npz")
    steps_pos, env_ids = gu.t(d["steps_pos"]), gu.t(d["env_ids"])
    poses, view_ids = rc.reference_stone_pose_write(steps_pos, env_ids)
    assert np.array_equal(view_ids.numpy(), d["view_ids"]) and np.array_equal(poses[view_ids].numpy(), d["rows"])
    N, S = steps_pos.shape[:2]
    want_ids = (torch.arange(S).unsqueeze(1) * N + env_ids).flatten()
    assert torch.equal(view_ids, want_ids)
    rows = poses[view_ids].reshape(S, len(env_ids), 7)
    assert torch.equal(rows[..., :3], steps_pos[env_ids].transpose(0, 1))
    assert torch.equal(rows[..., 3:], torch.tensor([0.0, 0.0, 0.0, 1.0]).expand(S, len(env_ids), 4))
