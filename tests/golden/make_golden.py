"""Generates tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (needs /root/reference; run in the build
container):   python tests/golden/make_golden.py

allsteps_replay_n16_quiet.npz  12 steps of 16 envs where nothing falls: most steps have no reset (single pass).
allsteps_replay_n64.npz   16 consecutive MDP steps of 64 envs through the reference's own hook methods
                          (oracle/ref_fake_env.py hosts them): per-step synthetic physics inputs, injected uniforms,
                          and every output / MDP buffer after the step, plus the arguments of the three PhysX writes.
stones_levels.npz         `_generate_foot_steps_allsteps` at curriculum levels 0..9 for given uniforms.
mirror_symmetry.npz       `get_symmetric_states_rl_games` / `get_symmetric_states_rsl_rl` (ENV:570-660) on random
                          observation / action / mu batches (incl. -0.0, inf and NaN entries).
stone_poses_view.npz      what `RigidObjectCollection.write_object_pose_to_sim` (rigid_object_collection.py:271-301)
                          hands to `root_physx_view.set_transforms` for the stone egress of ENV:119-120.
math_helpers.npz          euler_xyz_from_quat / quat_rotate_inverse / subtract_frame_transforms / scale_transform /
                          unscale_transform of the reference's utils/math.py on random inputs.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from scenario import Scenario, install_mdp_state  # noqa: E402
from allsteps_isaaclab_b200.config import BODY_NAMES, JOINT_NAMES  # noqa: E402
from oracle import ref_fake_env as rf  # noqa: E402
from oracle import ref_loader  # noqa: E402

PHYS_KEYS = ["root_pos_w", "root_quat_w", "root_lin_vel_w", "root_ang_vel_w", "body_pos_w", "joint_pos", "joint_vel",
             "force_matrix_right", "force_matrix_left", "actions"]
STATE_KEYS = ["curr_target_index", "prev_target_index", "next_target_index", "swing_leg", "target_reach_count",
              "episode_length_buf", "curriculum", "potentials", "old_potentials"]


def replay(num_envs=64, steps=16, seed=2024, high_index_from=8, fall_fraction=0.02):
    sc = Scenario(num_envs, seed=seed, full_bodies=True, fall_fraction=fall_fraction)
    su = sc.stone_uniforms(0)
    st0 = sc.initial_mdp_state()
    stones0 = None
    phys0 = None
    world = None
    out = {"num_envs": num_envs, "steps": steps, "seed": seed}
    # build the reference env (its __init__ generates the stones from `su`)
    dummy_stones = torch.zeros(num_envs, 20, 3) + sc.env_origins[:, None, :]
    phys0 = sc.physics(dummy_stones, st0["curr_target_index"], st0["swing_leg"])
    world = dict(phys0)
    world["env_origins"] = sc.env_origins
    world["joint_pos_limits"] = sc.joint_limits.unsqueeze(0).repeat(num_envs, 1, 1)
    env = rf.make_reference_env(world, sc.cfg, BODY_NAMES, JOINT_NAMES, su)
    install_mdp_state(env, st0)
    out["init_stone_uniforms"] = su.numpy()
    out["init_steps_pos"] = env.steps_pos.numpy().copy()
    out["init_steps_dphi"] = env.steps_dphi.numpy().copy()
    for k in ("curr_target_index", "swing_leg", "target_reach_count", "episode_length_buf", "curriculum",
              "potentials"):
        out[f"init_{k}"] = st0[k].numpy().copy()
    out["env_origins"] = sc.env_origins.numpy()
    out["joint_limits"] = sc.joint_limits.numpy()
    out["body_indices"] = np.array(sc.body_indices)
    for step in range(steps):
        if step == high_index_from:
            # push most envs far along so that mean(curr_target_index) > 12 and the promotion rule fires (ENV:471)
            env.curr_target_index[:] = torch.randint(13, 20, (num_envs,), generator=sc.gen)
            env.prev_target_index = torch.clamp(env.curr_target_index - 1, 0, 19)
            env.next_target_index = torch.clamp(env.curr_target_index + 1, 0, 19)
            out[f"s{step}_forced_index"] = env.curr_target_index.numpy().copy()
        phys = sc.physics(env.steps_pos, env.curr_target_index, env.swing_leg)
        m, n = sc.reset_uniforms(step)
        rf.load_physics(env, phys)
        obs, rew, term, to, ids = rf.step_mdp(env, phys["actions"], rf.UniformTables(m, n, sc.stone_uniforms(step)))
        for k in PHYS_KEYS:
            out[f"s{step}_in_{k}"] = phys[k].numpy().copy()
        out[f"s{step}_mirror_u"] = m.numpy().copy()
        out[f"s{step}_noise_u"] = n.numpy().copy()
        out[f"s{step}_obs"] = obs.numpy().copy()
        out[f"s{step}_reward"] = rew.numpy().copy()
        out[f"s{step}_terminated"] = term.numpy().copy()
        out[f"s{step}_time_out"] = to.numpy().copy()
        out[f"s{step}_reset_ids"] = ids.numpy().copy()
        for k in STATE_KEYS:
            out[f"s{step}_{k}"] = getattr(env, k).numpy().copy()
        if len(ids):
            c = env.robot.rec.calls
            out[f"s{step}_w_root_pose"] = c["root_pose"][0].numpy().copy()
            out[f"s{step}_w_root_velocity"] = c["root_velocity"][0].numpy().copy()
            out[f"s{step}_w_joint_pos"] = c["joint_state"][0].numpy().copy()
            out[f"s{step}_w_joint_vel"] = c["joint_state"][1].numpy().copy()
    return out


def stones_levels(num_envs=40, seed=7):
    ref = ref_loader.load_reference()
    sc = Scenario(num_envs, seed=seed, full_bodies=True)
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(5, num_envs, 20, generator=g)
    levels = torch.arange(num_envs) % 10
    st0 = sc.initial_mdp_state()
    phys0 = sc.physics(torch.zeros(num_envs, 20, 3), st0["curr_target_index"], st0["swing_leg"])
    world = dict(phys0)
    world["env_origins"] = sc.env_origins
    world["joint_pos_limits"] = sc.joint_limits.unsqueeze(0).repeat(num_envs, 1, 1)
    env = rf.make_reference_env(world, sc.cfg, BODY_NAMES, JOINT_NAMES, u)
    env.curriculum[:] = levels
    with rf.injected_uniforms(env, rf.UniformTables(stones=u)):
        pos, dphi, legs = env._generate_foot_steps_allsteps()
    return {"uniforms": u.numpy(), "levels": levels.numpy(), "pos_local": pos.numpy(), "dphi": dphi.numpy(),
            "swing_legs": legs.numpy()}


def math_helpers(n=256, seed=11):
    ref = ref_loader.load_reference()
    M = ref.math
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(n, 4, generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    q[0] = torch.tensor([1.0, 0.0, 0.0, 0.0])
    q[1] = torch.tensor([1.0, -0.0, -0.0, -0.0])
    q[2] = torch.tensor([0.70710678, 0.0, 0.70710678, 0.0])  # sin_pitch = 1: the copysign branch
    q[3] = torch.tensor([0.70710678, 0.0, -0.70710678, 0.0])
    v = torch.randn(n, 3, generator=g)
    p = 10 * torch.randn(n, 3, generator=g)
    t = 10 * torch.randn(n, 3, generator=g)
    x = torch.randn(n, 21, generator=g)
    lo = -1.0 - torch.rand(21, generator=g)
    hi = 0.5 + torch.rand(21, generator=g)
    roll, pitch, yaw = M.euler_xyz_from_quat(q)
    return {"q": q.numpy(), "v": v.numpy(), "p": p.numpy(), "t": t.numpy(), "x": x.numpy(), "lo": lo.numpy(),
            "hi": hi.numpy(), "roll": roll.numpy(), "pitch": pitch.numpy(), "yaw": yaw.numpy(),
            "rotate_inverse": M.quat_rotate_inverse(q, v).numpy(),
            "frame_point": M.subtract_frame_transforms(p, q, t)[0].numpy(),
            "scaled": M.scale_transform(x, lo, hi).numpy(),
            "unscaled": M.unscale_transform(x, lo, hi).numpy()}


def fake_wrapped_env(cfg, num_envs=8, device="cpu"):
    """What the reference's symmetry helpers read from `env` (ENV:574-584): three index tensors and two batched spaces."""
    import types

    base = types.SimpleNamespace(
        right_body_indices=torch.tensor(cfg.right_joint_indices, dtype=torch.int64, device=device),
        left_body_indices=torch.tensor(cfg.left_joint_indices, dtype=torch.int64, device=device),
        negation_body_indices=torch.tensor(cfg.negation_joint_indices, dtype=torch.int64, device=device),
        observation_space=types.SimpleNamespace(shape=(num_envs, 59)),
        action_space=types.SimpleNamespace(shape=(num_envs, 21)), device=device)
    return types.SimpleNamespace(unwrapped=base, device=device)


def mirror_symmetry(rows=96, seed=31):
    ref = ref_loader.load_reference()
    from allsteps_isaaclab_b200.config import AllstepsCfg

    g = torch.Generator().manual_seed(seed)
    obs = 3.0 * torch.randn(rows, 59, generator=g)
    actions = torch.randn(rows, 21, generator=g)
    mus = torch.randn(rows, 21, generator=g)
    specials = torch.tensor([0.0, -0.0, float("inf"), -float("inf"), float("nan")])
    for r in range(5):   # special values in every column kind: kept / swapped / negated
        obs[r, :] = specials[r]
        actions[r, :] = specials[r]
        mus[r, ::2] = specials[r]
    env = fake_wrapped_env(AllstepsCfg())
    o2, a2, m2 = ref.env_module.get_symmetric_states_rl_games(obs, actions, env, False, mus)
    o3, a3 = ref.env_module.get_symmetric_states_rsl_rl(obs, actions, env)
    n1, n2, n3 = ref.env_module.get_symmetric_states_rl_games(None, actions, env, False, None)
    assert n1 is None and n3 is None
    as_bits = lambda t: t.numpy().view(np.uint32).copy()  # noqa: E731  (bit patterns: NaN-safe comparisons)
    return {"obs": as_bits(obs), "actions": as_bits(actions), "mus": as_bits(mus),
            "rl_games_obs": as_bits(o2), "rl_games_actions": as_bits(a2), "rl_games_mus": as_bits(m2),
            "rsl_rl_obs": as_bits(o3), "rsl_rl_actions": as_bits(a3), "actions_only": as_bits(n2)}


def stone_poses_view(num_envs=96, seed=43):
    from oracle import ref_rigid_collection as rc

    g = torch.Generator().manual_seed(seed)
    steps_pos = 5.0 * torch.randn(num_envs, 20, 3, generator=g)
    env_ids = torch.randperm(num_envs, generator=g)[:31].sort().values
    poses, view_ids = rc.reference_stone_pose_write(steps_pos, env_ids)
    poses_all, view_ids_all = rc.reference_stone_pose_write(steps_pos, torch.arange(num_envs))
    return {"steps_pos": steps_pos.numpy(), "env_ids": env_ids.numpy(), "view_ids": view_ids.numpy(),
            "rows": poses[view_ids].numpy(), "view_ids_all": view_ids_all.numpy(), "poses_all": poses_all.numpy()}


if __name__ == "__main__":
    assert ref_loader.reference_available(), "needs the reference checkout under /root/reference"
    torch.set_num_threads(1)
    np.savez_compressed(os.path.join(HERE, "allsteps_replay_n64.npz"), **replay())
    # a quiet replay: nothing falls, so most steps have NO reset and the reference runs a single pass (DRL:360)
    np.savez_compressed(os.path.join(HERE, "allsteps_replay_n16_quiet.npz"),
                        **replay(num_envs=16, steps=12, seed=5, high_index_from=-1, fall_fraction=0.0))
    np.savez_compressed(os.path.join(HERE, "stones_levels.npz"), **stones_levels())
    np.savez_compressed(os.path.join(HERE, "math_helpers.npz"), **math_helpers())
    np.savez_compressed(os.path.join(HERE, "mirror_symmetry.npz"), **mirror_symmetry())
    np.savez_compressed(os.path.join(HERE, "stone_poses_view.npz"), **stone_poses_view())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
