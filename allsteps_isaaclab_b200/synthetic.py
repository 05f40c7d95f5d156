"""Seeded synthetic stand-in for the PhysX side of the Allsteps-v0 step.

The physics engine is out of scope; the MDP step sees its results only as tensors.  These generators
produce tensors of exactly the shapes/layouts the reference reads from `robot.data.*`,
`sensor_{left,right}.data.force_matrix_w` and `scene.env_origins`, with distributions chosen so that
every branch of the step is exercised (contacts on/off, feet inside/outside the stone radius, both signs
of roll/pitch, joint positions beyond the limits, actions beyond the clamp, all three fall causes).

Everything is generated with an explicit `torch.Generator` on the device of `stones`, so a
(seed, step) pair always reproduces the same state on CPU; CUDA callers generate on CPU and copy, or use a
CUDA generator when only throughput matters (bench.py).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from .config import AllstepsCfg, NUM_JOINTS, NUM_STONES


def env_origins_grid(num_envs: int, env_spacing: float, device="cpu") -> torch.Tensor:
    """Grid of env origins, the formula of terrain_importer.py:349-362 (rows x cols, centred)."""
    num_rows = np.ceil(num_envs / int(np.sqrt(num_envs)))
    num_cols = np.ceil(num_envs / num_rows)
    ii, jj = torch.meshgrid(torch.arange(num_rows), torch.arange(num_cols), indexing="ij")
    origins = torch.zeros(num_envs, 3)
    origins[:, 0] = -(ii.flatten()[:num_envs] - (num_rows - 1) / 2) * env_spacing
    origins[:, 1] = (jj.flatten()[:num_envs] - (num_cols - 1) / 2) * env_spacing
    return origins.to(device)


def joint_limits_tensor(cfg: AllstepsCfg, device="cpu") -> torch.Tensor:
    """(J,2) fp32 [lower, upper] in radians."""
    return torch.tensor(cfg.joint_limits_rad(), dtype=torch.float64).to(torch.float32).to(device)


def random_mdp_state(cfg: AllstepsCfg, num_envs: int, gen: torch.Generator, device="cpu",
                     per_env_levels: bool = False) -> Dict[str, torch.Tensor]:
    """Mid-episode MDP buffers: idx ~ U{1..19}, leg ~ U{0,1}, count ~ U{0,1}, ep_len ~ U{0..899}."""
    S = cfg.num_steps
    r = lambda lo, hi: torch.randint(lo, hi, (num_envs,), generator=gen, device=device, dtype=torch.int64)  # noqa
    curr = r(1, S)
    out = {
        "curr_target_index": curr,
        "swing_leg": r(0, 2),
        "target_reach_count": r(0, cfg.stop_frames),
        "episode_length_buf": r(0, cfg.max_episode_length),
        "potentials": -torch.rand(num_envs, generator=gen, device=device) * 60.0,
        "curriculum": r(0, cfg.max_curriculum + 1) if per_env_levels
        else torch.zeros(num_envs, dtype=torch.int64, device=device),
    }
    return out


def random_physics_state(
    cfg: AllstepsCfg,
    stones: torch.Tensor,  # (N,S,3) world frame
    curr_target_index: torch.Tensor,  # (N,) int64
    swing_leg: torch.Tensor,  # (N,) int64
    gen: torch.Generator,
    num_bodies: int = 3,
    body_indices=(0, 1, 2),  # rows (right_foot, left_foot, torso) inside body_pos_w
    fall_fraction: float = 0.02,
    fast_fraction: float = 0.001,
    contact_noise_fraction: float = 0.01,
) -> Dict[str, torch.Tensor]:
    """One synthetic post-physics state conditioned on the current stone / swing leg of each env."""
    N, S, _ = stones.shape
    dev = stones.device
    J = NUM_JOINTS
    randn = lambda *s: torch.randn(*s, generator=gen, device=dev)  # noqa: E731
    rand = lambda *s: torch.rand(*s, generator=gen, device=dev)  # noqa: E731
    ar = torch.arange(N, device=dev)
    curr = curr_target_index.clamp(0, S - 1)
    prev = (curr - 1).clamp(0, S - 1)
    stone_c = stones[ar, curr]
    stone_p = stones[ar, prev]

    root_pos = stone_p.clone()
    root_pos[:, :2] += 0.15 * randn(N, 2)
    root_pos[:, 2] = stone_p[:, 2] + 1.30 + 0.10 * randn(N)
    falling = rand(N) < fall_fraction
    root_pos[:, 2] = torch.where(falling, stone_p[:, 2] * 0 + 0.2 + 0.19 * rand(N), root_pos[:, 2])

    quat = torch.zeros(N, 4, device=dev)
    quat[:, 0] = 1.0
    quat = quat + 0.15 * randn(N, 4)
    quat = quat / quat.norm(dim=-1, keepdim=True)

    lin_vel = randn(N, 3)
    fast = rand(N) < fast_fraction
    lin_vel = torch.where(fast[:, None], lin_vel * 10.0, lin_vel)
    ang_vel = randn(N, 3)

    lim = joint_limits_tensor(cfg, dev)
    span = lim[:, 1] - lim[:, 0]
    joint_pos = (lim[:, 0] - 0.02 * span) + rand(N, J) * (1.04 * span)
    joint_vel = 3.0 * randn(N, J)
    actions = -1.2 + 2.4 * rand(N, J)

    # feet: swing foot scattered around the current stone, stance foot on the previous stone
    swing_xy = stone_c[:, :2] + 0.2 * randn(N, 2)
    swing = torch.cat([swing_xy, stone_c[:, 2:3] + 0.11], dim=-1)
    stance = torch.cat([stone_p[:, :2] + 0.03 * randn(N, 2), stone_p[:, 2:3] + 0.11], dim=-1)
    is_left_swing = (swing_leg == 1)[:, None]
    right_foot = torch.where(is_left_swing, stance, swing)
    left_foot = torch.where(is_left_swing, swing, stance)
    torso = root_pos + torch.tensor([0.0, 0.0, 0.25], device=dev)
    # a few envs get a crouched torso so that `fell` fires at every curriculum level
    crouch = rand(N) < fall_fraction
    torso[:, 2] = torch.where(crouch, torch.minimum(left_foot[:, 2], right_foot[:, 2]) + 0.1 + 0.3 * rand(N),
                              torso[:, 2])
    body_pos = randn(N, num_bodies, 3) + root_pos[:, None, :]
    body_pos[:, body_indices[0]] = right_foot
    body_pos[:, body_indices[1]] = left_foot
    body_pos[:, body_indices[2]] = torso

    # filtered contact matrices (N,1,S,3): the swing foot presses the current stone half of the time,
    # the stance foot presses the previous one, a sprinkle of spurious contacts elsewhere
    def contact_for(foot_is_swing: torch.Tensor) -> torch.Tensor:
        f = torch.zeros(N, S, 3, device=dev)
        col = torch.where(foot_is_swing, curr, prev)
        on = rand(N) < 0.5
        fz = torch.where(on, (200.0 * randn(N)).abs(), torch.zeros(N, device=dev))
        f[ar, col, 2] = fz
        f[ar, col, 0] = torch.where(on, 5.0 * randn(N), torch.zeros(N, device=dev))
        spurious = rand(N, S) < contact_noise_fraction
        f[:, :, 2] = torch.where(spurious, (50.0 * randn(N, S)).abs(), f[:, :, 2])
        return f.unsqueeze(1).contiguous()

    force_right = contact_for(swing_leg == 0)
    force_left = contact_for(swing_leg == 1)

    return {
        "root_pos_w": root_pos.contiguous(),
        "root_quat_w": quat.contiguous(),
        "root_lin_vel_w": lin_vel.contiguous(),
        "root_ang_vel_w": ang_vel.contiguous(),
        "body_pos_w": body_pos.contiguous(),
        "joint_pos": joint_pos.contiguous(),
        "joint_vel": joint_vel.contiguous(),
        "force_matrix_right": force_right,
        "force_matrix_left": force_left,
        "actions": actions.contiguous(),
    }


def straight_stones(cfg: AllstepsCfg, env_origins: torch.Tensor) -> torch.Tensor:
    """Level-0 stones: a straight flat line, 0.75 m apart (what ENV:125-174 yields at level 0)."""
    N = env_origins.shape[0]
    S = cfg.num_steps
    x = torch.zeros(S, device=env_origins.device)
    x[1:] = cfg.init_step_separation
    x = torch.cumsum(x, 0)
    stones = torch.zeros(N, S, 3, device=env_origins.device)
    stones[:, :, 0] = x
    return stones + env_origins[:, None, :]
