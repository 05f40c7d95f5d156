"""TEST INFRASTRUCTURE ONLY -- CPU torch restatement ("port") of the reference Allsteps-v0 MDP step.

This file is the checker for the CUDA path and the timed CPU baseline; the product never imports it.
It restates, op for op in fp32 torch, what the reference computes, so that on CPU it is bit-identical to the
reference's own methods (pinned by tests/test_oracle_vs_reference.py where /root/reference is mounted, and by
the golden vectors in tests/golden/ that were produced by executing the reference itself).

Citations: ENV = source/isaaclab_tasks/isaaclab_tasks/direct/allsteps/allsteps_env.py,
MATH = source/isaaclab/isaaclab/utils/math.py, DRL = source/isaaclab/isaaclab/envs/direct_rl_env.py,
ART = source/isaaclab/isaaclab/assets/articulation/articulation.py (all under /root/reference).

Unlike the reference, random numbers are explicit inputs (`uniforms`), state lives in one small class, and the
Isaac Sim objects are replaced by a dict of tensors in the reference's layouts:
    root_pos_w (N,3)  root_quat_w (N,4 wxyz)  root_lin_vel_w (N,3)  root_ang_vel_w (N,3)  body_pos_w (N,B,3)
    joint_pos (N,J)  joint_vel (N,J)  force_matrix_left/right (N,1,S,3)
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

TWO_PI = 2 * math.pi
RIGHT, LEFT = 0, 1


# --------------------------------------------------------------------------------------------- math helpers
def scale_to_unit(x, lower, upper):
    """MATH:22-40 scale_transform."""
    offset = (lower + upper) * 0.5
    return 2 * (x - offset) / (upper - lower)


def unscale_from_unit(x, lower, upper):
    """MATH:43-61 unscale_transform."""
    offset = (lower + upper) * 0.5
    return x * (upper - lower) * 0.5 + offset


def euler_xyz_wrapped(quat):
    """MATH:413-444 euler_xyz_from_quat: angles end up in [0, 2*pi)."""
    w, x, y, z = quat[:, 0], quat[:, 1], quat[:, 2], quat[:, 3]
    sin_roll = 2.0 * (w * x + y * z)
    cos_roll = 1 - 2 * (x * x + y * y)
    roll = torch.atan2(sin_roll, cos_roll)
    sin_pitch = 2.0 * (w * y - z * x)
    # MATH:120-135 copysign(): |pi/2| * sign(sin_pitch)  (sign(0) = 0)
    half_pi = torch.abs(torch.full_like(sin_pitch, math.pi / 2.0)) * torch.sign(sin_pitch)
    pitch = torch.where(torch.abs(sin_pitch) >= 1, half_pi, torch.asin(sin_pitch))
    sin_yaw = 2.0 * (w * z + x * y)
    cos_yaw = 1 - 2 * (y * y + z * z)
    yaw = torch.atan2(sin_yaw, cos_yaw)
    return roll % TWO_PI, pitch % TWO_PI, yaw % TWO_PI


def rotate_by_inverse(q, v):
    """MATH:605-625 quat_rotate_inverse (2-D branch, degenerate bmm as dot product)."""
    q_w = q[..., 0]
    q_vec = q[..., 1:]
    a = v * (2.0 * q_w**2 - 1.0).unsqueeze(-1)
    b = torch.cross(q_vec, v, dim=-1) * q_w.unsqueeze(-1) * 2.0
    dot = torch.bmm(q_vec.view(q.shape[0], 1, 3), v.view(q.shape[0], 3, 1)).squeeze(-1)
    c = q_vec * dot * 2.0
    return a - b + c


def point_in_frame(frame_pos, frame_quat, point):
    """MATH:785-817 subtract_frame_transforms(t01, q01, t02)[0] with MATH:238-248 quat_inv, :81-92 normalize,
    :223-235 quat_conjugate and :545-564 quat_apply."""
    conj = torch.cat((frame_quat[:, 0:1], -frame_quat[:, 1:]), dim=-1)
    inv = conj / conj.norm(p=2, dim=-1).clamp(min=1e-9, max=None).unsqueeze(-1)
    vec = point - frame_pos
    xyz = inv[:, 1:]
    t = xyz.cross(vec, dim=-1) * 2
    return vec + inv[:, 0:1] * t + xyz.cross(t, dim=-1)


# --------------------------------------------------------------------------------------------- stones
def generate_stones(cfg, level: torch.Tensor, uniforms: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """ENV:125-174 `_generate_foot_steps_allsteps` in the env-local frame (origins are added by the caller,
    ENV:111).  `level` (N,) int64 already clamped to max_curriculum (ENV:126); `uniforms` (>=3,N,S) are the
    draws for dr, dphi, dtheta (the x/y tilt draws ENV:140-141 are sampled by the reference but never used).
    Returns (pos (N,S,3), cumulative yaw (N,S))."""
    N = level.shape[0]
    S = cfg.num_steps
    max_level = torch.tensor(cfg.max_curriculum, dtype=torch.int64)
    dist_lohi = torch.tensor(cfg.dist_range, dtype=torch.float32)
    yaw_lohi = torch.tensor(cfg.yaw_range_deg, dtype=torch.float32)
    pitch_lohi = torch.tensor(cfg.pitch_range_deg, dtype=torch.float32)
    level = torch.minimum(level, max_level)
    ratio = level / max_level
    dist_upper = torch.linspace(*dist_lohi, cfg.max_curriculum + 1)
    dist_range = torch.stack([dist_lohi[0].repeat(N), dist_upper[level]], dim=-1)
    yaw_range = torch.deg2rad(yaw_lohi.unsqueeze(0) * ratio.unsqueeze(1))
    pitch_range = torch.deg2rad(pitch_lohi.unsqueeze(0) * ratio.unsqueeze(1)) + torch.pi / 2

    dr = torch.lerp(dist_range[:, 0].unsqueeze(1), dist_range[:, 1].unsqueeze(1), uniforms[0])
    dphi = torch.lerp(yaw_range[:, 0].unsqueeze(1), yaw_range[:, 1].unsqueeze(1), uniforms[1])
    dtheta = torch.lerp(pitch_range[:, 0].unsqueeze(1), pitch_range[:, 1].unsqueeze(1), uniforms[2])

    dr[:, 0] = 0.0
    dphi[:, 0] = 0.0
    dtheta[:, 0] = torch.pi / 2
    dr[:, 1:3] = cfg.init_step_separation
    dphi[:, 1:3] = 0.0
    dtheta[:, 1:3] = torch.pi / 2

    dphi = torch.cumsum(dphi, dim=1)
    dx = dr * torch.sin(dtheta) * torch.cos(dphi)
    dy = dr * torch.sin(dtheta) * torch.sin(dphi)
    dz = dr * torch.cos(dtheta)
    pos = torch.stack((torch.cumsum(dx, dim=1), torch.cumsum(dy, dim=1), torch.cumsum(dz, dim=1)), dim=2)
    return pos, dphi


# --------------------------------------------------------------------------------------------- the MDP
# --------------------------------------------------------------------------------------------- mirror symmetry
def symmetric_states(x, right_ids, left_ids, negate_ids, kind: str, num_joints: int = 21):
    """ENV:570-660 for one tensor: `kind` "obs" (N,59) or "actions" (N,21; also used for `mus`).  Returns
    vstack((x, mirrored(x))), or None for None."""
    if x is None:
        return None
    right_ids, left_ids, negate_ids = (torch.as_tensor(t, dtype=torch.int64) for t in (right_ids, left_ids, negate_ids))
    if kind == "obs":
        J = num_joints
        steps_neg = torch.tensor([3 * i + 1 for i in range(3)], dtype=torch.int64)  # y of prev / curr / next, ENV:620
        root_neg = torch.tensor([1, 4], dtype=torch.int64)  # roll, v_y, ENV:621
        right = torch.cat((right_ids + 6, right_ids + 6 + J, torch.tensor([6 + 2 * J])))  # ENV:623
        left = torch.cat((left_ids + 6, left_ids + 6 + J, torch.tensor([6 + 2 * J + 1])))  # ENV:624
        neg = torch.cat((root_neg, 6 + negate_ids, 6 + J + negate_ids, 6 + 2 * J + 2 + steps_neg))  # ENV:625
    else:
        right, left, neg = right_ids, left_ids, negate_ids
    m = x.clone()
    m[:, right] = x[:, left]
    m[:, left] = x[:, right]
    m[:, neg] = -x[:, neg]
    return torch.vstack((x, m))


class AllstepsOracle:
    """Reference-exact MDP state machine on CPU tensors (reference dtypes: int64 indices, bool masks)."""

    def __init__(self, cfg, num_envs: int, env_origins: torch.Tensor, joint_limits: torch.Tensor,
                 body_indices=(0, 1, 2), stone_uniforms: Optional[torch.Tensor] = None,
                 intended_regen: bool = False, grid=None, seed: int = 0,
                 missed_step_height: Optional[float] = None):
        self.cfg = cfg
        self.N = N = num_envs
        self.S = S = cfg.num_steps
        self.J = joint_limits.shape[0]
        self.env_origins = env_origins
        self.lower = joint_limits[:, 0].unsqueeze(0).repeat(N, 1)  # robot.data.joint_pos_limits[:, :, 0]
        self.upper = joint_limits[:, 1].unsqueeze(0).repeat(N, 1)
        self.right_foot_row, self.left_foot_row, self.torso_row = body_indices
        self.step_dt = cfg.sim_dt * cfg.decimation
        self.max_episode_length = cfg.max_episode_length
        self.intended_regen = intended_regen  # extension: regen mask taken BEFORE the index reset
        self.grid = grid  # extension: oracle.grid_curriculum.GridCurriculum (every reset env is re-binned + regenerated)
        self.seed = seed
        self.step_index = 0
        # EXTENSION, no reference counterpart (BASELINE north_star "missed-step termination") -- this
        # method IS its specification, parity unpinned in the reference.  None = off = reference behaviour.
        # An env also terminates when its swing foot has come down below the top of the stone it is heading for,
        # outside that stone's footprint; swing leg and current stone as they are BEFORE the pass updates them.
        self.missed_step_height = missed_step_height
        # ENV:45-48
        self.termination_curriculum = torch.linspace(*cfg.termination_height_range, cfg.max_curriculum + 1)
        self.applied_gain_curriculum = torch.linspace(*cfg.applied_gain_range, cfg.max_curriculum + 1)
        self.joint_gears = torch.tensor(cfg.joint_gears, dtype=torch.float32)
        self.curriculum = torch.zeros(N, dtype=torch.int64)
        # ENV:66-71 stones (generated once at level 0)
        self.steps_pos = torch.zeros(N, S, 3)
        self.steps_dphi = torch.zeros(N, S)
        if stone_uniforms is None:
            stone_uniforms = torch.zeros(3, N, S)
        self.regenerate_stones(torch.arange(N), stone_uniforms)
        # ENV:74-79, 95-96
        self.swing_leg = torch.zeros(N, dtype=torch.int64)
        self.curr_target_index = torch.ones(N, dtype=torch.int64)
        self.prev_target_index = torch.clamp(self.curr_target_index - 1, 0, S - 1)
        self.next_target_index = torch.clamp(self.curr_target_index + 1, 0, S - 1)
        self.target_reach_count = torch.zeros(N, dtype=torch.int64)
        self.foot_contact = torch.zeros(N, 2)
        self.targets_w = torch.zeros(N, 3, 3)
        self.targets_b = torch.zeros(N, 3, 3)
        self.potentials = torch.zeros(N)
        self.old_potentials = torch.zeros(N)
        # DRL:179-182
        self.episode_length_buf = torch.zeros(N, dtype=torch.int64)
        self.reset_terminated = torch.zeros(N, dtype=torch.bool)
        self.reset_time_outs = torch.zeros(N, dtype=torch.bool)
        self.actions = torch.zeros(N, self.J)
        self.reset_writes: Dict[str, torch.Tensor] = {}
        # mirror tables, ENV:90-92
        self.right_ids = torch.tensor(cfg.right_joint_indices, dtype=torch.int64)
        self.left_ids = torch.tensor(cfg.left_joint_indices, dtype=torch.int64)
        self.negate_ids = torch.tensor(cfg.negation_joint_indices, dtype=torch.int64)
        self.reset_pose = torch.tensor(cfg.reset_joint_pose(), dtype=torch.float64).to(torch.float32)

    # ------------------------------------------------------------------ stones, ENV:106-123
    def regenerate_stones(self, env_ids: torch.Tensor, uniforms: torch.Tensor):
        pos, dphi = generate_stones(self.cfg, self.curriculum, uniforms)
        pos = pos + self.env_origins.unsqueeze(1)
        self.steps_pos[env_ids] = pos[env_ids]
        self.steps_dphi[env_ids] = dphi[env_ids]

    # ------------------------------------------------------------------ physics views
    def load_physics(self, phys: Dict[str, torch.Tensor]):
        """Install one post-physics state; tensors are cloned because reset writes rows back (ART:316-489)."""
        self.phys = {k: v.clone() for k, v in phys.items() if k != "actions"}

    # ------------------------------------------------------------------ one pass, ENV:276-324
    def mdp_pass(self):
        p = self.phys
        rows = torch.arange(self.N)
        right_foot = p["body_pos_w"][:, self.right_foot_row]
        left_foot = p["body_pos_w"][:, self.left_foot_row]
        torso = p["body_pos_w"][:, self.torso_row]
        self.torso_to_feet_height = torso[:, 2] - torch.minimum(left_foot[:, 2], right_foot[:, 2])  # ENV:281-283
        self.roll, self.pitch, self.yaw = euler_xyz_wrapped(p["root_quat_w"])  # ENV:285
        self.joint_pos_scaled = scale_to_unit(p["joint_pos"], self.lower, self.upper)  # ENV:287-291
        self.root_vec_b = rotate_by_inverse(p["root_quat_w"], p["root_lin_vel_w"])  # ENV:293
        self.root_ang_vec_b = rotate_by_inverse(p["root_quat_w"], p["root_ang_vel_w"])  # ENV:295 (unused)

        # ---- foot state machine, ENV:418-457
        force_left = torch.linalg.vector_norm(p["force_matrix_left"], dim=-1).squeeze(dim=1)
        force_right = torch.linalg.vector_norm(p["force_matrix_right"], dim=-1).squeeze(dim=1)
        force = torch.stack((force_right, force_left), dim=-1)  # (N,S,2) right first
        pressed = force[rows, self.curr_target_index] > self.cfg.contact_epsilon
        self.foot_contact[:] = pressed.float()
        stone_xy = self.steps_pos[rows, self.curr_target_index, :2]
        feet_xy = torch.stack((right_foot[:, :2], left_foot[:, :2]), dim=1)  # body_pos_w[:, foot_indices, :2]
        self.foot_to_target_dist_xy = torch.linalg.vector_norm(feet_xy - stone_xy[:, None, :], dim=-1)
        if self.missed_step_height is not None:  # extension (see __init__): evaluated on the pre-update leg / stone
            feet_z = torch.stack((right_foot[:, 2], left_foot[:, 2]), dim=1)
            swing_z = feet_z[rows, self.swing_leg]
            stone_z = self.steps_pos[rows, self.curr_target_index, 2]
            swing_d = self.foot_to_target_dist_xy[rows, self.swing_leg]
            self.missed_step = (swing_z < stone_z + self.missed_step_height) & (swing_d >= self.cfg.step_radius)
        self.target_reached = (pressed[rows, self.swing_leg] > 0) & (
            (self.foot_to_target_dist_xy < self.cfg.step_radius)[rows, self.swing_leg])
        self.target_reach_count[self.target_reached] += 1
        advance = self.target_reach_count >= self.cfg.stop_frames
        self.swing_leg[advance] = self.swing_leg[advance] ^ 1
        last = self.S - 1
        self.curr_target_index[advance] = torch.clamp(self.curr_target_index[advance] + 1, 0, last)
        self.prev_target_index[advance] = torch.clamp(self.curr_target_index[advance] - 1, 0, last)
        self.next_target_index[advance] = torch.clamp(self.curr_target_index[advance] + 1, 0, last)
        self.target_reach_count[advance] = 0
        self.advanced = advance

        # ---- targets in world and root frame, ENV:459-467, 302-316
        self.targets_w[:] = torch.stack([self.steps_pos[rows, self.prev_target_index],
                                         self.steps_pos[rows, self.curr_target_index],
                                         self.steps_pos[rows, self.next_target_index]], dim=1)
        for k in range(3):
            self.targets_b[:, k] = point_in_frame(p["root_pos_w"], p["root_quat_w"], self.targets_w[:, k])

        # ---- potentials, ENV:407-416
        to_next = self.targets_w[:, -1] - p["root_pos_w"]
        self.body_dist_to_target_xy = torch.linalg.vector_norm(to_next[:, 0:2], dim=-1)
        self.old_potentials = self.potentials.clone()
        self.potentials = -(self.body_dist_to_target_xy) / self.step_dt
        # ENV:323-324: the camera follow does a device->host read of env 0 every pass
        _ = tuple(p["root_pos_w"][0].tolist())

    # ------------------------------------------------------------------ dones, ENV:396-405
    def dones(self):
        self.mdp_pass()
        p = self.phys
        time_out = self.episode_length_buf >= self.max_episode_length - 1
        fell = self.torso_to_feet_height < self.termination_curriculum[self.curriculum]
        so_fast = torch.linalg.vector_norm(p["root_lin_vel_w"], dim=-1) > self.cfg.max_root_speed
        died = p["root_pos_w"][:, 2] < self.cfg.termination_height_absolute
        self.fell, self.so_fast, self.died = fell, so_fast, died
        if self.missed_step_height is not None:
            self.missed_step_pass1 = self.missed_step.clone()
            return fell | so_fast | died | self.missed_step, time_out
        return fell | so_fast | died, time_out

    # ------------------------------------------------------------------ rewards, ENV:347-394
    def rewards(self):
        c = self.cfg
        p = self.phys
        alive = torch.ones_like(self.torso_to_feet_height) * c.alive_reward_scale
        progress = self.potentials - self.old_potentials
        roll_bad = (self.roll > 0.4) | (self.roll < -0.4)
        pitch_bad = (self.pitch > 0.4) | (self.pitch < -0.2)
        roll_cost = torch.where(roll_bad, self.roll.abs(), torch.zeros_like(self.roll))
        pitch_cost = torch.where(pitch_bad, self.pitch.abs(), torch.zeros_like(self.pitch))
        speed = torch.linalg.vector_norm(p["root_lin_vel_w"], dim=-1)
        speed_cost = torch.where(speed > 1.6, speed - 1.6, torch.zeros_like(speed))
        action_cost = c.actions_cost_scale * torch.linalg.vector_norm(self.actions, dim=-1)
        energy_cost = c.energy_cost_scale * torch.sum(torch.abs(p["joint_vel"] * self.actions), dim=-1)
        limit_cost = torch.count_nonzero(torch.abs(self.joint_pos_scaled) > 0.99, dim=-1).float() \
            * c.joint_at_limit_cost_scale
        rows = torch.arange(self.N)
        pays_step = self.target_reached & (self.target_reach_count == 1) & (self.curr_target_index < self.S - 1)
        dist = self.foot_to_target_dist_xy[rows, self.swing_leg]
        step_reward = torch.where(pays_step, 50 * torch.exp(-dist / 0.25), torch.zeros_like(pays_step))
        at_goal = (self.curr_target_index == self.S - 1) & (self.body_dist_to_target_xy < 0.15)
        bonus = torch.where(at_goal, 10 * torch.ones_like(at_goal), torch.zeros_like(at_goal))
        total = (alive + progress - roll_cost - pitch_cost - speed_cost - energy_cost - action_cost
                 - limit_cost + step_reward + bonus)
        return torch.where(self.reset_terminated, c.death_cost * torch.ones_like(total), total)

    # ------------------------------------------------------------------ reset, ENV:469-567 + DRL:563-584
    def reset_rows(self, env_ids: torch.Tensor, mirror_u: torch.Tensor, noise_u: torch.Tensor,
                   stone_uniforms: Optional[torch.Tensor] = None):
        """`mirror_u` (k,), `noise_u` (k,J): the draws of ENV:518 and ENV:542 for rows `env_ids`."""
        c = self.cfg
        p = self.phys
        S = self.S
        # promotion rule, ENV:471-479
        if self.curr_target_index.float().mean() > c.curriculum_progress_threshold:
            self.curriculum = torch.clamp(self.curriculum + 1, 0, c.max_curriculum)
        regen_mask_before = self.curr_target_index > S // 2
        index_at_end = self.curr_target_index[env_ids].clone()
        # scene.reset + base class, DRL:563-584, contact_sensor.py:155
        p["force_matrix_left"][env_ids] = 0.0
        p["force_matrix_right"][env_ids] = 0.0
        self.episode_length_buf[env_ids] = 0
        # ENV:487-494
        self.old_potentials[env_ids] = 0.0
        self.potentials[env_ids] = 0.0
        self.target_reach_count[env_ids] = 0
        self.swing_leg[env_ids] = 0
        self.curr_target_index[env_ids] = 1
        self.prev_target_index[env_ids] = torch.clamp(self.curr_target_index[env_ids] - 1, 0, S - 1)
        self.next_target_index[env_ids] = torch.clamp(self.curr_target_index[env_ids] + 1, 0, S - 1)
        # ENV:497-500: evaluated after the index reset => never true in the reference
        regen_mask = regen_mask_before if self.intended_regen else (self.curr_target_index > S // 2)
        replace_ids = env_ids[torch.isin(env_ids, regen_mask.nonzero(as_tuple=False).flatten())]
        if self.grid is not None:
            # grid-curriculum extension: outcome -> histogram, new bin by inverse CDF, stones at the bin's difficulty
            from . import grid_curriculum as gc

            replace_ids = env_ids
            self.grid.episode_end(env_ids.numpy(), index_at_end.numpy(), S, self.seed, self.step_index)
            pos, dphi = gc.generate_stones_for_bins(c, torch.from_numpy(self.grid.bins.copy()), self.grid.B,
                                                    stone_uniforms)
            pos = pos + self.env_origins.unsqueeze(1)
            self.steps_pos[env_ids] = pos[env_ids]
            self.steps_dphi[env_ids] = dphi[env_ids]
        elif len(replace_ids) > 0:
            self.regenerate_stones(replace_ids, stone_uniforms)
        self.regenerated_ids = replace_ids
        # running-start pose, ENV:505-515
        k = env_ids.shape[0]
        joint_pos = self.reset_pose.unsqueeze(0).repeat(k, 1)
        joint_vel = torch.zeros(k, self.J)
        root = torch.zeros(k, 13)
        root[:, 0:3] = torch.tensor(c.default_root_pos)
        root[:, 3] = 1.0
        root[:, :3] += self.env_origins[env_ids]
        # mirror, ENV:518-538
        flip = mirror_u > 0.5
        sub = torch.nonzero(flip, as_tuple=True)[0]
        for buf in (joint_pos, joint_vel):
            m = buf.clone()
            m[sub[:, None], self.right_ids] = buf[sub[:, None], self.left_ids]
            m[sub[:, None], self.left_ids] = buf[sub[:, None], self.right_ids]
            m[sub[:, None], self.negate_ids] *= -1
            buf[sub] = m[sub]
        root[sub, 4:7] *= -1
        flipped_envs = env_ids[flip]
        self.swing_leg[flipped_envs] = self.swing_leg[flipped_envs] ^ 1
        # noise + clip, ENV:542-560 (sample_uniform MATH:1313-1331)
        lo, hi = c.initial_joint_angle_range
        joint_pos[:] += noise_u * (hi - lo) + lo
        unit = scale_to_unit(joint_pos, self.lower[env_ids], self.upper[env_ids])
        unit = torch.clamp(unit, c.initial_joint_angle_clip_range[0], c.initial_joint_angle_clip_range[1])
        joint_pos = unscale_from_unit(unit, self.lower[env_ids], self.upper[env_ids])
        # the three PhysX writes, ENV:563-565; the data views change immediately (ART:316-341,400-420,472-489)
        self.reset_writes = {"env_ids": env_ids.clone(), "root_pose": root[:, :7].clone(),
                             "root_velocity": root[:, 7:].clone(), "joint_pos": joint_pos.clone(),
                             "joint_vel": joint_vel.clone()}
        p["root_pos_w"][env_ids] = root[:, 0:3]
        p["root_quat_w"][env_ids] = root[:, 3:7]
        p["root_lin_vel_w"][env_ids] = root[:, 7:10]
        p["root_ang_vel_w"][env_ids] = root[:, 10:13]
        p["joint_pos"][env_ids] = joint_pos
        p["joint_vel"][env_ids] = joint_vel
        self.mdp_pass()  # ENV:567 -- pass 2 over ALL envs

    # ------------------------------------------------------------------ observations, ENV:326-345
    def observations(self):
        c = self.cfg
        return torch.cat((
            self.torso_to_feet_height.unsqueeze(-1), self.roll.unsqueeze(-1), self.pitch.unsqueeze(-1),
            self.root_vec_b, self.joint_pos_scaled,
            torch.clamp(self.phys["joint_vel"] * c.dof_vel_scale, -5, 5),
            self.foot_contact, self.targets_b.reshape(self.N, -1)), dim=-1)

    # ------------------------------------------------------------------ action path, ENV:257-274
    def clamp_actions(self, actions):
        self.actions = torch.clamp(actions.clone(), -1.0, 1.0)

    def joint_efforts(self):
        return self.applied_gain_curriculum[self.curriculum].unsqueeze(-1) * self.joint_gears.unsqueeze(0) \
            * self.actions

    # ------------------------------------------------------------------ one env step, DRL:326,351-375
    def step(self, phys: Dict[str, torch.Tensor], actions: torch.Tensor,
             mirror_u: Optional[torch.Tensor] = None, noise_u: Optional[torch.Tensor] = None,
             stone_uniforms: Optional[torch.Tensor] = None):
        """`mirror_u` (N,), `noise_u` (N,J), `stone_uniforms` (>=3,N,S) are per-env tables; rows of the envs
        that reset are consumed."""
        self.load_physics(phys)
        self.clamp_actions(actions)
        self.episode_length_buf += 1
        self.reset_terminated[:], self.reset_time_outs[:] = self.dones()
        reset_buf = self.reset_terminated | self.reset_time_outs
        reward = self.rewards()
        self.pass1 = {"curr_target_index": self.curr_target_index.clone(),
                      "swing_leg": self.swing_leg.clone(),
                      "target_reach_count": self.target_reach_count.clone(),
                      "potentials": self.potentials.clone()}
        ids = reset_buf.nonzero(as_tuple=False).squeeze(-1)
        self.reset_writes = {}
        if len(ids) > 0:
            self.reset_rows(ids, mirror_u[ids], noise_u[ids], stone_uniforms)
        obs = self.observations()
        self.step_index += 1
        return obs, reward, self.reset_terminated.clone(), self.reset_time_outs.clone(), ids
