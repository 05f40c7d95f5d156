"""Constants of the Allsteps-v0 MDP step, gathered into one plain dataclass.

Every field cites where the reference defines it.  `ENV` = source/isaaclab_tasks/isaaclab_tasks/direct/
allsteps/allsteps_env.py, `CFG` = .../allsteps_env_cfg.py, `XML` = source/isaaclab_assets/data/mjcf/walker3d.xml,
`WALKER` = source/isaaclab_assets/isaaclab_assets/robots/walker3d.py (all under /root/reference).

The reference hard-codes half of these in `AllstepsEnv.__init__` (ENV:41-59) and keeps the rest in the
config class (CFG:52-235); the kernels receive all of them as one POD struct (`AsParams`, include/allsteps_b200.h).
"""
from __future__ import annotations

import dataclasses
import math
from typing import List, Sequence, Tuple

NUM_JOINTS = 21  # CFG:57 action_space
NUM_STONES = 20  # CFG:90 num_steps
OBS_DIM = 59  # CFG:58 observation_space
NUM_LEVELS = 10  # ENV:45 max_curriculum = 9 -> levels 0..9

RIGHT_FOOT = 0  # ENV:29
LEFT_FOOT = 1  # ENV:30

# PhysX (BFS) joint order as documented by the gear table comments, CFG:133-155.
JOINT_NAMES: Tuple[str, ...] = (
    "abdomen_z", "abdomen_y",
    "right_shoulder_x", "right_shoulder_y", "right_shoulder_z",
    "left_shoulder_x", "left_shoulder_y", "left_shoulder_z",
    "abdomen_x", "right_elbow", "left_elbow",
    "right_hip_x", "right_hip_y", "right_hip_z",
    "left_hip_x", "left_hip_y", "left_hip_z",
    "right_knee", "left_knee", "right_ankle", "left_ankle",
)

# Joint ranges in degrees, XML:39-98 (`range="lo hi"` of each hinge).
_JOINT_RANGE_DEG = {
    "abdomen_z": (-35, 35), "abdomen_y": (-80, 15), "abdomen_x": (-25, 25),
    "right_hip_x": (-25, 5), "right_hip_z": (-40, 35), "right_hip_y": (-100, 20),
    "right_knee": (-150, 0), "right_ankle": (-20, 40),
    "left_hip_x": (-25, 5), "left_hip_z": (-40, 35), "left_hip_y": (-100, 20),
    "left_knee": (-150, 0), "left_ankle": (-20, 40),
    "right_shoulder_x": (-60, 100), "right_shoulder_z": (-35, 120), "right_shoulder_y": (-60, 60),
    "right_elbow": (0, 120),
    "left_shoulder_x": (-60, 100), "left_shoulder_z": (-35, 120), "left_shoulder_y": (-60, 60),
    "left_elbow": (0, 120),
}

# Body names in MJCF document order (XML:28-101); only three are read by the task (CFG:214-215).
BODY_NAMES: Tuple[str, ...] = (
    "walker3d", "head", "torso", "waist", "pelvis",
    "right_thigh", "right_shin", "right_foot", "left_thigh", "left_shin", "left_foot",
    "right_upper_arm", "right_lower_arm", "right_hand", "left_upper_arm", "left_lower_arm", "left_hand",
)


def _names_to_indices(names: Sequence[str]) -> Tuple[int, ...]:
    return tuple(JOINT_NAMES.index(n) for n in names)


@dataclasses.dataclass
class AllstepsCfg:
    # --- env timing (CFG:54-55,62; DRL:248-250) ---------------------------------------------------------
    episode_length_s: float = 15.0
    decimation: int = 4
    sim_dt: float = 1.0 / 240.0
    # --- stones (CFG:90,97; ENV:41-44,50) ---------------------------------------------------------------
    num_steps: int = NUM_STONES
    step_radius: float = 0.25
    dist_range: Tuple[float, float] = (0.75, 0.9)
    pitch_range_deg: Tuple[float, float] = (-30.0, 30.0)
    yaw_range_deg: Tuple[float, float] = (-20.0, 20.0)
    tilt_range_deg: Tuple[float, float] = (-15.0, 15.0)
    init_step_separation: float = 0.75
    # --- curriculum (ENV:45-48,53) ----------------------------------------------------------------------
    max_curriculum: int = 9
    termination_height_range: Tuple[float, float] = (0.75, 0.45)  # linspace over the 10 levels, ENV:46
    applied_gain_range: Tuple[float, float] = (1.2, 1.2)  # ENV:47
    curriculum_progress_threshold: float = 12.0  # ENV:53 (sic: curriculum_progess_theshold)
    # --- foot state machine (ENV:32,56) -----------------------------------------------------------------
    contact_epsilon: float = 1e-4
    stop_frames: int = 2
    # --- rewards (CFG:222-228, ENV:356-375) -------------------------------------------------------------
    energy_cost_scale: float = 0.009
    actions_cost_scale: float = 0.01
    alive_reward_scale: float = 2.0
    dof_vel_scale: float = 0.1
    joint_at_limit_cost_scale: float = 0.1
    death_cost: float = -1.0
    # --- terminations (CFG:228, ENV:402) ----------------------------------------------------------------
    termination_height_absolute: float = 0.4
    max_root_speed: float = 5.0
    # extension (no reference counterpart): missed-step termination, used when AS_FLAG_MISSED_STEP is set
    missed_step_height: float = 0.05
    # --- reset (CFG:232-233, ENV:505-511, WALKER:36-39) -------------------------------------------------
    initial_joint_angle_range: Tuple[float, float] = (-0.1, 0.1)
    initial_joint_angle_clip_range: Tuple[float, float] = (-0.95, 0.95)
    default_root_pos: Tuple[float, float, float] = (0.2, 0.0, 1.5)
    env_spacing: float = 4.0  # CFG:78
    # --- robot tables -----------------------------------------------------------------------------------
    joint_gears: Tuple[float, ...] = (
        60, 80, 60, 50, 60, 60, 50, 60, 60, 60, 60, 80, 100, 60, 80, 100, 60, 90, 90, 60, 60
    )  # CFG:133-155
    right_body_names: Tuple[str, ...] = (
        "right_shoulder_x", "right_shoulder_y", "right_shoulder_z", "right_elbow",
        "right_hip_x", "right_hip_y", "right_hip_z", "right_knee", "right_ankle",
    )  # CFG:217
    left_body_names: Tuple[str, ...] = (
        "left_shoulder_x", "left_shoulder_y", "left_shoulder_z", "left_elbow",
        "left_hip_x", "left_hip_y", "left_hip_z", "left_knee", "left_ankle",
    )  # CFG:218
    negation_body_names: Tuple[str, ...] = ("abdomen_z", "abdomen_x")  # CFG:219
    foot_names: Tuple[str, str] = ("right_foot", "left_foot")  # CFG:215
    torso_name: str = "torso"  # CFG:214

    # ---- derived -------------------------------------------------------------------------------------
    @property
    def step_dt(self) -> float:
        return self.sim_dt * self.decimation  # DRL: step_dt = sim.dt * decimation

    @property
    def max_episode_length(self) -> int:
        return math.ceil(self.episode_length_s / (self.sim_dt * self.decimation))  # DRL:248-250 -> 900

    @property
    def right_joint_indices(self) -> Tuple[int, ...]:
        return _names_to_indices(self.right_body_names)

    @property
    def left_joint_indices(self) -> Tuple[int, ...]:
        return _names_to_indices(self.left_body_names)

    @property
    def negation_joint_indices(self) -> Tuple[int, ...]:
        return _names_to_indices(self.negation_body_names)

    def joint_limits_rad(self) -> List[Tuple[float, float]]:
        """(lower, upper) per joint in radians, PhysX order. Converted in double, stored as fp32 by callers."""
        return [tuple(math.radians(v) for v in _JOINT_RANGE_DEG[n]) for n in JOINT_NAMES]

    def reset_joint_pose(self) -> List[float]:
        """Running-start pose written over the all-zero default joint positions, ENV:505-511."""
        q = [0.0] * NUM_JOINTS
        q[12] = q[17] = -math.pi / 8
        q[15] = math.pi / 10
        q[2] = q[5] = math.pi / 3
        q[4] = -math.pi / 6
        q[7] = math.pi / 6
        q[9] = q[10] = math.pi / 3
        return q

    def mirror_permutation(self) -> Tuple[List[int], List[float]]:
        """(source index, sign) per joint for the left/right mirror, ENV:522-526:
        right[i] <- left[i], left[i] <- right[i], negation joints *= -1."""
        src = list(range(NUM_JOINTS))
        sign = [1.0] * NUM_JOINTS
        for r, l in zip(self.right_joint_indices, self.left_joint_indices):
            src[r], src[l] = l, r
        for n in self.negation_joint_indices:
            sign[n] = -1.0
        return src, sign

    def body_indices(self) -> Tuple[int, int, int]:
        """(right_foot, left_foot, torso) rows of `body_pos_w`, ENV:87-88."""
        return (BODY_NAMES.index(self.foot_names[0]), BODY_NAMES.index(self.foot_names[1]),
                BODY_NAMES.index(self.torso_name))
