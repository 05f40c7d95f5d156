"""Helpers to replay the committed golden fixtures (outputs of the unmodified reference, tests/golden/)."""
from __future__ import annotations

import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PHYS_KEYS = ["root_pos_w", "root_quat_w", "root_lin_vel_w", "root_ang_vel_w", "body_pos_w", "joint_pos", "joint_vel",
             "force_matrix_right", "force_matrix_left", "actions"]
STATE_KEYS = ["curr_target_index", "prev_target_index", "next_target_index", "swing_leg", "target_reach_count",
              "episode_length_buf", "curriculum", "potentials"]
REPLAYS = ["allsteps_replay_n64.npz", "allsteps_replay_n16_quiet.npz"]


def load(name: str):
    return np.load(os.path.join(GOLDEN_DIR, name))


def t(a) -> torch.Tensor:
    return torch.from_numpy(np.asarray(a).copy())


def initial_state(d):
    return {k: t(d[f"init_{k}"]) for k in ("curr_target_index", "swing_leg", "target_reach_count",
                                           "episode_length_buf", "curriculum", "potentials")}


def step_inputs(d, step: int):
    return {k: t(d[f"s{step}_in_{k}"]) for k in PHYS_KEYS}
