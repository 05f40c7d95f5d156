"""Shared replay driver for the parity tests (test infrastructure).

A *scenario* is a seeded sequence of K synthetic post-physics states, each conditioned on the MDP state the
oracle holds at that moment (so feet land near the stone the env is currently heading for), plus the uniform
tables of that step taken from the Philox twin.  The same (phys, actions, tables) triple is handed to every
implementation under test.
"""
from __future__ import annotations

import numpy as np
import torch

from allsteps_isaaclab_b200 import synthetic as syn
from allsteps_isaaclab_b200.config import AllstepsCfg, BODY_NAMES
from oracle import philox


class Scenario:
    def __init__(self, num_envs: int, seed: int = 1234, full_bodies: bool = False, env_id_offset: int = 0,
                 fall_fraction: float = 0.02, cfg: AllstepsCfg | None = None):
        self.cfg = cfg or AllstepsCfg()
        self.N = num_envs
        self.seed = seed
        self.gen = torch.Generator().manual_seed(seed)
        self.env_id_offset = env_id_offset
        self.global_ids = np.arange(num_envs, dtype=np.int64) + env_id_offset
        self.fall_fraction = fall_fraction
        if full_bodies:
            self.num_bodies = len(BODY_NAMES)
            self.body_indices = self.cfg.body_indices()
        else:
            self.num_bodies = 3
            self.body_indices = (0, 1, 2)
        self.env_origins = syn.env_origins_grid(num_envs, self.cfg.env_spacing)
        self.joint_limits = syn.joint_limits_tensor(self.cfg)

    def stone_uniforms(self, step: int) -> torch.Tensor:
        return torch.from_numpy(philox.stone_tables(self.seed, step, self.global_ids, self.cfg.num_steps))

    def reset_uniforms(self, step: int):
        m, n = philox.reset_tables(self.seed, step, self.global_ids)
        return torch.from_numpy(m), torch.from_numpy(n)

    def initial_mdp_state(self, per_env_levels: bool = False):
        return syn.random_mdp_state(self.cfg, self.N, self.gen, per_env_levels=per_env_levels)

    def physics(self, stones, curr_target_index, swing_leg):
        return syn.random_physics_state(self.cfg, stones, curr_target_index, swing_leg, self.gen,
                                        num_bodies=self.num_bodies, body_indices=self.body_indices,
                                        fall_fraction=self.fall_fraction)


def install_mdp_state(target, state: dict, num_steps: int = 20):
    """Write a `random_mdp_state` dict into an oracle / hosted-reference object (reference attribute names)."""
    for name in ("curr_target_index", "swing_leg", "target_reach_count", "episode_length_buf", "curriculum"):
        getattr(target, name)[:] = state[name]
    target.potentials = state["potentials"].clone()
    target.old_potentials = state["potentials"].clone()
    target.prev_target_index = torch.clamp(target.curr_target_index - 1, 0, num_steps - 1)
    target.next_target_index = torch.clamp(target.curr_target_index + 1, 0, num_steps - 1)
