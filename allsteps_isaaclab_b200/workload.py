"""Synthetic workload that stays on the distribution of SURVEY.md section 8(d) for as long as it is stepped.

The physics engine is out of scope, so the step is timed on synthetic post-physics states.  Such a state only means
something relative to the MDP state it meets: the swing foot is scattered around the env's CURRENT stone, the contact
matrices have their entries in the CURRENT stone's column.  Rotating a handful of states that were all generated from
the initial MDP state lets the two drift apart after the first step (feet and contacts no longer sit where the env is
heading: almost no env advances any more, and the slide-window / re-gather / step-reward branches go quiet).

`ChainedWorkload` therefore builds a CYCLE of P states, state k generated from the MDP state that k steps of the cycle
lead to, and rewinds the MDP state (AllstepsMDP.restore of a snapshot taken at the cycle start, Philox position
included) every P steps.  Every step of every cycle then sees exactly the pairing it was generated for; the rewind is
a device-to-device copy of the state words and stone windows (80 B per env) issued on the stream like a step.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import synthetic as syn
from .config import AllstepsCfg
from .mdp import AllstepsMDP, PhysicsViews, StepBuffers

ISAAC_NUM_BODIES = 17
ISAAC_BODY_ROWS = (16, 13, 0)  # right_foot, left_foot, torso rows used for the Isaac-layout variant


def to_isaac_layout(d: Dict[str, torch.Tensor]) -> Tuple[Dict[str, torch.Tensor], Tuple[int, int, int]]:
    """The same state in the layout Isaac Lab hands out: root pos / quat / lin vel as slices of ONE (N,13)
    `root_state_w` tensor (articulation_data.py:366-380) and body positions as a slice of the (N,B,13) `body_state_w`
    tensor (articulation_data.py:430-449)."""
    N = d["root_pos_w"].shape[0]
    dev = d["root_pos_w"].device
    root_state = torch.zeros(N, 13, device=dev)
    root_state[:, 0:3], root_state[:, 3:7], root_state[:, 7:10] = d["root_pos_w"], d["root_quat_w"], d["root_lin_vel_w"]
    body_state = torch.zeros(N, ISAAC_NUM_BODIES, 13, device=dev)
    for k, r in enumerate(ISAAC_BODY_ROWS):
        body_state[:, r, 0:3] = d["body_pos_w"][:, k]
    out = dict(d)
    out["root_pos_w"], out["root_quat_w"], out["root_lin_vel_w"] = (root_state[:, 0:3], root_state[:, 3:7],
                                                                    root_state[:, 7:10])
    out["body_pos_w"] = body_state[..., 0:3]
    out["_keep"] = (root_state, body_state)
    return out, ISAAC_BODY_ROWS


class ChainedWorkload:
    def __init__(self, mdp: AllstepsMDP, origins: torch.Tensor, out: StepBuffers, period: int, seed: int,
                 cfg: Optional[AllstepsCfg] = None, fall_fraction: float = 0.02, layout: str = "dense",
                 stones_change: bool = False, close_step=None):
        """`close_step(mdp)`: how a step is closed when the builder advances the chain (default: mdp.step closes it
        itself); sharded callers with a global promotion route pass their own so that the chain sees the same
        promotions the timed steps will."""
        self.mdp, self.out, self.period = mdp, out, int(period)
        self.cfg = cfg or mdp.cfg
        self.stones_change = bool(stones_change)
        self.layout = layout
        dev = mdp.device
        gen = torch.Generator(device=dev).manual_seed(seed)
        self.snap = mdp.snapshot(include_stones=self.stones_change)
        self.sets: List[Tuple[PhysicsViews, Dict[str, torch.Tensor]]] = []
        adv, rst = [], []
        self._step_fn = close_step
        for _ in range(self.period):
            st = mdp.export_state(("steps_pos", "curr_target_index", "swing_leg"))
            d = syn.random_physics_state(self.cfg, st["steps_pos"], st["curr_target_index"], st["swing_leg"], gen,
                                         fall_fraction=fall_fraction)
            del st
            d.pop("root_ang_vel_w", None)
            rows = (0, 1, 2)
            if layout == "isaac":
                d, rows = to_isaac_layout(d)
            v = PhysicsViews.from_dict(d, origins, rows)
            self.sets.append((v, d))
            self._run(v, d)
            s = mdp.read_stats()
            adv.append(s["n_advanced"])
            rst.append(s["n_reset"])
        n = float(mdp.num_envs)
        self.advance_rate = sum(adv) / (n * self.period)
        self.reset_rate = sum(rst) / (n * self.period)
        self.set_bytes = sum(t.numel() * t.element_size() for k, t in self.sets[0][1].items() if torch.is_tensor(t))
        if layout == "isaac":
            self.set_bytes = sum(t.numel() * t.element_size() for t in self.sets[0][1]["_keep"]) + sum(
                self.sets[0][1][k].numel() * 4 for k in ("joint_pos", "joint_vel", "actions", "force_matrix_right",
                                                          "force_matrix_left"))
        self.rewind_bytes = int(self.snap.numel())
        mdp.restore(self.snap, include_stones=self.stones_change)
        self.j = 0

    def _run(self, v, d):
        if self._step_fn is not None:
            self._step_fn(self.mdp, v, d, self.out)
        else:
            self.mdp.step(v, d["actions"], self.out)

    def rewind(self):
        self.mdp.restore(self.snap, include_stones=self.stones_change)
        self.j = 0

    def step(self):
        """One step of the cycle (rewinding first when a cycle starts)."""
        k = self.j % self.period
        if k == 0 and self.j > 0:
            self.mdp.restore(self.snap, include_stones=self.stones_change)
        v, d = self.sets[k]
        self._run(v, d)
        self.j += 1

    def current(self):
        return self.sets[self.j % self.period]
