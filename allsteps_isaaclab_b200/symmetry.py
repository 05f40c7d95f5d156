"""Mirror-symmetry augmentation with the reference's own signatures.

Drop-ins for `get_symmetric_states_rsl_rl` (ENV:570-609) and `get_symmetric_states_rl_games` (ENV:611-660) of the
reference's `allsteps_env.py`, and for the body of `A2CAgentSymmetry.play_steps` (learning/a2c_ppo_mirroring.py:20-40):
the clone + three fancy-index permutations + negation + `vstack` per tensor become one copy-and-permute kernel
(`as_mirror_rows`), and the three tensors rl_games hands over (obses, actions, mus) travel in ONE launch
(`as_mirror_batch`).  Results are bit-identical to the reference functions (tests/golden/mirror_symmetry.npz holds
outputs of the reference itself).

The permutation tables live in the CUDA library (built from AsParams.mirror_src / mirror_sign, i.e. CFG:217-219); the
index tensors `env.unwrapped.right_body_indices / left_body_indices / negation_body_indices` the reference functions
read are checked against them once per env, so an env with other tables is refused instead of mirrored wrongly.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _cabi
from .config import AllstepsCfg, NUM_JOINTS, OBS_DIM
from .mdp import AllstepsMDP

_KEY = "_allsteps_b200_symmetry"


def _unwrapped(env):
    return getattr(env, "unwrapped", env)


def _mdp_for(env) -> AllstepsMDP:
    """The handle whose mirror tables serve `env`: the env's own (`AllstepsHooksB200.mdp`), or a one-env handle created
    on the env's device the first time (the tables are constants of the task, not of the batch)."""
    base = _unwrapped(env)
    cached = getattr(base, _KEY, None)
    if cached is not None:
        return cached
    mdp = getattr(base, "mdp", None)
    if not isinstance(mdp, AllstepsMDP):
        dev = torch.device(getattr(env, "device", None) or getattr(base, "device"))
        mdp = AllstepsMDP(1, device=dev)
    cfg: AllstepsCfg = mdp.cfg
    for name, want in (("right_body_indices", cfg.right_joint_indices), ("left_body_indices", cfg.left_joint_indices),
                       ("negation_body_indices", cfg.negation_joint_indices)):
        have = getattr(base, name, None)
        if have is not None and [int(x) for x in torch.as_tensor(have).tolist()] != list(want):
            raise ValueError(f"env.unwrapped.{name} = {torch.as_tensor(have).tolist()} differs from the mirror tables the "
                             f"Allsteps kernels are built with ({list(want)}, CFG:217-219)")
    for space, dim in (("observation_space", OBS_DIM), ("action_space", NUM_JOINTS)):
        sp = getattr(base, space, None)
        shape = getattr(sp, "shape", None)
        if shape is not None and len(shape) >= 2 and int(shape[1]) != dim:
            raise ValueError(f"env.unwrapped.{space}.shape[1] = {shape[1]}, the Allsteps kernels are built for {dim}")
    try:
        setattr(base, _KEY, mdp)
    except Exception:
        pass
    return mdp


def _check(t: Optional[torch.Tensor], dim: int, name: str, dev) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dim() != 2 or t.shape[1] != dim:
        raise ValueError(f"{name} must be (rows,{dim}), got {tuple(t.shape)}")
    if t.device != dev:
        raise ValueError(f"{name} is on {t.device}; the mirror kernel runs on {dev} (there is no CPU path)")
    return t.detach().to(torch.float32).contiguous()


def mirror_batch(mdp: AllstepsMDP, obs: Optional[torch.Tensor], *action_like: Optional[torch.Tensor]):
    """vstack((x, mirrored(x))) for the observations and any number (<= 3) of action-shaped tensors, one launch."""
    dev = mdp.device
    tensors = [(_check(obs, OBS_DIM, "obs", dev), 0)] + [(_check(a, NUM_JOINTS, "actions", dev), 1) for a in action_like]
    jobs, outs = [], []
    for t, kind in tensors:
        if t is None:
            outs.append(None)
            continue
        o = torch.empty(2 * t.shape[0], t.shape[1], dtype=torch.float32, device=dev)
        outs.append(o)
        jobs.append(_cabi.AsMirrorJob(t.data_ptr(), o.data_ptr(), t.shape[0], kind, 0))
    if jobs:
        arr = (_cabi.AsMirrorJob * len(jobs))(*jobs)
        _cabi.check(mdp.lib.as_mirror_batch(mdp.handle, arr, len(jobs), mdp._stream()), "as_mirror_batch")
        mdp._keepalive_mirror = (tensors, outs)
    return outs


def get_symmetric_states_rl_games(obs: Optional[torch.Tensor], actions: Optional[torch.Tensor], env, is_critic: bool,
                                  mus: Optional[torch.Tensor]) -> Tuple[Optional[torch.Tensor], ...]:
    """ENV:611-660, same arguments and return value: (vstack(obs, mirrored), vstack(actions, mirrored),
    vstack(mus, mirrored)), `None` in -> `None` out.  `is_critic` is unused, as in the reference."""
    o, a, m = mirror_batch(_mdp_for(env), obs, actions, mus)
    return o, a, m


def get_symmetric_states_rsl_rl(obs: Optional[torch.Tensor], actions: Optional[torch.Tensor], env,
                                is_critic: bool = False) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """ENV:570-609, same arguments and return value."""
    o, a = mirror_batch(_mdp_for(env), obs, actions)
    return o, a


def augment_play_steps_batch(normal_batch: Dict[str, torch.Tensor], env) -> Dict[str, torch.Tensor]:
    """Body of `A2CAgentSymmetry.play_steps` with `symmetry: True` (learning/a2c_ppo_mirroring.py:23-38): the batch of
    `A2CAgent.play_steps()` doubled -- returns / dones / values / sigmas / neglogpacs repeated, obses / actions / mus
    mirrored (one kernel launch for the three).  Modifies and returns `normal_batch` like the reference."""
    normal_batch["returns"] = normal_batch["returns"].repeat(2, 1)
    normal_batch["dones"] = normal_batch["dones"].repeat(2)
    normal_batch["values"] = normal_batch["values"].repeat(2, 1)
    normal_batch["sigmas"] = normal_batch["sigmas"].repeat(2, 1)
    normal_batch["neglogpacs"] = normal_batch["neglogpacs"].repeat(2)
    new_obs, new_actions, new_mus = get_symmetric_states_rl_games(normal_batch["obses"], normal_batch["actions"], env,
                                                                  False, normal_batch["mus"])
    normal_batch["obses"] = new_obs
    normal_batch["actions"] = new_actions
    normal_batch["mus"] = new_mus
    return normal_batch
