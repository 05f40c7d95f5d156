"""TEST INFRASTRUCTURE ONLY -- host the UNMODIFIED reference `AllstepsEnv` methods on synthetic tensors.

`make_reference_env(world)` builds an instance of the real reference class (loaded by `ref_loader`) whose
Isaac Sim facing attributes (`robot`, `sensor_left/right`, `scene`, `sim`, `steps`, `marker`) are small
tensor-holding fakes, then `step_mdp()` drives the reference hooks in the order of
source/isaaclab/isaaclab/envs/direct_rl_env.py:351-375.

Fakes reproduce the only side effects the MDP arithmetic can observe:
* `Articulation.write_root_pose_to_sim / write_root_velocity_to_sim / write_joint_state_to_sim` update the
  `robot.data` root/joint rows immediately (articulation.py:316-341, 400-420, 472-489); body positions are NOT
  refreshed until the next physics step, so they stay as they were ("stale") for pass 2.
* `scene.reset(env_ids)` zeroes the contact-matrix rows (contact_sensor.py:142-161) and the base class zeroes
  `episode_length_buf` rows (direct_rl_env.py:584) -- done by the base stand-in in `ref_loader`.

Uniform draws are injected: `torch.rand` is patched while reference code runs so the caller decides every
random number (per-env tables indexed by env id), which is what lets a counter-based Philox stream in the CUDA
kernels be compared with the reference's sequential `torch.rand` calls.
"""
from __future__ import annotations

import contextlib
import types
from typing import Dict, Optional

import torch

from . import ref_loader


class _Recorder:
    """Collects arguments of the PhysX write calls -- they are outputs of the path."""

    def __init__(self):
        self.calls = {}

    def put(self, name, *tensors):
        self.calls[name] = tuple(t.clone() if torch.is_tensor(t) else t for t in tensors)


class FakeRobot:
    def __init__(self, world: Dict[str, torch.Tensor], body_names, joint_names, default_root_pos):
        N = world["root_pos_w"].shape[0]
        J = world["joint_pos"].shape[1]
        dev = world["root_pos_w"].device
        self.rec = _Recorder()
        self.data = types.SimpleNamespace(
            body_names=list(body_names),
            joint_names=list(joint_names),
            root_pos_w=world["root_pos_w"],
            root_quat_w=world["root_quat_w"],
            root_lin_vel_w=world["root_lin_vel_w"],
            root_ang_vel_w=world["root_ang_vel_w"],
            body_pos_w=world["body_pos_w"],
            joint_pos=world["joint_pos"],
            joint_vel=world["joint_vel"],
            joint_pos_limits=world["joint_pos_limits"],
            default_joint_pos=torch.zeros(N, J, device=dev),
            default_joint_vel=torch.zeros(N, J, device=dev),
            default_root_state=torch.zeros(N, 13, device=dev),
        )
        self.data.default_root_state[:, 0:3] = torch.tensor(default_root_pos, device=dev)
        self.data.default_root_state[:, 3] = 1.0
        self._ALL_INDICES = torch.arange(N, dtype=torch.long, device=dev)

    def load_physics(self, world: Dict[str, torch.Tensor]):
        d = self.data
        for k in ("root_pos_w", "root_quat_w", "root_lin_vel_w", "root_ang_vel_w", "body_pos_w",
                  "joint_pos", "joint_vel"):
            setattr(d, k, world[k].clone())

    def reset(self, env_ids=None):
        pass

    def set_joint_effort_target(self, forces):
        self.rec.put("joint_effort_target", forces)

    def write_root_pose_to_sim(self, root_pose, env_ids=None):
        self.rec.put("root_pose", root_pose, env_ids)
        self.data.root_pos_w[env_ids] = root_pose[:, 0:3]
        self.data.root_quat_w[env_ids] = root_pose[:, 3:7]

    def write_root_velocity_to_sim(self, root_velocity, env_ids=None):
        self.rec.put("root_velocity", root_velocity, env_ids)
        self.data.root_lin_vel_w[env_ids] = root_velocity[:, 0:3]
        self.data.root_ang_vel_w[env_ids] = root_velocity[:, 3:6]

    def write_joint_state_to_sim(self, position, velocity, joint_ids=None, env_ids=None):
        self.rec.put("joint_state", position, velocity, env_ids)
        self.data.joint_pos[env_ids] = position
        self.data.joint_vel[env_ids] = velocity


class _Sensor:
    def __init__(self, force_matrix_w):
        self.data = types.SimpleNamespace(force_matrix_w=force_matrix_w)


class UniformTables:
    """Per-env uniform numbers handed to the patched `torch.rand`.

    mirror (N,), noise (N,J): rows are picked by the env ids being reset (ENV:518, ENV:542 via MATH:1313);
    stones (5,N,S): consumed in call order dr, dphi, dtheta, x_tilt, y_tilt (ENV:137-141).
    """

    def __init__(self, mirror=None, noise=None, stones=None):
        self.mirror, self.noise, self.stones = mirror, noise, stones
        self._stone_call = 0


@contextlib.contextmanager
def injected_uniforms(env, tables: Optional[UniformTables]):
    if tables is None:
        yield
        return
    real_rand = torch.rand
    tables._stone_call = 0
    N, S = env.num_envs, env.num_steps

    def fake_rand(*size, **kwargs):
        shape = tuple(size[0]) if len(size) == 1 and not isinstance(size[0], int) else tuple(size)
        ids = getattr(env, "_reset_ids_now", None)
        if shape == (N, S):  # stone generation, ENV:137-141 (S=20 never collides with J=21)
            out = tables.stones[tables._stone_call % 5]
            tables._stone_call += 1
            return out.clone()
        if ids is not None and shape == tuple(ids.shape):  # mirror coin, ENV:518
            return tables.mirror[ids].clone()
        if ids is not None and len(shape) == 2 and shape == (ids.shape[0], tables.noise.shape[1]):  # ENV:542
            return tables.noise[ids].clone()
        raise AssertionError(f"unexpected torch.rand{shape} inside reference code")

    torch.rand = fake_rand
    try:
        yield
    finally:
        torch.rand = real_rand


def make_reference_env(world: Dict[str, torch.Tensor], cfg_like, body_names, joint_names,
                       stone_uniforms: Optional[torch.Tensor] = None):
    """Instantiate the real reference class on synthetic tensors.

    `world` holds reference-layout tensors: root_pos_w, root_quat_w, root_lin_vel_w, root_ang_vel_w, body_pos_w,
    joint_pos, joint_vel, joint_pos_limits (N,J,2), force_matrix_left/right (N,1,S,3), env_origins (N,3).
    `cfg_like` supplies default_root_pos; everything else comes from the reference's own cfg class.
    `stone_uniforms` (5,N,S) feeds the one stone generation the reference performs in __init__ (ENV:71).
    """
    ref = ref_loader.load_reference()
    N = world["root_pos_w"].shape[0]
    dev = world["root_pos_w"].device

    class HostedAllstepsEnv(ref.AllstepsEnv):
        # properties the real DirectRLEnv base would provide (direct_rl_env.py:229-250)
        device = dev
        num_envs = N
        step_dt = (1.0 / 240.0) * 4
        max_episode_length = 900

        def _generate_foot_steps(self, env_ids=None):
            # The reference's version allocates an (N*N, S) temporary (ENV:112) which cannot be held
            # for N >= 65536.  Rows [0, N) of that temporary equal the un-repeated tensor, so for large N the
            # same scatter is done without the repeat; for small N the reference's own method runs untouched.
            if self.num_envs <= 4096:
                return super()._generate_foot_steps(env_ids)
            if env_ids is None:
                env_ids = torch.arange(self.num_envs, dtype=torch.long, device=self.device)
            pos, dphi, _ = self._generate_foot_steps_allsteps()
            pos = pos + self.scene.env_origins.unsqueeze(1)
            self.steps_pos[env_ids] = pos[env_ids]
            self.steps_dphi[env_ids] = dphi[env_ids]

        def _reset_idx(self, env_ids):
            self._reset_ids_now = env_ids
            try:
                super()._reset_idx(env_ids)
            finally:
                self._reset_ids_now = None

    env = HostedAllstepsEnv.__new__(HostedAllstepsEnv)
    # ---- what DirectRLEnv.__init__ + _setup_scene would have created (direct_rl_env.py:71-190, ENV:222-255)
    env.cfg = ref.AllstepsEnvCfg
    env.robot = FakeRobot(world, body_names, joint_names, cfg_like.default_root_pos)
    env.sensor_left = _Sensor(world["force_matrix_left"])
    env.sensor_right = _Sensor(world["force_matrix_right"])
    env.scene = types.SimpleNamespace(env_origins=world["env_origins"])
    env.sim = types.SimpleNamespace(set_camera_view=lambda **kw: None)
    env.steps = types.SimpleNamespace(write_object_pose_to_sim=lambda *a, **k: None)
    env.marker = types.SimpleNamespace(visualize=lambda **kw: None)
    env.episode_length_buf = torch.zeros(N, dtype=torch.long, device=dev)
    env.reset_terminated = torch.zeros(N, dtype=torch.bool, device=dev)
    env.reset_time_outs = torch.zeros(N, dtype=torch.bool, device=dev)
    env.reset_buf = torch.zeros(N, dtype=torch.bool, device=dev)
    env.actions = torch.zeros(N, 21, device=dev)
    env.extras = {}
    env._reset_ids_now = None

    # ---- the task's own __init__ body (ENV:40-102) runs unmodified; only `super().__init__` is skipped,
    # by temporarily giving the stand-in base an __init__ that accepts the arguments.
    tables = UniformTables(stones=stone_uniforms) if stone_uniforms is not None else None
    base = ref.BaseEnv
    had = "__init__" in base.__dict__
    old = base.__dict__.get("__init__")
    base.__init__ = lambda self, *a, **k: None
    try:
        with injected_uniforms(_ShapeOnly(N, 20), tables):
            ref.AllstepsEnv.__init__(env, ref.AllstepsEnvCfg)
    finally:
        if had:
            base.__init__ = old
        else:
            del base.__init__
    return env


class _ShapeOnly:
    def __init__(self, n, s):
        self.num_envs, self.num_steps = n, s


def load_physics(env, world: Dict[str, torch.Tensor]):
    """Install one synthetic post-physics state (what `scene.update` would have published)."""
    env.robot.load_physics(world)
    env.sensor_left.data.force_matrix_w = world["force_matrix_left"].clone()
    env.sensor_right.data.force_matrix_w = world["force_matrix_right"].clone()


def step_mdp(env, actions: torch.Tensor, tables: Optional[UniformTables] = None):
    """One MDP step in the order of direct_rl_env.py:326,351-375, running the reference's own hooks."""
    with injected_uniforms(env, tables):
        env._pre_physics_step(actions)  # DRL:326
        env.episode_length_buf += 1  # DRL:351
        env.reset_terminated[:], env.reset_time_outs[:] = env._get_dones()  # DRL:354
        env.reset_buf = env.reset_terminated | env.reset_time_outs  # DRL:355
        reward = env._get_rewards()  # DRL:356
        reset_env_ids = env.reset_buf.nonzero(as_tuple=False).squeeze(-1)  # DRL:359
        if len(reset_env_ids) > 0:
            env._reset_idx(reset_env_ids)  # DRL:361
        obs = env._get_observations()  # DRL:375
    return obs["policy"], reward, env.reset_terminated.clone(), env.reset_time_outs.clone(), reset_env_ids
