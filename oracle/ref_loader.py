"""TEST INFRASTRUCTURE ONLY -- load the UNMODIFIED reference task code from /root/reference.

The reference (`source/isaaclab_tasks/isaaclab_tasks/direct/allsteps/allsteps_env.py`) imports Isaac Sim
bound packages (`isaaclab.sim`, `isaaclab.assets`, `gymnasium`, ...) that do not exist outside an Isaac Sim
install.  None of them is touched by the MDP arithmetic, so they are replaced by inert stand-ins and the two
files that *are* the hot path are executed as they lie on disk:

* ``isaaclab/utils/math.py``        -- real module (depends on torch/numpy only)
* ``direct/allsteps/allsteps_env.py`` and ``allsteps_env_cfg.py`` -- real modules

Nothing is copied: the files are exec'd from the read-only checkout.  This loader only works where
``/root/reference`` is mounted (the build container); the GPU box uses the committed golden vectors.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ALLSTEPS_REFERENCE_ROOT", "/root/reference")
_SRC = os.path.join(REFERENCE_ROOT, "source")
_TASK_DIR = os.path.join(_SRC, "isaaclab_tasks", "isaaclab_tasks", "direct", "allsteps")
_MATH_PY = os.path.join(_SRC, "isaaclab", "isaaclab", "utils", "math.py")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_TASK_DIR, "allsteps_env.py")) and os.path.isfile(_MATH_PY)


class _InertMeta(type):
    """Class-level attribute access (``RigidObjectCfg.InitialStateCfg``) also yields an inert class."""

    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return Inert


class Inert(metaclass=_InertMeta):
    """Swallows construction, attribute access, calls and ``.replace(...)``."""

    def __init__(self, *args, **kwargs):
        pass

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return Inert()

    def __call__(self, *args, **kwargs):
        return Inert()

    def replace(self, **kwargs):
        return self


class _BaseDirectRLEnv:
    """Stand-in for DirectRLEnv: only `_reset_idx` matters to the task code.

    Mirrors what the real base does to the tensors the MDP reads
    (direct_rl_env.py:563-584 -> scene.reset -> contact_sensor.py:142-161):
    contact rows of the reset envs are zeroed and the episode counter restarts.
    """

    def _reset_idx(self, env_ids):
        self.sensor_left.data.force_matrix_w[env_ids] = 0.0
        self.sensor_right.data.force_matrix_w[env_ids] = 0.0
        self.episode_length_buf[env_ids] = 0


class _BaseDirectRLEnvCfg:
    pass


_loaded = None


def _module(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def _exec_file(name: str, path: str) -> types.ModuleType:
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    """Returns a namespace with `math` (module), `AllstepsEnv`, `AllstepsEnvCfg`, `env_module`."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference checkout not found under {REFERENCE_ROOT}")

    mathmod = _exec_file("isaaclab.utils.math", _MATH_PY)

    inert_names = lambda *names: {n: Inert for n in names}  # noqa: E731
    _module("gymnasium")
    isaaclab = _module("isaaclab")
    isaaclab.sim = _module(
        "isaaclab.sim",
        **inert_names(
            "SimulationCfg", "CuboidCfg", "RigidBodyPropertiesCfg", "CollisionPropertiesCfg",
            "PreviewSurfaceCfg", "RigidBodyMaterialCfg", "CylinderCfg", "DomeLightCfg",
        ),
    )
    _module("isaaclab.sim.spawners")
    _module("isaaclab.sim.spawners.from_files", **inert_names("GroundPlaneCfg", "spawn_ground_plane"))
    _module(
        "isaaclab.assets",
        **inert_names(
            "Articulation", "RigidObject", "RigidObjectCollection",
            "ArticulationCfg", "RigidObjectCfg", "RigidObjectCollectionCfg",
        ),
    )
    _module("isaaclab.envs", DirectRLEnv=_BaseDirectRLEnv, DirectRLEnvCfg=_BaseDirectRLEnvCfg)
    _module("isaaclab.utils", configclass=lambda cls: cls, math=mathmod)
    _module("isaaclab.markers", **inert_names("VisualizationMarkers", "VisualizationMarkersCfg"))
    _module("isaaclab.sensors", **inert_names("ContactSensor", "ContactSensorCfg"))
    _module("isaaclab.scene", **inert_names("InteractiveSceneCfg"))
    _module("isaaclab.terrains", **inert_names("TerrainImporterCfg"))
    _module("isaaclab_assets", WALKER_CFG=Inert(), HUMANOID_28_CFG=Inert(), HUMANOID_CFG=Inert())
    _module("isaaclab_rl")
    _module("isaaclab_rl.rsl_rl")
    _module("isaaclab_rl.rsl_rl.vecenv_wrapper", RslRlVecEnvWrapper=Inert)
    _module("isaaclab_rl.rl_games", RlGamesVecEnvWrapper=Inert)
    _module("isaaclab_tasks")
    _module("isaaclab_tasks.direct")
    pkg = _module("isaaclab_tasks.direct.allsteps")
    pkg.__path__ = [_TASK_DIR]

    cfgmod = _exec_file(
        "isaaclab_tasks.direct.allsteps.allsteps_env_cfg", os.path.join(_TASK_DIR, "allsteps_env_cfg.py")
    )
    envmod = _exec_file(
        "isaaclab_tasks.direct.allsteps.allsteps_env", os.path.join(_TASK_DIR, "allsteps_env.py")
    )
    _loaded = types.SimpleNamespace(
        math=mathmod,
        env_module=envmod,
        AllstepsEnv=envmod.AllstepsEnv,
        AllstepsEnvCfg=cfgmod.AllstepsEnvCfg,
        BaseEnv=_BaseDirectRLEnv,
    )
    return _loaded
