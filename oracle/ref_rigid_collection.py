"""TEST INFRASTRUCTURE ONLY -- executes the reference's own stone-pose egress, `RigidObjectCollection.write_object_pose_to_sim`
(source/isaaclab/isaaclab/assets/rigid_object_collection/rigid_object_collection.py:271-301 with `reshape_data_to_view`
:650-659 and `_env_obj_ids_to_view_ids` :661-680), unmodified, on a plain-tensor stand-in for `self`, and records what it
hands to `root_physx_view.set_transforms(poses, indices=view_ids)`.  Pins SURVEY 8 f4 (egress) to the live reference:
`as_export_stone_poses` must produce the same rows and the same index list.  Only where /root/reference is mounted."""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch

from .ref_loader import REFERENCE_ROOT, Inert, reference_available

_ISAACLAB = os.path.join(REFERENCE_ROOT, "source", "isaaclab", "isaaclab")
_cls = None


def _load_class():
    global _cls
    if _cls is not None:
        return _cls
    if not reference_available():
        raise RuntimeError("reference checkout not mounted")
    ours = ("isaaclab", "omni", "pxr")
    mine = lambda k: k.split(".")[0] in ours  # noqa: E731
    saved = {k: v for k, v in sys.modules.items() if mine(k)}
    for k in saved:
        del sys.modules[k]
    try:
        def shell(name, **attrs):
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
            return m

        def load(name, path):
            spec = importlib.util.spec_from_file_location(name, path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
            return mod

        for n in ("omni", "omni.kit", "omni.kit.app", "omni.log", "omni.physics", "omni.physics.tensors",
                  "omni.physics.tensors.impl", "omni.physics.tensors.impl.api", "omni.timeline"):
            shell(n)
        shell("pxr", UsdPhysics=Inert)
        il = shell("isaaclab")
        il.__path__ = [_ISAACLAB]
        shell("isaaclab.sim")
        utils = shell("isaaclab.utils")
        utils.__path__ = [os.path.join(_ISAACLAB, "utils")]
        utils.math = load("isaaclab.utils.math", os.path.join(_ISAACLAB, "utils", "math.py"))
        utils.string = load("isaaclab.utils.string", os.path.join(_ISAACLAB, "utils", "string.py"))
        assets = shell("isaaclab.assets")
        assets.__path__ = [os.path.join(_ISAACLAB, "assets")]
        shell("isaaclab.assets.asset_base", AssetBase=type("AssetBase", (), {}))
        pkg = shell("isaaclab.assets.rigid_object_collection")
        pkg.__path__ = [os.path.join(_ISAACLAB, "assets", "rigid_object_collection")]
        shell("isaaclab.assets.rigid_object_collection.rigid_object_collection_data",
              RigidObjectCollectionData=type("RigidObjectCollectionData", (), {}))
        mod = load("isaaclab.assets.rigid_object_collection.rigid_object_collection",
                   os.path.join(_ISAACLAB, "assets", "rigid_object_collection", "rigid_object_collection.py"))
        _cls = mod.RigidObjectCollection
    finally:
        for k in [k for k in sys.modules if mine(k)]:
            del sys.modules[k]
        sys.modules.update(saved)
    return _cls


def reference_stone_pose_write(steps_pos: torch.Tensor, env_ids: torch.Tensor):
    """ENV:119-120 (`pose = cat(steps_pos, [1,0,0,0])`, `steps.write_object_pose_to_sim(pose[env_ids], env_ids)`) through
    the reference's method.  Returns (view_poses (S*N,7) x,y,z,w; view_ids (S*k)) as given to `set_transforms`."""
    cls = _load_class()
    N, S, _ = steps_pos.shape
    rec = {}
    me = types.SimpleNamespace()
    me.num_instances, me.num_objects, me.device = N, S, "cpu"
    me._ALL_ENV_INDICES = torch.arange(N, dtype=torch.long)
    me._ALL_OBJ_INDICES = torch.arange(S, dtype=torch.long)
    state = torch.zeros(N, S, 13)
    state[..., 3] = 1.0  # identity quaternions w,x,y,z (the default state of the collection)
    me._data = types.SimpleNamespace(object_state_w=state)
    me.root_physx_view = types.SimpleNamespace(
        set_transforms=lambda poses, indices: rec.update(poses=poses.clone(), indices=indices.clone()))
    me._env_obj_ids_to_view_ids = types.MethodType(cls._env_obj_ids_to_view_ids, me)
    me.reshape_data_to_view = types.MethodType(cls.reshape_data_to_view, me)
    quat = torch.tensor([1.0, 0.0, 0.0, 0.0]).expand(N, S, 4)
    pose = torch.cat((steps_pos, quat), dim=-1)                                  # ENV:119
    cls.write_object_pose_to_sim(me, pose[env_ids], env_ids)                     # ENV:120
    return rec["poses"], rec["indices"]
