"""TEST INFRASTRUCTURE ONLY -- loads the reference's UNMODIFIED manager framework (`isaaclab/managers/*.py`, SURVEY 2.1:
`ManagerBase._resolve_common_term_cfg` manager_base.py:219-298, `ObservationManager`, `RewardManager`,
`TerminationManager`, `EventManager`, `CurriculumManager`, the term-cfg classes and `SceneEntityCfg`) from
/root/reference so that the B2 face (allsteps_isaaclab_b200/terms.py + manager_cfg.py) can be validated against the
code that would really call it.  Only where the reference checkout is mounted (the build container).

The real modules are executed as they are; what they import from outside `managers/` is provided as follows:
  isaaclab.utils.{string,array,dict,configclass,modifiers,noise,buffers}   the reference's own files (pure Python/torch)
  warp                      a shell (`isaaclab/utils/array.py` only names `wp.array` in a type union)
  omni.log                  print-free no-ops
  prettytable.PrettyTable   a minimal table (only the managers' __str__ uses it)
  isaaclab.assets / isaaclab.scene                class shells for the isinstance checks of SceneEntityCfg.resolve
Every sys.modules entry this loader creates is removed again afterwards (the loaded module objects keep working), so
it cannot interfere with oracle/ref_loader.py's own stand-ins for the same package names.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

from .ref_loader import REFERENCE_ROOT, reference_available

_ISAACLAB = os.path.join(REFERENCE_ROOT, "source", "isaaclab", "isaaclab")
_loaded = None


class PrettyTable:
    def __init__(self, *a, **k):
        self.title = ""
        self.field_names = []
        self.align = {}
        self.rows = []

    def add_row(self, row):
        self.rows.append(list(row))

    def get_string(self):
        head = " | ".join(str(f) for f in self.field_names)
        return "\n".join([str(self.title), head] + [" | ".join(str(c) for c in r) for r in self.rows])

    __str__ = get_string


def _exec(name: str, path: str, is_pkg: bool = False):
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[os.path.dirname(path)]
                                                  if is_pkg else None)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_managers():
    """Returns the reference's `isaaclab.managers` namespace: ObservationManager, RewardManager, TerminationManager,
    EventManager, CurriculumManager, *TermCfg, ObservationGroupCfg, SceneEntityCfg, ManagerTermBase, configclass."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference checkout not found under {REFERENCE_ROOT}")
    ours = ("isaaclab", "omni", "warp", "prettytable")  # package names this loader stands in for
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

        shell("warp", array=type("array", (), {}))
        omni = shell("omni")
        omni.log = shell("omni.log", info=lambda *a, **k: None, warn=lambda *a, **k: None,
                         error=lambda *a, **k: None, verbose=lambda *a, **k: None)
        shell("prettytable", PrettyTable=PrettyTable)
        isaaclab = shell("isaaclab")
        isaaclab.__path__ = [_ISAACLAB]
        shell("isaaclab.assets", Articulation=type("Articulation", (), {}), RigidObject=type("RigidObject", (), {}),
              RigidObjectCollection=type("RigidObjectCollection", (), {}))
        shell("isaaclab.scene", InteractiveScene=type("InteractiveScene", (), {}))
        u = os.path.join(_ISAACLAB, "utils")
        utils = shell("isaaclab.utils")
        utils.__path__ = [u]
        utils.string = _exec("isaaclab.utils.string", os.path.join(u, "string.py"))
        utils.array = _exec("isaaclab.utils.array", os.path.join(u, "array.py"))
        utils.dict = _exec("isaaclab.utils.dict", os.path.join(u, "dict.py"))
        cc = _exec("isaaclab.utils.configclass", os.path.join(u, "configclass.py"))
        utils.configclass = cc.configclass
        utils.string_to_callable = utils.string.string_to_callable
        utils.buffers = _exec("isaaclab.utils.buffers", os.path.join(u, "buffers", "__init__.py"), is_pkg=True)
        utils.modifiers = _exec("isaaclab.utils.modifiers", os.path.join(u, "modifiers", "__init__.py"), is_pkg=True)
        utils.noise = _exec("isaaclab.utils.noise", os.path.join(u, "noise", "__init__.py"), is_pkg=True)
        m = os.path.join(_ISAACLAB, "managers")
        pkg = shell("isaaclab.managers")
        pkg.__path__ = [m]
        ns = types.SimpleNamespace(configclass=cc.configclass)
        for name in ("scene_entity_cfg", "manager_term_cfg", "manager_base", "observation_manager", "reward_manager",
                     "termination_manager", "event_manager", "curriculum_manager"):
            mod = _exec(f"isaaclab.managers.{name}", os.path.join(m, f"{name}.py"))
            setattr(pkg, name, mod)
            setattr(ns, name, mod)
        for mod_name, names in (
                ("scene_entity_cfg", ["SceneEntityCfg"]),
                ("manager_term_cfg", ["ObservationTermCfg", "ObservationGroupCfg", "RewardTermCfg", "TerminationTermCfg",
                                      "EventTermCfg", "CurriculumTermCfg", "ManagerTermBaseCfg"]),
                ("manager_base", ["ManagerBase", "ManagerTermBase"]),
                ("observation_manager", ["ObservationManager"]), ("reward_manager", ["RewardManager"]),
                ("termination_manager", ["TerminationManager"]), ("event_manager", ["EventManager"]),
                ("curriculum_manager", ["CurriculumManager"])):
            for n in names:
                setattr(ns, n, getattr(getattr(ns, mod_name), n))
        ns.Articulation = sys.modules["isaaclab.assets"].Articulation
        _loaded = ns
    finally:
        for k in [k for k in sys.modules if mine(k)]:
            del sys.modules[k]
        sys.modules.update(saved)
    return _loaded


def load_env_methods(rel_path: str, class_name: str, names):
    """Methods of a reference env class as plain functions of `self`: the bodies are compiled from the reference's own
    source text -- unmodified but for the argument / return annotations, which name types of modules that cannot be
    imported here (`gymnasium`, the simulator) -- so that a test can let the REFERENCE decide the order of the calls
    around a step instead of replaying that order by hand."""
    import ast

    import torch

    path = os.path.join(_ISAACLAB, rel_path)
    with open(path) as f:
        tree = ast.parse(f.read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name)
    out = {}
    for node in cls.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            node.returns = None
            for a in node.args.args + node.args.kwonlyargs:
                a.annotation = None
            ns = {"torch": torch}
            exec(compile(ast.fix_missing_locations(ast.Module(body=[node], type_ignores=[])), path, "exec"), ns)
            out[node.name] = ns[node.name]
    assert set(out) == set(names), "reference layout changed"
    return out


def load_rl_env_methods():
    """`ManagerBasedRLEnv.step` and `._reset_idx` (manager_based_rl_env.py:153-239, 347-392)."""
    return load_env_methods(os.path.join("envs", "manager_based_rl_env.py"), "ManagerBasedRLEnv", ("step", "_reset_idx"))


def load_direct_env_methods():
    """`DirectRLEnv.reset`, `.step` and `._reset_idx` (direct_rl_env.py:256-294, 296-383, 563-584)."""
    return load_env_methods(os.path.join("envs", "direct_rl_env.py"), "DirectRLEnv", ("reset", "step", "_reset_idx"))
