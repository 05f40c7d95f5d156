"""CPU: the C-ABI library loads and exports every symbol the header declares; host-side logic; loud failure paths.
No compute call can succeed here (no GPU) -- that is part of what is checked."""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from allsteps_isaaclab_b200 import _cabi, build
from allsteps_isaaclab_b200.config import AllstepsCfg, JOINT_NAMES, NUM_JOINTS
from allsteps_isaaclab_b200.params import make_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _cabi.load()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "allsteps_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(as_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    names = header_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/allsteps_b200.h but not exported"
        assert n in _cabi.SIGNATURES, f"{n} has no ctypes signature in _cabi.py"
    assert sorted(_cabi.SIGNATURES) == names


def test_struct_layouts_match_the_library(lib):
    for i, t in enumerate([_cabi.AsParams, _cabi.AsStateIn, _cabi.AsStepOut, _cabi.AsResetOut, _cabi.AsStats,
                           _cabi.AsMdpState, _cabi.AsMirrorJob, _cabi.AsExchange]):
        assert C.sizeof(t) == lib.as_sizeof(i), t.__name__
    assert lib.as_sizeof(99) == -1
    assert lib.as_abi_version() == _cabi.ABI_VERSION


def test_workspace_size_is_monotonic_and_aligned(lib):
    prev = 0
    for n in (1, 64, 4096, 65536, 1 << 20):
        b = lib.as_workspace_bytes(n)
        assert b % 256 == 0 and b > prev
        prev = b
    assert lib.as_workspace_bytes(0) == 0
    # 1M envs: stones 320 B + stone window 64 B + 2 x 8 B state + 2 x 4 B lists + 16 B contact norms + 2 B grid
    # + 36 B dense body rows (used when the caller's body tensor is strided) + 48 B pass-1 observation tails and
    # 1 B reset flags (3-call path) + 1 bit window-stale flags
    assert lib.as_workspace_bytes(1 << 20) < (1 << 20) * 520


def test_argument_validation_and_loud_failure_without_gpu(lib):
    p = make_params(AllstepsCfg(), seed=1)
    h = C.c_void_p()
    assert lib.as_create(None, 64, 0, 0, None, 0, None, C.byref(h)) == -1
    assert b"null" in lib.as_last_error()
    assert lib.as_create(C.byref(p), 0, 0, 0, None, 0, None, C.byref(h)) == -1
    bad = make_params(AllstepsCfg(), seed=1)
    bad.stop_frames = 7
    assert lib.as_create(C.byref(bad), 64, 0, 0, None, 0, None, C.byref(h)) == -1
    assert b"stop_frames" in lib.as_last_error()
    # workspace too small
    buf = (C.c_uint8 * 1024)()
    assert lib.as_create(C.byref(p), 64, 0, 0, C.addressof(buf), 1024, None, C.byref(h)) == -1
    if not torch.cuda.is_available():
        n = lib.as_workspace_bytes(64)
        raw = np.zeros(n + 256, dtype=np.uint8)
        addr = (raw.ctypes.data + 255) // 256 * 256
        rc = lib.as_create(C.byref(p), 64, 0, 0, addr, n, None, C.byref(h))
        assert rc == -2, "without a CUDA device as_create must fail with AS_ERR_CUDA, never fall back"
        assert lib.as_last_error()


def test_host_object_refuses_cpu():
    from allsteps_isaaclab_b200.mdp import AllstepsMDP

    with pytest.raises(_cabi.AllstepsLibraryError):
        AllstepsMDP(64, device="cpu")


def test_missing_library_is_reported_not_papered_over(tmp_path, monkeypatch):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_cabi.AllstepsLibraryError, match="no CPU or PyTorch fallback"):
        _cabi.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "allsteps_isaaclab_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports oracle/"


def test_params_tables():
    cfg = AllstepsCfg()
    p = make_params(cfg, seed=0x1234567890)
    assert p.seed == 0x1234567890
    assert cfg.max_episode_length == 900 and p.max_episode_length == 900
    th = torch.linspace(0.75, 0.45, 10)
    assert [p.termination_height[i] for i in range(10)] == [float(x) for x in th]
    du = torch.linspace(0.75, 0.9, 10)
    assert [p.dist_upper[i] for i in range(10)] == [float(x) for x in du]
    assert abs(p.step_dt - 4.0 / 240.0) < 1e-9
    assert p.noise_span == np.float32(0.2) and p.noise_lower == np.float32(-0.1)
    # mirror permutation: an involution that swaps right<->left and negates abdomen_z / abdomen_x
    src = [p.mirror_src[j] for j in range(NUM_JOINTS)]
    assert [src[s] for s in src] == list(range(NUM_JOINTS))
    assert src[JOINT_NAMES.index("right_knee")] == JOINT_NAMES.index("left_knee")
    neg = [j for j in range(NUM_JOINTS) if p.mirror_sign[j] < 0]
    assert neg == [JOINT_NAMES.index("abdomen_z"), JOINT_NAMES.index("abdomen_x")] == [0, 8]
    assert cfg.right_joint_indices == (2, 3, 4, 9, 11, 12, 13, 17, 19)
    assert cfg.left_joint_indices == (5, 6, 7, 10, 14, 15, 16, 18, 20)
    assert cfg.body_indices() == (7, 10, 2)
    lim = cfg.joint_limits_rad()
    assert abs(lim[JOINT_NAMES.index("right_knee")][0] + np.deg2rad(150)) < 1e-12
    pose = cfg.reset_joint_pose()
    assert pose[12] == pose[17] == -np.pi / 8 and pose[15] == np.pi / 10


def test_physics_views_validation():
    from allsteps_isaaclab_b200.mdp import PhysicsViews

    N = 8
    good = dict(root_pos_w=torch.zeros(N, 3), root_quat_w=torch.zeros(N, 4), root_lin_vel_w=torch.zeros(N, 3),
                body_pos_w=torch.zeros(N, 17, 3), joint_pos=torch.zeros(N, 21), joint_vel=torch.zeros(N, 21),
                force_matrix_right=torch.zeros(N, 1, 20, 3), force_matrix_left=torch.zeros(N, 1, 20, 3),
                env_origins=torch.zeros(N, 3))
    v = PhysicsViews(**good, body_rows=(7, 10, 2))
    assert v.struct.body_env_stride == 51 and v.struct.body_row_stride == 3 and v.struct.torso_row == 2
    # Isaac Lab's real views: slices of (N,13) root_state_w and (N,B,13) body_state_w
    root_state = torch.zeros(N, 13)
    body_state = torch.zeros(N, 17, 13)
    v2 = PhysicsViews(**{**good, "root_pos_w": root_state[:, 0:3], "root_quat_w": root_state[:, 3:7],
                         "root_lin_vel_w": root_state[:, 7:10], "body_pos_w": body_state[..., 0:3]},
                      body_rows=(7, 10, 2))
    assert v2.struct.root_pos_stride == 13 and v2.struct.root_quat_stride == 13
    assert v2.struct.body_env_stride == 17 * 13 and v2.struct.body_row_stride == 13
    assert v2.struct.root_quat == root_state.data_ptr() + 12
    with pytest.raises(ValueError):
        PhysicsViews(**{**good, "joint_pos": torch.zeros(N, 20)})
    with pytest.raises(TypeError):
        PhysicsViews(**{**good, "root_pos_w": torch.zeros(N, 3, dtype=torch.float64)})
    with pytest.raises(ValueError):
        PhysicsViews(**{**good, "force_matrix_left": torch.zeros(N, 1, 19, 3)})
    with pytest.raises(ValueError):
        PhysicsViews(**good, body_rows=(7, 10, 17))
