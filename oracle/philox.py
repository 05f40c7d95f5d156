"""TEST INFRASTRUCTURE ONLY -- numpy twin of the kernels' counter-based random stream.

The CUDA kernels draw every random number from Philox4x32-10 (Salmon et al., SC'11; the published round
function and constants below) keyed by (seed, global env id, step counter, stream), so a draw does not depend
on which envs reset together or on how envs are sharded over GPUs.  The reference instead calls `torch.rand`
sequentially (ENV:137-141, ENV:518, MATH:1331); the tests feed the reference/port the numbers produced here.

Counter layout (must match allsteps_isaaclab_b200/csrc/philox.cuh):
    counter = (env_id, block, step_lo, stream | step_hi << 8)      key = (seed_lo, seed_hi)
    draw d of a stream lives in lane d % 4 of block d // 4
    uniform = (u32 >> 8) * 2**-24          (24-bit mantissa, in [0, 1))
Streams: 0 = reset (draw 0 mirror coin, draws 1..J joint noise), 1 = stones (draw k*S + s, k in dr,dphi,dtheta,
x_tilt,y_tilt order of ENV:137-141).
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)

STREAM_RESET = 0
STREAM_STONES = 1


def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """counter (...,4) uint32, key (...,2) uint32 -> (...,4) uint32."""
    c = [counter[..., i].astype(np.uint32) for i in range(4)]
    k0 = key[..., 0].astype(np.uint32).copy()
    k1 = key[..., 1].astype(np.uint32).copy()
    with np.errstate(over="ignore"):
        for r in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK32).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            if r != 9:
                k0 = (k0 + W0).astype(np.uint32)
                k1 = (k1 + W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def u32_to_unit_float(x: np.ndarray) -> np.ndarray:
    return ((x >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def stream_uniforms(seed: int, step: int, stream: int, env_ids: np.ndarray, num_draws: int) -> np.ndarray:
    """(len(env_ids), num_draws) float32 uniforms of one stream at one step."""
    env_ids = np.asarray(env_ids, dtype=np.uint32)
    n = env_ids.shape[0]
    nblk = (num_draws + 3) // 4
    counter = np.zeros((n, nblk, 4), dtype=np.uint32)
    counter[..., 0] = env_ids[:, None]
    counter[..., 1] = np.arange(nblk, dtype=np.uint32)[None, :]
    counter[..., 2] = np.uint32(step & 0xFFFFFFFF)
    counter[..., 3] = np.uint32((stream & 0xFF) | (((step >> 32) & 0xFFFFFF) << 8))
    key = np.zeros((n, nblk, 2), dtype=np.uint32)
    key[..., 0] = np.uint32(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    bits = philox4x32_10(counter, key).reshape(n, nblk * 4)
    return u32_to_unit_float(bits[:, :num_draws])


def reset_tables(seed: int, step: int, env_ids: np.ndarray, num_joints: int = 21):
    """(mirror (n,), noise (n,J)) float32 for the reset of `env_ids` at `step`."""
    u = stream_uniforms(seed, step, STREAM_RESET, env_ids, 1 + num_joints)
    return u[:, 0].copy(), u[:, 1:].copy()


def stone_tables(seed: int, step: int, env_ids: np.ndarray, num_stones: int = 20) -> np.ndarray:
    """(5, n, S) float32 in the call order dr, dphi, dtheta, x_tilt, y_tilt."""
    u = stream_uniforms(seed, step, STREAM_STONES, env_ids, 5 * num_stones)
    return np.ascontiguousarray(u.reshape(-1, 5, num_stones).transpose(1, 0, 2))
