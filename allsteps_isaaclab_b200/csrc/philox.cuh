// Counter-based random stream of the Allsteps kernels: Philox4x32-10 (Salmon et al., SC'11).
//
// The reference draws with sequential `torch.rand` calls (allsteps_env.py:137-141, :518; math.py:1331), whose
// values depend on how many envs reset together.  Here every draw is a pure function of
// (seed, global env id, step counter, stream, draw index), so results do not depend on the reset set, the
// launch geometry or the env-id sharding over GPUs.  oracle/philox.py is the numpy twin used by the tests.
//
//   counter = (env_id, block, step_lo, stream | step_hi << 8)      key = (seed_lo, seed_hi)
//   draw d of a stream = lane d % 4 of block d / 4;   uniform = (u32 >> 8) * 2^-24  in [0, 1)
#pragma once
#include <cstdint>

namespace as {

constexpr uint32_t kStreamReset = 0;   // draw 0: mirror coin, draws 1..J: joint noise
constexpr uint32_t kStreamStones = 1;  // draw k*S + s, k in (dr, dphi, dtheta, x_tilt, y_tilt)

struct PhiloxKey {
  uint32_t k0, k1;
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, PhiloxKey key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t k0 = key.k0, k1 = key.k1;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += W0;
    k1 += W1;
  }
  return c;
}

__device__ __forceinline__ float u32_to_unit(uint32_t x) { return static_cast<float>(x >> 8) * 5.9604644775390625e-08f; }

// Four consecutive draws (one Philox block) of `stream` for `env_id` at `step`.
__device__ __forceinline__ uint4 philox_block(uint64_t seed, uint64_t step, uint32_t stream, uint32_t env_id,
                                              uint32_t block) {
  const uint4 ctr = make_uint4(env_id, block, static_cast<uint32_t>(step),
                               (stream & 0xFFu) | ((static_cast<uint32_t>(step >> 32) & 0xFFFFFFu) << 8));
  const PhiloxKey key{static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)};
  return philox4x32_10(ctr, key);
}

__device__ __forceinline__ uint32_t lane_of(const uint4& b, int lane) {
  return lane == 0 ? b.x : lane == 1 ? b.y : lane == 2 ? b.z : b.w;
}

// Single draw `d` of a stream.
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint64_t step, uint32_t stream, uint32_t env_id,
                                                int d) {
  const uint4 b = philox_block(seed, step, stream, env_id, static_cast<uint32_t>(d >> 2));
  return u32_to_unit(lane_of(b, d & 3));
}

}  // namespace as
