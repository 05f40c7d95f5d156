// Checks the straight-line arithmetic of csrc/as_math.cuh against the IEEE operations it replaces, on the GPU.
//
//   sqrt_rn(x)            vs sqrtf(x)      all 2^32 bit patterns, bit for bit (NaN vs NaN counts as equal)
//   div_by_const(n,d,inv) vs n / d         d = step_dt and every joint range of the task, 2^28 numerators each
//   div_with_rcp(a,b,r)   vs a / b         unit-quaternion-like operands: |a| <= 2, b in [1e-9, 4], 2^28 pairs
//
// Build and run (GPU box):  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false \
//                                -o tools/exact_math_check tools/exact_math_check.cu && tools/exact_math_check
// Exit code 0 iff there is no mismatch.  tests/test_gpu_parity.py::test_straight_line_math_is_exact runs it.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../allsteps_isaaclab_b200/csrc/as_math.cuh"

using namespace as;

__device__ unsigned long long g_bad[4];
__device__ unsigned int g_first[4];

__device__ __forceinline__ bool same(float a, float b) {
  if (a != a && b != b) return true;
  return __float_as_uint(a) == __float_as_uint(b);
}

__global__ void k_sqrt_all() {
  const unsigned long long n = 1ull << 32;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const float x = __uint_as_float(static_cast<uint32_t>(i));
    if (!same(sqrt_rn(x), sqrtf(x))) {
      if (atomicAdd(&g_bad[0], 1ull) == 0) g_first[0] = static_cast<uint32_t>(i);
    }
  }
}

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// numerators: every 16th float bit pattern with magnitude in [2^-40, 2^40], both signs, plus zero
__global__ void k_div_const(float d, float inv, int which) {
  const unsigned long long n = 1ull << 28;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t lo = 0x2b800000u, hi = 0x53800000u;  // 2^-40 .. 2^40
    uint32_t bits = lo + static_cast<uint32_t>((mix(static_cast<uint32_t>(i)) % ((hi - lo) >> 0)));
    if (i & 1) bits |= 0x80000000u;
    const float x = (i == 0) ? 0.0f : __uint_as_float(bits);
    if (!same(div_by_const(x, d, inv), x / d)) {
      if (atomicAdd(&g_bad[which], 1ull) == 0) g_first[which] = bits;
    }
  }
}

__global__ void k_div_rcp() {
  const unsigned long long n = 1ull << 28;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t h1 = mix(static_cast<uint32_t>(i)), h2 = mix(h1 ^ 0x9e3779b9u);
    // a: magnitude in [2^-30, 2], random sign (and exact zero now and then); b in [1e-9, 4]
    uint32_t abits = 0x30800000u + h1 % (0x40000000u - 0x30800000u);
    if (h2 & 1) abits |= 0x80000000u;
    float a = __uint_as_float(abits);
    if ((h1 & 0xfffu) == 0) a = 0.0f;
    const uint32_t blo = 0x3089705fu /* 1e-9 */, bhi = 0x40800000u /* 4 */;
    const float b = __uint_as_float(blo + h2 % (bhi - blo));
    const float r = refined_rcp(b);
    const float got = div_with_rcp(a, b, r), want = a / b;
    if (!(got == want)) {  // (+0 == -0: the sign of a zero quotient is not kept, as_math.cuh)
      if (atomicAdd(&g_bad[3], 1ull) == 0) g_first[3] = abits;
    }
  }
}

static float rn_reciprocal(float d) {  // correctly rounded 1/d via double (as_api.cu does the same)
  return static_cast<float>(1.0 / static_cast<double>(d));
}

int main() {
  unsigned long long bad[4] = {0, 0, 0, 0};
  cudaMemcpyToSymbol(g_bad, bad, sizeof(bad));
  k_sqrt_all<<<148 * 8, 256>>>();
  const float step_dt = 4.0f / 240.0f;
  k_div_const<<<148 * 8, 256>>>(step_dt, rn_reciprocal(step_dt), 1);
  // joint ranges of walker3d.xml in radians (upper - lower), the divisors of MATH:22-40
  const float deg = 0.017453292519943295f;
  const float ranges_deg[] = {70, 95, 160, 120, 155, 50, 120, 30, 75, 150, 60};
  for (float rd : ranges_deg) {
    const float d = rd * deg;
    k_div_const<<<148 * 8, 256>>>(d, rn_reciprocal(d), 2);
  }
  k_div_rcp<<<148 * 8, 256>>>();
  if (cudaDeviceSynchronize() != cudaSuccess) {
    std::printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 2;
  }
  unsigned int first[4];
  cudaMemcpyFromSymbol(bad, g_bad, sizeof(bad));
  cudaMemcpyFromSymbol(first, g_first, sizeof(first));
  std::printf("sqrt_rn vs sqrtf, 2^32 inputs:            %llu mismatches (first bits %08x)\n", bad[0], first[0]);
  std::printf("div_by_const vs /, step_dt:               %llu mismatches (first bits %08x)\n", bad[1], first[1]);
  std::printf("div_by_const vs /, joint ranges:          %llu mismatches (first bits %08x)\n", bad[2], first[2]);
  std::printf("div_with_rcp vs /, quaternion operands:   %llu mismatches (first bits %08x)\n", bad[3], first[3]);
  return (bad[0] | bad[1] | bad[2] | bad[3]) ? 1 : 0;
}
