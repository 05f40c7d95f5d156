// Development probe (not product code): how many DRAM bytes does one random 12-byte read out of a (N,20,3) fp32
// matrix cost on B200, per load flavour?  Variants of the contact gather (two matrices, one vector per env each):
//   0  ld.global.nc  (what __ldg emits), one lane per env, 1-2 x 128-bit
//   1  ld.global.nc, two lanes per env, one 128-bit load each in ONE instruction (k_contact_gather_paired)
//   2  as 1 with plain ld.global (L1-allocating, coherent path)
//   3  as 1 with ld.global.cg (L2 only)
//   4  as 1 with ld.global.nc.L1::no_allocate
//   5  as 1 with ld.global.cv (volatile, no caching)
//   6  three scalar ld.global.nc per vector, one lane per env
//   7  as 1 with ld.global.nc.L2::64B  (explicit 64-byte L2 prefetch size)
//   8  TMA bulk copies (cp.async.bulk.shared.global), one 16-byte copy per chunk the vector touches, one lane per env
//   9  TMA bulk copies, one 32-byte copy per vector (the two chunks it can span)
//  10  cp.async.cg (LDGSTS) 16-byte copies global -> shared, no L2 size hint
//  11  cp.async.cg 16-byte copies with L2::64B
// Run under ncu for dram__bytes_read.sum / lts sectors; prints the time per launch itself.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/gather_probe tools/gather_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kRow = 60;  // floats per env row

template <int V>
__device__ __forceinline__ float4 load16(const float4* p) {
  float4 r;
  if (V == 2) asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  else if (V == 3) asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  else if (V == 4) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  else if (V == 5) asm volatile("ld.global.cv.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  else if (V == 7) asm volatile("ld.global.nc.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  else r = __ldg(p);
  return r;
}

template <int V>
__global__ void __launch_bounds__(256) k_paired(const float* cr, const float* cl, const uint8_t* idxs, float2* out, int64_t n) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t e = t >> 1;
  const bool live = e < n;
  const int half = (int)(t & 1);
  const int idx = live ? idxs[e] : 0;
  const int o = idx * 3, k = o & 3;
  const bool fetch = live && (half == 0 || k >= 2);
  float4 r = make_float4(0, 0, 0, 0), l = r;
  if (fetch) {
    r = load16<V>(reinterpret_cast<const float4*>(cr + e * kRow) + (o >> 2) + half);
    l = load16<V>(reinterpret_cast<const float4*>(cl + e * kRow) + (o >> 2) + half);
  }
  const float r1x = __shfl_down_sync(0xffffffffu, r.x, 1), r1y = __shfl_down_sync(0xffffffffu, r.y, 1);
  const float l1x = __shfl_down_sync(0xffffffffu, l.x, 1), l1y = __shfl_down_sync(0xffffffffu, l.y, 1);
  if (!live || half) return;
  float a, b;
  if (k == 0) { a = r.x + r.y + r.z; b = l.x + l.y + l.z; }
  else if (k == 1) { a = r.y + r.z + r.w; b = l.y + l.z + l.w; }
  else if (k == 2) { a = r.z + r.w + r1x; b = l.z + l.w + l1x; }
  else { a = r.w + r1x + r1y; b = l.w + l1x + l1y; }
  out[e] = make_float2(a, b);
}

template <int V>
__global__ void __launch_bounds__(256) k_single(const float* cr, const float* cl, const uint8_t* idxs, float2* out, int64_t n) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int idx = idxs[e];
  float a, b;
  if (V == 6) {
    const float* f = cr + e * kRow + idx * 3;
    const float* g = cl + e * kRow + idx * 3;
    a = __ldg(f) + __ldg(f + 1) + __ldg(f + 2);
    b = __ldg(g) + __ldg(g + 1) + __ldg(g + 2);
  } else {
    const int o = idx * 3, k = o & 3;
    const float4* c = reinterpret_cast<const float4*>(cr + e * kRow) + (o >> 2);
    const float4* d = reinterpret_cast<const float4*>(cl + e * kRow) + (o >> 2);
    float4 c0 = __ldg(c), d0 = __ldg(d), c1 = make_float4(0, 0, 0, 0), d1 = c1;
    if (k >= 2) { c1 = __ldg(c + 1); d1 = __ldg(d + 1); }
    a = c0.x + c0.y + c0.z + c0.w + c1.x + c1.y;
    b = d0.x + d0.y + d0.z + d0.w + d1.x + d1.y;
  }
  out[e] = make_float2(a, b);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane per env; every lane copies the chunk(s) of its two vectors into its own 64-byte slot of shared memory
template <int V>
__global__ void __launch_bounds__(256) k_async(const float* cr, const float* cl, const uint8_t* idxs, float2* out, int64_t n) {
  __shared__ __align__(128) float slot[256][16];   // [0..7] right, [8..15] left
  __shared__ __align__(8) unsigned long long bar;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = e < n;
  const int idx = live ? idxs[e] : 0;
  const int o = idx * 3, k = o & 3;
  const int chunks = (k >= 2) ? 2 : 1;
  const uint32_t b = smem_u32(&bar);
  if (V == 8 || V == 9) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(256) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t bytes = live ? (V == 9 ? 64u : 32u * chunks) : 0u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    if (live) {
      const float* pr = cr + e * kRow + (o & ~3);
      const float* pl = cl + e * kRow + (o & ~3);
      const uint32_t d = smem_u32(&slot[threadIdx.x][0]);
      const uint32_t sz = V == 9 ? 32u : 16u * chunks;
      // (variant 9 may read 16 bytes past the vector's last chunk; the probe's rows are padded by the next row)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(pr), "r"(sz), "r"(b) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d + 32), "l"(pl), "r"(sz), "r"(b) : "memory");
    }
    uint32_t ok;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(0) : "memory");
    } while (!ok);
  } else {
    if (live) {
      const float* pr = cr + e * kRow + (o & ~3);
      const float* pl = cl + e * kRow + (o & ~3);
      const uint32_t d = smem_u32(&slot[threadIdx.x][0]);
      for (int c = 0; c < chunks; ++c) {
        if (V == 10) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16 * c), "l"(pr + 4 * c) : "memory");
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 32 + 16 * c), "l"(pl + 4 * c) : "memory");
        } else {
          asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" ::"r"(d + 16 * c), "l"(pr + 4 * c) : "memory");
          asm volatile("cp.async.cg.shared.global.L2::64B [%0], [%1], 16;" ::"r"(d + 32 + 16 * c), "l"(pl + 4 * c) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
  if (!live) return;
  const float* r = &slot[threadIdx.x][0];
  const float* l = &slot[threadIdx.x][8];
  out[e] = make_float2(r[k] + r[k + 1] + r[k + 2], l[k] + l[k + 1] + l[k + 2]);
}

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : (1 << 20);
  const int sets = 3;  // rotate so that nothing survives in L2 (2 x 252 MB per set)
  float *cr[sets], *cl[sets];
  uint8_t* idxs;
  float2* out;
  for (int s = 0; s < sets; ++s) {
    CK(cudaMalloc(&cr[s], n * kRow * 4));
    CK(cudaMalloc(&cl[s], n * kRow * 4));
    CK(cudaMemset(cr[s], 0, n * kRow * 4));
    CK(cudaMemset(cl[s], 0, n * kRow * 4));
  }
  CK(cudaMalloc(&idxs, n));
  CK(cudaMalloc(&out, n * 8));
  uint8_t* h = (uint8_t*)malloc(n);
  uint32_t x = 12345;
  for (int64_t i = 0; i < n; ++i) { x = x * 1664525u + 1013904223u; h[i] = (uint8_t)((x >> 16) % 20); }
  CK(cudaMemcpy(idxs, h, n, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const char* names[12] = {"nc single 1-2x128", "nc paired", "plain paired", "cg paired", "nc no_allocate paired", "cv paired",
                           "nc single 3x32", "nc L2::64B paired", "TMA bulk 16B chunks", "TMA bulk 32B", "cp.async.cg 16B",
                           "cp.async.cg 16B L2::64B"};
  const int only = argc > 2 ? atoi(argv[2]) : -1;
  for (int v = 0; v < 12; ++v) {
    if (only >= 0 && v != only) continue;
    const unsigned b1 = (unsigned)((n + 255) / 256), b2 = (unsigned)((2 * n + 255) / 256);
    const int reps = 9;
    float ms = 0;
    for (int r = -3; r < reps; ++r) {
      const int s = (r + 3) % sets;
      if (r == 0) CK(cudaEventRecord(e0));
      switch (v) {
        case 0: k_single<0><<<b1, 256>>>(cr[s], cl[s], idxs, out, n); break;
        case 1: k_paired<1><<<b2, 256>>>(cr[s], cl[s], idxs, out, n); break;
        case 2: k_paired<2><<<b2, 256>>>(cr[s], cl[s], idxs, out, n); break;
        case 3: k_paired<3><<<b2, 256>>>(cr[s], cl[s], idxs, out, n); break;
        case 4: k_paired<4><<<b2, 256>>>(cr[s], cl[s], idxs, out, n); break;
        case 5: k_paired<5><<<b2, 256>>>(cr[s], cl[s], idxs, out, n); break;
        case 6: k_single<6><<<b1, 256>>>(cr[s], cl[s], idxs, out, n); break;
        case 7: k_paired<7><<<b2, 256>>>(cr[s], cl[s], idxs, out, n); break;
        case 8: k_async<8><<<b1, 256>>>(cr[s], cl[s], idxs, out, n); break;
        case 9: k_async<9><<<b1, 256>>>(cr[s], cl[s], idxs, out, n); break;
        case 10: k_async<10><<<b1, 256>>>(cr[s], cl[s], idxs, out, n); break;
        case 11: k_async<11><<<b1, 256>>>(cr[s], cl[s], idxs, out, n); break;
      }
    }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("variant %d  %-24s %8.2f us per launch (%lld envs)\n", v, names[v], ms * 1e3 / reps, (long long)n);
  }
  return 0;
}
