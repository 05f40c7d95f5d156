// Development probe (not product code): what bounds the Allsteps step's MEMORY ACCESS PATTERN on B200?
// A skeleton of k_step<fused> with no arithmetic: TMA bulk loads of the seven input tiles, the state word, the
// data-dependent gathers, the TMA bulk store of the observation tile and the per-thread stores.  Each piece can be
// switched off and the occupancy varied, to separate "pattern-bound" from "latency/compute-bound".
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/membound_probe tools/membound_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kJ = 21, kS = 20, kObs = 59;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct Bufs {
  const float *jp, *jv, *act, *rp, *rq, *rv, *body, *cr, *cl;
  const float4 *stones, *window;
  const uint2* state_in;
  uint2* state_out;
  float *obs, *reward;
  uint8_t *term, *tout;
  int64_t n;
};

enum { F_BULK_IN = 1, F_STATE = 2, F_CONTACT = 4, F_STONES = 8, F_WINDOW = 16, F_BULK_OUT = 32, F_SMALL_OUT = 64,
       F_DIRECT_IN = 128, F_DIRECT_OUT = 256, F_TOUCH = 512 };

template <int T>
__global__ void __launch_bounds__(T) k_probe(Bufs b, int flags) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int64_t env0 = (int64_t)blockIdx.x * T;
  const int64_t e = env0 + tid;
  float* s = reinterpret_cast<float*>(smem + 16);
  unsigned long long* mb = reinterpret_cast<unsigned long long*>(smem);
  const uint32_t bar = smem_u32(mb);
  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();
  float acc = 0.f;
  if (flags & F_BULK_IN) {
    if (tid == 0) {
      const uint32_t tx = T * (kJ * 4 * 3 + 12 + 16 + 12 + 36);
      mbar_arrive_expect_tx(bar, tx);
      uint32_t o = smem_u32(s);
      bulk_g2s(o, b.rp + env0 * 3, T * 12, bar); o += T * 12;
      bulk_g2s(o, b.rq + env0 * 4, T * 16, bar); o += T * 16;
      bulk_g2s(o, b.rv + env0 * 3, T * 12, bar); o += T * 12;
      bulk_g2s(o, b.body + env0 * 9, T * 36, bar); o += T * 36;
      bulk_g2s(o, b.jp + env0 * kJ, T * 84, bar); o += T * 84;
      bulk_g2s(o, b.jv + env0 * kJ, T * 84, bar); o += T * 84;
      bulk_g2s(o, b.act + env0 * kJ, T * 84, bar);
    }
  }
  if (flags & F_DIRECT_IN) {  // same bytes with plain coalesced float4 loads by all threads (no TMA, no smem)
    const float4* srcs[7] = {(const float4*)(b.rp + env0 * 3), (const float4*)(b.rq + env0 * 4), (const float4*)(b.rv + env0 * 3),
                             (const float4*)(b.body + env0 * 9), (const float4*)(b.jp + env0 * kJ),
                             (const float4*)(b.jv + env0 * kJ), (const float4*)(b.act + env0 * kJ)};
    const int n4[7] = {T * 3 / 4, T, T * 3 / 4, T * 9 / 4, T * 21 / 4, T * 21 / 4, T * 21 / 4};
#pragma unroll
    for (int a = 0; a < 7; ++a)
      for (int i = tid; i < n4[a]; i += T) { float4 v = __ldg(srcs[a] + i); acc += v.x + v.y + v.z + v.w; }
  }
  int idx = 1;
  if (flags & F_STATE) {
    const uint2 sw = b.state_in[e];
    idx = sw.x % kS;
    acc += __uint_as_float(sw.y);
  } else {
    idx = (int)((e * 2654435761u) >> 7) % kS;
  }
  if (flags & F_WINDOW) {
    const float4* w = b.window + e * 4;
    const float4 w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
    acc += w0.x + w1.y + w2.z + w3.w;
  }
  if (flags & F_CONTACT) {
    const float* r = b.cr + e * 60 + idx * 3;
    const float* l = b.cl + e * 60 + idx * 3;
    acc += __ldg(r) + __ldg(r + 1) + __ldg(r + 2) + __ldg(l) + __ldg(l + 1) + __ldg(l + 2);
  }
  if (flags & F_STONES) {
    const float4* st = b.stones + e * kS;
    const float4 a0 = __ldg(st + max(idx - 1, 0)), a1 = __ldg(st + idx), a2 = __ldg(st + min(idx + 1, kS - 1));
    acc += a0.x + a1.y + a2.z;
  }
  if (flags & F_BULK_IN) {
    mbar_wait(bar, 0);
    if (flags & F_TOUCH) {  // read own rows from smem like the real kernel (63 LDS per thread)
      const float* jp = s + T * 19 + tid * kJ;
#pragma unroll
      for (int j = 0; j < kJ; ++j) acc += jp[j] + jp[T * kJ + j] + jp[2 * T * kJ + j];
    }
  }
  __syncthreads();
  if (flags & F_BULK_OUT) {
    if (flags & F_TOUCH) {
      float* row = s + tid * kObs;
#pragma unroll
      for (int j = 0; j < kObs; ++j) row[j] = acc + j;
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      bulk_s2g(b.obs + env0 * kObs, smem_u32(s), T * kObs * 4);
      bulk_commit();
    }
  }
  if (flags & F_DIRECT_OUT) {
    float4* dst = (float4*)(b.obs + env0 * kObs);
    for (int i = tid; i < T * kObs / 4; i += T) dst[i] = make_float4(acc, acc, acc, acc);
  }
  if (flags & F_SMALL_OUT) {
    b.reward[e] = acc;
    b.term[e] = acc > 1.f;
    b.tout[e] = acc < 0.f;
    b.state_out[e] = make_uint2((uint32_t)idx + 1, __float_as_uint(acc));
  }
  if ((flags & F_BULK_OUT) && tid == 0) bulk_wait_read_all();
}

__global__ void k_init_state(uint2* st, int64_t n) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) st[e] = make_uint2((uint32_t)((e * 2654435761u) >> 7) % kS, 0x3f800000u);
}

int main(int argc, char** argv) {
  const int64_t N = 1 << 20;
  Bufs b{};
  b.n = N;
  auto alloc = [&](size_t bytes) { void* p; CK(cudaMalloc(&p, bytes)); CK(cudaMemset(p, 1, bytes)); return p; };
  b.jp = (float*)alloc(N * 84); b.jv = (float*)alloc(N * 84); b.act = (float*)alloc(N * 84);
  b.rp = (float*)alloc(N * 12); b.rq = (float*)alloc(N * 16); b.rv = (float*)alloc(N * 12); b.body = (float*)alloc(N * 36);
  b.cr = (float*)alloc(N * 240); b.cl = (float*)alloc(N * 240);
  b.stones = (float4*)alloc(N * 320); b.window = (float4*)alloc(N * 64);
  b.state_in = (uint2*)alloc(N * 8); b.state_out = (uint2*)alloc(N * 8);
  b.obs = (float*)alloc(N * 236); b.reward = (float*)alloc(N * 4); b.term = (uint8_t*)alloc(N); b.tout = (uint8_t*)alloc(N);
  k_init_state<<<(unsigned)(N / 256), 256>>>(const_cast<uint2*>(b.state_in), N);
  CK(cudaDeviceSynchronize());
  // a second set of inputs so that consecutive launches do not hit L2 (126 MB): flush with a big memset between
  void* flush = alloc(512ull << 20);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  constexpr int T = 128;
  CK(cudaFuncSetAttribute(k_probe<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  struct Cfg { const char* name; int flags; int smem_kb; double bytes_per_env; };
  const int IN = F_BULK_IN, ST = F_STATE, OUT = F_BULK_OUT | F_SMALL_OUT;
  Cfg cfgs[] = {
      {"bulk in only                         (5 CTA/SM)", IN, 43, 328},
      {"bulk in + bulk out                   (5 CTA/SM)", IN | OUT, 43, 328 + 250},
      {"direct in + direct out (no TMA)      (5 CTA/SM)", F_DIRECT_IN | F_DIRECT_OUT | F_SMALL_OUT, 43, 328 + 250},
      {"direct in + direct out (no TMA)     (16 CTA/SM)", F_DIRECT_IN | F_DIRECT_OUT | F_SMALL_OUT, 1, 328 + 250},
      {"bulk io + state                      (5 CTA/SM)", IN | OUT | ST, 43, 328 + 250 + 8},
      {"bulk io + state + contact            (5 CTA/SM)", IN | OUT | ST | F_CONTACT, 43, 328 + 250 + 8 + 24},
      {"bulk io + state + contact + stones   (5 CTA/SM)", IN | OUT | ST | F_CONTACT | F_STONES, 43, 328 + 250 + 8 + 24 + 36},
      {"bulk io + state + contact + window   (5 CTA/SM)", IN | OUT | ST | F_CONTACT | F_WINDOW, 43, 328 + 250 + 8 + 24 + 64},
      {"bulk io + contact(no state dep)+stones (5 CTA/SM)", IN | OUT | F_CONTACT | F_STONES, 43, 328 + 250 + 24 + 36},
      {"full skeleton + smem touch           (5 CTA/SM)", IN | OUT | ST | F_CONTACT | F_STONES | F_TOUCH, 43, 652},
      {"full skeleton + smem touch           (4 CTA/SM)", IN | OUT | ST | F_CONTACT | F_STONES | F_TOUCH, 54, 652},
      {"full skeleton + smem touch           (3 CTA/SM)", IN | OUT | ST | F_CONTACT | F_STONES | F_TOUCH, 72, 652},
      {"full skeleton + smem touch           (2 CTA/SM)", IN | OUT | ST | F_CONTACT | F_STONES | F_TOUCH, 100, 652},
      {"full skeleton (no touch)             (5 CTA/SM)", IN | OUT | ST | F_CONTACT | F_STONES, 43, 652},
  };
  printf("%-52s %9s %9s %9s\n", "config", "us", "GB/s alg", "B/env");
  for (auto& c : cfgs) {
    float best = 1e9f, sum = 0;
    const int reps = 6;
    for (int r = 0; r < reps + 1; ++r) {
      CK(cudaMemsetAsync(flush, r, 512ull << 20));
      CK(cudaEventRecord(e0));
      k_probe<T><<<(unsigned)(N / T), T, c.smem_kb * 1024>>>(b, c.flags);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (r > 0) { sum += ms; best = ms < best ? ms : best; }
    }
    const float avg = sum / reps;
    printf("%-52s %9.1f %9.1f %9.0f\n", c.name, avg * 1e3, N * c.bytes_per_env / (avg * 1e-3) / 1e9, c.bytes_per_env);
  }
  return 0;
}
