// Probe (development tool, not part of the library): issue rate and accumulation exactness of
// tcgen05.mma for kind::i8 (u8 x u8 -> s32, K = 32), kind::f8f6f4 (e4m3 -> f32, K = 32) and
// kind::mxf4 (e2m1 x e2m1, unit ue8m0 block scales -> f32, K = 64) on sm_100a.
//
// Operands are constant-filled shared memory (every element of A equals a, every element of B
// equals b), so the shared-memory layout does not matter: every accumulator element receives
// K * a * b per instruction.  The accumulator is seeded through tcgen05.st with X, n instructions
// are issued back to back by one thread, and the result must be X + n * K * a * b exactly.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp4_probe tools/fp4_probe.cu
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

enum { KIND_I8 = 0, KIND_F8 = 1, KIND_MXF4 = 2 };

template <int KIND>
__global__ void __launch_bounds__(128, 1) k_probe(int nmma, uint32_t seed_bits, uint32_t byteA, uint32_t byteB,
                                                  long long* cycles, uint32_t* out_d, int commit_every,
                                                  int fence_every, int ncols) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sa = smem;             // 16 KB: 128 rows x 128 bytes
  unsigned char* sb = smem + 16384;     // 32 KB: 256 rows x 128 bytes
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[4];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 16384 / 4; i += 128) reinterpret_cast<uint32_t*>(sa)[i] = byteA * 0x01010101u;
  for (int i = threadIdx.x; i < 32768 / 4; i += 128) reinterpret_cast<uint32_t*>(sb)[i] = byteB * 0x01010101u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar2[i])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
  // seed the accumulator (columns 0..255) and fill the scale-factor columns (256..511) with ue8m0 1.0
  for (int c = 0; c < 512; ++c) {
    const uint32_t v = c < 256 ? seed_bits : 0x7F7F7F7Fu;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(lane_addr + (uint32_t)c), "r"(v) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (threadIdx.x == 0) {
    const uint64_t da0 = make_smem_desc(smem_u32(sa));
    const uint64_t db0 = make_smem_desc(smem_u32(sb));
    uint32_t idesc;
    const uint32_t nc = (uint32_t)ncols;
    if (KIND == KIND_I8) idesc = (2u << 4) | ((nc >> 3) << 17) | ((128u >> 4) << 24);
    else if (KIND == KIND_F8) idesc = (1u << 4) | ((nc >> 3) << 17) | ((128u >> 4) << 24);
    else idesc = (1u << 7) | (1u << 10) | ((nc >> 3) << 17) | (1u << 23) | ((128u >> 4) << 24);
    const long long t0 = clock64();
    for (int i = 0; i < nmma; ++i) {
      if (fence_every && (i & 3) == 0) {
        if (fence_every & 2) {      // poll a barrier whose phase is already complete, like a full-stage wait
          uint32_t dn;
          do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(dn) : "r"(smem_u32(&bar2[3])), "r"(1u) : "memory");
          } while (!dn);
        }
        if (fence_every & 1) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const uint64_t da = da0 + (uint64_t)((i & 3) * 2);      // 32 bytes along K inside the 128-byte row
      const uint64_t db = db0 + (uint64_t)((i & 3) * 2);
      if (KIND == KIND_I8) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                     ::"r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(1u), "r"(0u) : "memory");
      } else if (KIND == KIND_F8) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                     ::"r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(1u), "r"(0u) : "memory");
      } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
                     ::"r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(1u), "r"(tmem_base + 256u), "r"(tmem_base + 320u)
                     : "memory");
      }
      if (commit_every && (i & (commit_every - 1)) == commit_every - 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[(i >> 2) & 1])) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    } while (!done);
    cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (blockIdx.x == 0) {
    for (int c = 0; c < 256; ++c) {
      uint32_t v;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(lane_addr + (uint32_t)c) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      out_d[(warp * 32 + lane) * 256 + c] = v;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

static long long* d_cycles;
static uint32_t* d_out;

struct Result { double cyc_per_mma; bool uniform; uint32_t first; };

static Result run(int kind, int nmma, uint32_t seed_bits, uint32_t a, uint32_t b, int grid, int commit_every = 0,
                  int fence_every = 0, int ncols = 256) {
  const size_t smem = 49152 + 1024;
  switch (kind) {
    case KIND_I8: CK(cudaFuncSetAttribute(k_probe<KIND_I8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_probe<KIND_I8><<<grid, 128, smem>>>(nmma, seed_bits, a, b, d_cycles, d_out, commit_every, fence_every, ncols); break;
    case KIND_F8: CK(cudaFuncSetAttribute(k_probe<KIND_F8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_probe<KIND_F8><<<grid, 128, smem>>>(nmma, seed_bits, a, b, d_cycles, d_out, commit_every, fence_every, ncols); break;
    default: CK(cudaFuncSetAttribute(k_probe<KIND_MXF4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); k_probe<KIND_MXF4><<<grid, 128, smem>>>(nmma, seed_bits, a, b, d_cycles, d_out, commit_every, fence_every, ncols); break;
  }
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<long long> cyc(grid);
  std::vector<uint32_t> out(128 * 256);
  CK(cudaMemcpy(cyc.data(), d_cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (long long c : cyc) mx = c > mx ? c : mx;
  Result r;
  r.cyc_per_mma = (double)mx / nmma;
  r.first = out[0];
  r.uniform = true;
  for (int l = 0; l < 128; ++l) for (int c = 0; c < ncols; ++c) if (out[l * 256 + c] != out[0]) r.uniform = false;
  return r;
}

static float bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static uint32_t f2bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

int main() {
  CK(cudaMalloc(&d_cycles, 1024 * sizeof(long long)));
  CK(cudaMalloc(&d_out, 128 * 256 * 4));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("%s, %d SMs\n", prop.name, prop.multiProcessorCount);
  const int grid = prop.multiProcessorCount;
  const char* names[3] = {"i8 (K32)", "f8f6f4 e4m3 (K32)", "mxf4 e2m1 (K64)"};
  const uint32_t one[3] = {0x01, 0x38, 0x22};
  const int kdim[3] = {32, 32, 64};
  // ---- issue rate: all SMs busy, 8192 back-to-back instructions per SM
  for (int rep = 0; rep < 2; ++rep)
    for (int kind = 0; kind < 3; ++kind) {
      Result r = run(kind, 8192, 0u, one[kind], one[kind], grid);
      const double expect = 8192.0 * kdim[kind];
      const double got = kind == KIND_I8 ? (double)(int)r.first : (double)bits2f(r.first);
      printf("rate %-20s %7.1f cycles per M128 N256 instruction (%d SMs busy); D = %.1f (expected %.1f) uniform %d\n",
             names[kind], r.cyc_per_mma, grid, got, expect, (int)r.uniform);
    }
  // ---- what a per-stage tcgen05.commit / fence costs the issue rate
  for (int kind : {0, 2})
    for (int ce : {0, 1, 2, 4, 8, 16})
      for (int fe : {0, 1, 2, 3}) {
        Result r = run(kind, 8192, 0u, one[kind], one[kind], grid, ce, fe);
        printf("rate %-20s commit every %2d, per 4 instructions: fence(1) | barrier poll(2) = %d: %7.1f cycles per instruction\n", names[kind], ce, fe, r.cyc_per_mma);
      }
  for (int kind : {0, 2})
    for (int nc : {64, 128, 192, 256}) {
      Result r = run(kind, 8192, 0u, one[kind], one[kind], grid, 0, 0, nc);
      printf("rate %-20s N = %3d: %7.1f cycles per instruction (uniform %d)\n", names[kind], nc, r.cyc_per_mma, (int)r.uniform);
    }
  // ---- exactness of the f32 accumulate: seed X (odd, just below 2^m), add n instructions of all-ones
  for (int kind = 1; kind < 3; ++kind) {
    printf("exactness %s: X + n*K (all-ones operands), X = 2^m - 4*K - 1\n", names[kind]);
    for (int m = 8; m <= 24; ++m) {
      const float X = ldexpf(1.0f, m) - 4.0f * kdim[kind] - 1.0f;
      if (X < 0) continue;
      int ok_all = 1;
      for (int n = 1; n <= 4; ++n) {
        Result r = run(kind, n, f2bits(X), one[kind], one[kind], 1);
        const float expect = X + (float)(n * kdim[kind]);
        if (!(r.uniform && bits2f(r.first) == expect)) { ok_all = 0; printf("   m=%d n=%d: got %.1f expected %.1f uniform %d\n", m, n, bits2f(r.first), expect, (int)r.uniform); }
      }
      printf("  m=%2d X=%.1f %s\n", m, X, ok_all ? "exact" : "INEXACT");
    }
    // products of different magnitude in one accumulator: 6*6 (mxf4) / 16*16 (e4m3 0x58 = 16) then ones
    const uint32_t big = kind == KIND_MXF4 ? 0x77 : 0x58;
    const float bigv = kind == KIND_MXF4 ? 36.0f : 256.0f;
    for (int n : {1, 16, 256, 1024}) {
      if (n * kdim[kind] * bigv >= 16777216.0f) continue;
      Result r = run(kind, n, f2bits(1.0f), big, big, 1);
      const float expect = 1.0f + n * kdim[kind] * bigv;
      printf("  1 + %d instructions of %.0f-products: got %.1f expected %.1f %s\n", n, bigv, bits2f(r.first), expect,
             (r.uniform && bits2f(r.first) == expect) ? "exact" : "INEXACT");
    }
    // long accumulation from 0 to just under 2^24 by all-ones instructions
    {
      const int n = (1 << 24) / kdim[kind] - 1;
      Result r = run(kind, n, 0u, one[kind], one[kind], 1);
      printf("  %d all-ones instructions from 0: got %.1f expected %.1f %s\n", n, bits2f(r.first), (double)n * kdim[kind],
             (r.uniform && bits2f(r.first) == (float)n * kdim[kind]) ? "exact" : "INEXACT");
    }
  }
  // mxf4 with half-valued operands: 0.5 * 0.5 * 64 = 16 per instruction onto a large accumulator
  for (int m : {12, 16, 20, 23}) {
    const float X = ldexpf(1.0f, m) + 1.0f;
    Result r = run(KIND_MXF4, 2, f2bits(X), 0x11, 0x11, 1);
    printf("  mxf4 0.25-products onto X=2^%d+1: got %.1f expected %.1f\n", m, bits2f(r.first), X + 32.0f);
  }
  return 0;
}
