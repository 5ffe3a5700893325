// Device-side building blocks shared by the dense tensor-core kernels (cv_dense.cu: u8 x u8 -> s32,
// cv_dense_fp4.cu: e2m1 x e2m1 -> f32): PTX wrappers for mbarrier / TMA / tcgen05, the tile raster,
// and the fused epilogue (R/compute_deviations.R:488-520 on the exact integer sums of one cell).
#pragma once
#include <cuda.h>

#include "cv_internal.cuh"

int cv_make_map_u8(cv_handle* h, CUtensorMap* map, void* base, uint64_t rows, uint64_t row_bytes,
                   uint64_t row_stride, uint32_t box_rows);

#define DN_BLOCK_M 128      // cells per tile (TMEM lanes)
#define DN_BLOCK_N 256      // stacked (annotation, iteration) rows per tile (TMEM columns)
#define DN_BLOCK_K 128      // peaks per pipeline stage (bytes of one SWIZZLE_128B row)
#define DN_UMMA_K 32        // peaks per tcgen05.mma.kind::i8
#define DN_STAGES 4
#define DN_THREADS 256
#define DN_MAX_G 16        // annotations per accumulator tile (bounded by the epilogue scratch)
#define DN_A_BYTES (DN_BLOCK_M * DN_BLOCK_K)
#define DN_B_BYTES (DN_BLOCK_N * DN_BLOCK_K)
#define DN_STAGE_BYTES (DN_A_BYTES + DN_B_BYTES)
#define DN_SMEM_BYTES (DN_STAGES * DN_STAGE_BYTES + 1024 /*align*/ + 4096 /*winv*/ + 256 /*barriers*/)
// CTA-pair kernel: each CTA stages its 128 cells and half of the 256 stacked rows
#define DN_STAGES2 6
#define DN_STAGE2_BYTES (2 * DN_A_BYTES)
#define DN_SMEM2_BYTES (DN_STAGES2 * DN_STAGE2_BYTES + 1024 + 4096 + 256)

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// waits that are long by design (epilogue warps wait a whole tile for the accumulator, producers
// wait for a free stage): back off between polls so the polling warps do not burn issue slots and
// power next to the tensor pipe
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, unsigned ns) {
  uint32_t done;
  const uint32_t addr = smem_u32(bar);
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(ns);
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
// one lane of a converged warp; the code around it runs on all 32 lanes with warp-uniform values, so
// the descriptors of the tcgen05 instructions stay in uniform registers (a lane-0-only branch makes
// the compiler move every operand through an ELECT / R2UR.BROADCAST loop per instruction)
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
  return p != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], u8 x u8 -> s32
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the cells operand read from tensor memory (128 lanes x 32 bytes)
__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
// K-major operand tile in the canonical SWIZZLE_128B layout TMA produces: rows of 128 bytes,
// 8-row groups 1024 bytes apart (SBO), 16-byte chunks XOR-swizzled by (row % 8).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                               // leading byte offset (unused for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                               // layout type SWIZZLE_128B
  return d;
}
// instruction descriptor: dense, no saturate, D = S32, A = B = UINT8, both K-major, N, M
__host__ __device__ constexpr uint32_t make_idesc_u8(int M, int N) {
  return (2u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- cluster / CTA-pair PTX wrappers --------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA) inside CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of one CTA of a pair: data lands in this CTA's shared memory, the transaction bytes
// are counted on the leader CTA's mbarrier (bar_cluster_addr is a shared::cluster address).
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst,
                                                 int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(x), "r"(y)
      : "memory");
}
// arrive (once the pair's MMAs issued so far have retired) on the barrier at this shared-memory
// offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
// D[tmem of both CTAs] (+)= A[256 x 32: 128 rows from each CTA] * B[32 x 256: 128 columns from each CTA]
__device__ __forceinline__ void umma_i8_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8, %9, %10, %11, %12}, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u),
        "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}

// ---- the GEMM + fused epilogue -----------------------------------------------------------------------
// epilogue arithmetic: FUSED = fp64 throughout; FUSED32 = per-iteration statistics without fp64:
// the mean over iterations in exact 64-bit fixed point (sum_j s_j * W_j, W_j = 1 / E_j scaled to 32
// bits), the variance in fp32 of the exact integer differences; fp64 only once per annotation;
// SUMS = raw integer sums (digit planes).
// FP64 instructions issue ~25x slower while the tensor pipe is busy (stall_math_pipe_throttle; the
// same code runs at 37 TFLOP/s alone, tools/fp64_probe.cu), so the fp64 epilogue of a tile takes
// ~265 us -- hidden behind the next tile's MMAs only when a tile has >= ~1000 k-blocks.
enum { DN_EPI_FUSED = 0, DN_EPI_SUMS = 1, DN_EPI_FUSED32 = 2 };

struct DenseParams {
  int N, K, B, G;            // cells (whole shard), annotations, iterations, annotations per 256-row group
  int cb0;                   // global index of this launch's first 128-cell block
  int c_end;                 // cells >= c_end are padding of this launch (no output)
  int n_cell_blocks;         // 128-cell blocks in this launch (chunk of the shard)
  int ncb_total;             // 128-cell blocks of the whole shard (stride of varpart)
  int n_groups, n_kblocks;
  int64_t n_tiles;           // work items: (cell block | cell block pair) x group
  int band;                  // groups per raster band
  const int64_t* annoptr;
  const double* etab;        // [K][B+1]
  const double* fps;         // [N]
  double* dev_rm;            // [K][N]
  double* z_rm;
  double* varpart;           // [K][ncb_total][4]
  long long* obs;            // [K][N] or null
  long long* samp;           // [K][B][N] or null
  // DN_EPI_SUMS: exact integer sums of one u8 digit plane, added into sums[(k*(B+1)+j)*sums_ld + (c - sums_c0)]
  long long* sums;
  int sums_ld, sums_c0, sums_shift, sums_acc;
  int debug_no_tma;          // experiment: the producer only signals the barrier (results are garbage)
  const uint32_t* winv_fx;   // [K][B+1] W[k][j] = round(2^F_k / E[k][j]) < 2^32 (DN_EPI_FUSED32)
  const double* wscale;      // [K] 2^-F_k
  // e2m1 path (cv_dense_fp4.cu): annotations per 256- / 208-row sub-tile, super-groups, and the
  // number of 256-peak k-blocks of this chunk (main + overflow columns), written on the device
  int Ga, Gb, n_sg;
  const uint32_t* n_kblocks_dev;
  uint32_t* sync_ctr;        // one counter per (wave of work items, sub-tile): producers start together
  int sync_lag;              // 0: wait for this sub-tile's counter, 1: for the previous sub-tile's
};

struct DWelford { double n, mean, m2; };
__device__ __forceinline__ DWelford dw_merge(const DWelford& a, const DWelford& b) {
  if (b.n == 0.0) return a;
  if (a.n == 0.0) return b;
  DWelford r;
  r.n = a.n + b.n;
  double d = b.mean - a.mean;
  r.mean = a.mean + d * (b.n / r.n);
  r.m2 = a.m2 + b.m2 + d * d * (a.n * b.n / r.n);
  return r;
}
// deterministic, lane-symmetric tree: every lane ends with the same bits
__device__ __forceinline__ void dw_warp_reduce(DWelford& wf, double& nonfinite, int lane) {
  const unsigned full = 0xffffffffu;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    DWelford b;
    b.n = __shfl_xor_sync(full, wf.n, o);
    b.mean = __shfl_xor_sync(full, wf.mean, o);
    b.m2 = __shfl_xor_sync(full, wf.m2, o);
    wf = (lane & o) ? dw_merge(b, wf) : dw_merge(wf, b);
    nonfinite += __shfl_xor_sync(full, nonfinite, o);
  }
}

// (n, mean, M2, #non-finite) of the z values of one warp's 32 cells: sums shifted by the first finite
// value (xor-shuffle trees: every lane ends with the same bits), converted once.  Non-live lanes
// (padding cells) do not count.
__device__ __forceinline__ void dn_warp_variability(double zv, bool live, int lane, double* w4) {
  const unsigned full = 0xffffffffu;
  const bool fin = live && isfinite(zv);
  const unsigned mfin = __ballot_sync(full, fin);
  const unsigned mbad = __ballot_sync(full, live && !fin);
  double s1 = 0.0, s2 = 0.0, p = 0.0;
  if (mfin) {
    p = __shfl_sync(full, zv, __ffs(mfin) - 1);
    const double d = fin ? zv - p : 0.0;
    s1 = d;
    s2 = d * d;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(full, s1, o);
      s2 += __shfl_xor_sync(full, s2, o);
    }
  }
  if (lane == 0) {
    const double n = (double)__popc(mfin);
    double mean = 0.0, m2 = 0.0;
    if (mfin) {
      mean = p + s1 / n;
      m2 = s2 - s1 * (s1 / n);
      if (m2 < 0.0) m2 = 0.0;
    }
    w4[0] = n; w4[1] = mean; w4[2] = m2; w4[3] = (double)__popc(mbad);
  }
}

// the same with the within-warp sums in fp32 (deviations from the first finite value; 32 terms):
// two fp64 operations per lane instead of twelve -- fp64 is throttled while the tensor pipe runs
__device__ __forceinline__ void dn_warp_variability_f32(double zv, bool live, int lane, double* w4) {
  const unsigned full = 0xffffffffu;
  const bool fin = live && isfinite(zv);
  const unsigned mfin = __ballot_sync(full, fin);
  const unsigned mbad = __ballot_sync(full, live && !fin);
  double p = 0.0;
  float s1 = 0.f, s2 = 0.f;
  if (mfin) {
    p = __shfl_sync(full, zv, __ffs(mfin) - 1);
    const float d = fin ? (float)(zv - p) : 0.f;
    s1 = d;
    s2 = d * d;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(full, s1, o);
      s2 += __shfl_xor_sync(full, s2, o);
    }
  }
  if (lane == 0) {
    const double n = (double)__popc(mfin);
    double mean = 0.0, m2 = 0.0;
    if (mfin) {
      mean = p + (double)s1 / n;
      m2 = (double)s2 - (double)s1 * ((double)s1 / n);
      if (m2 < 0.0) m2 = 0.0;
    }
    w4[0] = n; w4[1] = mean; w4[2] = m2; w4[3] = (double)__popc(mbad);
  }
}

// per (annotation, cell) running state over the B+1 stacked sums (R/compute_deviations.R:488-520)
struct DnCell { double obs_v, pivot, sum, sumsq; };
__device__ __forceinline__ void dn_cell_add(DnCell& a, int j, double sv, double winv_j, double rf) {
  if (j == 0) {
    a.obs_v = sv;
    a.sum = 0.0;
    a.sumsq = 0.0;
  } else {
    // (sampled - sampled_expected) / sampled_expected with sampled_expected = E_j * fps (:502-504),
    // accumulated relative to the first iteration's value (exact zeros when all agree)
    const double d = sv * winv_j * rf - 1.0;
    if (j == 1) a.pivot = d;
    const double dd = d - a.pivot;
    a.sum += dd;
    a.sumsq += dd * dd;
  }
}
__device__ __forceinline__ void dn_cell_finish(const DnCell& a, double E0, double fpsv, int B, bool empty,
                                               double& devv, double& zv) {
  const double expected = E0 * fpsv;                                  // :488
  const double obs_dev = (a.obs_v - expected) / expected;             // :489
  const double Bd = (double)B;
  const double mshift = a.sum / Bd;
  const double mean = a.pivot + mshift;                               // :511
  const double var = (a.sumsq - a.sum * mshift) / (Bd - 1.0);
  const double sd = sqrt(var > 0.0 ? var : (var == var ? 0.0 : var)); // :512
  devv = obs_dev - mean;                                              // :514
  zv = devv / sd;                                                     // :515
  if (empty || expected < 1.0) devv = zv = cv_na_real();              // :458-461, :509, :517-520
}

// raster: bands of `band` groups; inside a band the cell units are the slow index so that the
// concurrently running tiles cover about band groups x (#CTAs / band) cell units (L2 reuse)
__device__ __forceinline__ void dn_decode_tile(int band, int n_groups, int n_units, int64_t tile, int& unit, int& g) {
  const int64_t per_band = (int64_t)band * n_units;
  const int bnd = (int)(tile / per_band);
  const int64_t r = tile - (int64_t)bnd * per_band;
  const int g0 = bnd * band;
  const int gw = min(band, n_groups - g0);
  unit = (int)(r / gw);
  g = g0 + (int)(r % gw);
}

// 1 / E[k][j] of the group's annotations -> shared memory (epilogue warps only: bar 1, 128 threads)
// NT = number of epilogue threads of the CTA (128: four warps; 256: eight warps, e2m1 kernel)
template <int NT>
__device__ __forceinline__ void dn_epi_bar() {
  if (NT == 128) asm volatile("bar.sync 1, 128;" ::: "memory");
  else asm volatile("bar.sync 1, 256;" ::: "memory");
}
template <int MODE, int NT = 128>
__device__ __forceinline__ void dn_epi_stage_tables(const DenseParams& prm, int k0, int nk, double* s_winv, int et) {
  const int I = prm.B + 1;
  dn_epi_bar<NT>();
  if (MODE == DN_EPI_FUSED32) {
    uint32_t* s2 = reinterpret_cast<uint32_t*>(s_winv);
    for (int i = et; i < nk * I; i += NT) s2[i] = prm.winv_fx[(int64_t)k0 * I + i];
  } else {
    for (int i = et; i < nk * I; i += NT) s_winv[i] = 1.0 / prm.etab[(int64_t)k0 * I + i];
  }
  dn_epi_bar<NT>();
}

// One epilogue thread = one cell (one TMEM lane): streams the 256 accumulator columns of its lane.
// FUSED: deviations / z / variability partials straight from TMEM.  SUMS: the integer sums of this
// digit plane are added (shifted) into the int64 scratch and finished by k_dense_finalize_sums.
// F32ACC: the accumulator holds the exact integer sums as f32 (kind::mxf4 path; sums < 2^24).
template <int MODE, bool F32ACC = false>
__device__ __forceinline__ void dn_epi_consume(const DenseParams& prm, int cb_local, int k0, int nk, uint32_t taddr,
                                               const double* s_winv, double* s_wred, int q4, int lane) {
  const int I = prm.B + 1;
  const int c = (prm.cb0 + cb_local) * DN_BLOCK_M + q4 * 32 + lane;
  const bool live = c < prm.c_end;
  const double fpsv = live ? prm.fps[c] : 1.0;
  const double rf = 1.0 / fpsv;
  int a = 0, j = 0;
  DnCell cell = {0.0, 0.0, 0.0, 0.0};
  // fixed-point form: per-annotation state between the two sweeps (indexed by the running annotation:
  // thread-local memory, written / read once per annotation)
  unsigned long long f_sumi = 0ull;
  unsigned long long f_tot[MODE == DN_EPI_FUSED32 ? DN_MAX_G : 1];
  int f_obs[MODE == DN_EPI_FUSED32 ? DN_MAX_G : 1];
  const double inv_b = 1.0 / (double)prm.B, inv_bm1 = 1.0 / (double)(prm.B - 1);
  const int ncols = nk * I;
  for (int col0 = 0; col0 < ncols; col0 += 32) {
    uint32_t r[32];
    tmem_ld32(taddr + (uint32_t)col0, r);
    tmem_ld_wait();
    if (F32ACC) {
#pragma unroll
      for (int u = 0; u < 32; ++u) r[u] = (uint32_t)__float2int_rn(__uint_as_float(r[u]));
    }
#pragma unroll
    for (int u = 0; u < 32; ++u) {
      if (col0 + u < ncols) {
        const int k = k0 + a;
        if (MODE == DN_EPI_SUMS) {
          if (live) {
            long long* dst = prm.sums + ((int64_t)k * I + j) * prm.sums_ld + (c - prm.sums_c0);
            const long long v = (long long)(int)r[u] << prm.sums_shift;
            *dst = prm.sums_acc ? *dst + v : v;
          }
          if (++j == I) { j = 0; ++a; }
        } else {
          if (live) {
            if (j == 0) { if (prm.obs) prm.obs[(int64_t)k * prm.N + c] = (long long)(int)r[u]; }
            else if (prm.samp) prm.samp[((int64_t)k * prm.B + (j - 1)) * prm.N + c] = (long long)(int)r[u];
          }
          if (MODE == DN_EPI_FUSED32) {
            // first sweep: y_j = s_j * W_j (exact, < 2^56); their sum gives the mean over iterations exactly
            if (j == 0) f_obs[a] = (int)r[u];
            else f_sumi += (unsigned long long)r[u] * (unsigned long long)reinterpret_cast<const uint32_t*>(s_winv)[a * I + j];
          } else {
            dn_cell_add(cell, j, (double)(int)r[u], s_winv[a * I + j], rf);
          }
          if (++j == I) {
            const bool empty = prm.annoptr[k + 1] == prm.annoptr[k];
            double devv, zv;
            if (MODE == DN_EPI_FUSED32) {
              f_tot[a] = f_sumi;                         // finished by the second sweep below
              f_sumi = 0ull;
              j = 0;
              ++a;
              continue;
            }
            dn_cell_finish(cell, prm.etab[(int64_t)k * I], fpsv, prm.B, empty, devv, zv);
            if (live) {
              prm.dev_rm[(int64_t)k * prm.N + c] = devv;
              prm.z_rm[(int64_t)k * prm.N + c] = zv;
            }
            // variability partial of (k, cell block): this warp's 32 cells, then the 4 epilogue warps
            dn_warp_variability(zv, live, lane, s_wred + (a * 4 + q4) * 4);
            j = 0;
            ++a;
          }
        }
      }
    }
  }
  if (MODE != DN_EPI_FUSED32) return;
  // Second sweep of the fixed-point form (the accumulator is still in tensor memory): squared
  // differences to a centre c within a few units of the integer mean sum_j y_j / B of this
  // (annotation, cell), in fp32.  Every term is >= 0 and the residual s1 = sum_j (y_j - c), an exact
  // integer of magnitude <= B, enters once -- nothing cancels whatever the background draws look like
  // (a pivot on the first iteration loses digits when that iteration is an outlier: one fragment
  // against 49 zeros amplified the fp32 rounding ~50x).  No fp64 inside the column loop: predicated-
  // off fp64 instructions still issue, and fp64 issue is throttled while the tensor pipe runs.
  const unsigned long long magic = ~0ull / (unsigned long long)prm.B;        // floor((2^64 - 1) / B)
  a = 0;
  j = 0;
  unsigned long long ctr = __umul64hi(f_tot[0], magic);                      // floor(tot / B) or one less
  float f_sumsq = 0.f;
  unsigned redo = 0u, redo_mine = 0u;          // annotations for the fp64 sweep: any lane of the warp / this cell
  for (int col0 = 0; col0 < ncols; col0 += 32) {
    uint32_t r[32];
    tmem_ld32(taddr + (uint32_t)col0, r);
    tmem_ld_wait();
    if (F32ACC) {
#pragma unroll
      for (int u = 0; u < 32; ++u) r[u] = (uint32_t)__float2int_rn(__uint_as_float(r[u]));
    }
#pragma unroll
    for (int u = 0; u < 32; ++u) {
      if (col0 + u < ncols) {
        if (j != 0) {
          const unsigned long long prod =
              (unsigned long long)r[u] * (unsigned long long)reinterpret_cast<const uint32_t*>(s_winv)[a * I + j];
          const float df = (float)(long long)(prod - ctr);
          f_sumsq = fmaf(df, df, f_sumsq);
        }
        if (++j == I) {
          // FP64 issues ~25x slower while the tensor pipe runs, so this tail (once per annotation and
          // cell) uses a dozen fp64 operations: reciprocals come from tables, the square root is an
          // fp32 rsqrt + one Newton step (the variance is an fp32 sum anyway)
          const int k = k0 + a;
          const bool empty = prm.annoptr[k + 1] == prm.annoptr[k];
          const unsigned long long tot = f_tot[a];
          const double4 cst = *reinterpret_cast<const double4*>(prm.wscale + 4 * (int64_t)k);   // 2^-F / B, 2^F, 1 / E0, E0
          const double expected = cst.w * fpsv;                                                 // :488
          const double obs_dev = (double)f_obs[a] * (rf * cst.z) - 1.0;                         // :489
          const double mean = (double)tot * (rf * cst.x) - 1.0;                                 // :511
          const float s1 = (float)(long long)(tot - (unsigned long long)prm.B * ctr);           // 0 <= s1 < 2 B
          float var = (f_sumsq - s1 * s1 * (float)inv_b) * (float)inv_bm1;                      // :512, scaled units
          var = var > 0.f ? var : (var == var ? 0.f : var);
          // The table entries W_j carry 32 bits, i.e. each y_j / c is off by up to ~2^-31: once the
          // spread of the background deviations falls below ~5e-4 (only under-dispersed counts do that
          // with sums < 2^24; Poisson spread there is >= 2.4e-4 ... 1) that shows in z.  Such
          // (annotation, cell) pairs are redone in fp64 below.
          const float fc = (float)ctr;
          const bool thin = live && !empty && var < 2.5e-7f * fc * fc;
          if (thin) redo_mine |= 1u << a;
          if (__any_sync(0xffffffffu, thin)) redo |= 1u << a;
          float rs = rsqrtf(var);
          rs = rs * (1.5f - 0.5f * ((var * rs) * rs));                                          // inf / nan pass through
          if (var == 0.f) rs = __int_as_float(0x7f800000);
          double devv = obs_dev - mean;                                                         // :514
          double zv = devv * ((fpsv * cst.y) * (double)rs);                                     // :515
          if (empty || expected < 1.0) devv = zv = cv_na_real();
          if (live) {
            prm.dev_rm[(int64_t)k * prm.N + c] = devv;
            prm.z_rm[(int64_t)k * prm.N + c] = zv;
          }
          dn_warp_variability_f32(zv, live, lane, s_wred + (a * 4 + q4) * 4);
          j = 0;
          ++a;
          ctr = __umul64hi(f_tot[a < nk ? a : 0], magic);
          f_sumsq = 0.f;
        }
      }
    }
  }
  if (!redo) return;
  // Third sweep, rare: the flagged (annotation, cell) pairs again, in the fp64 form of DN_EPI_FUSED
  // (reciprocals of the expectation sums straight from the table in global memory).  The sweep is
  // warp-wide (tensor-memory loads are), but only a flagged cell takes the fp64 value: what a cell
  // gets never depends on which other cells share its warp, so sharding the cells differently
  // leaves every result bit for bit the same.
  a = 0;
  j = 0;
  for (int col0 = 0; col0 < ncols; col0 += 32) {
    uint32_t r[32];
    tmem_ld32(taddr + (uint32_t)col0, r);
    tmem_ld_wait();
    if (F32ACC) {
#pragma unroll
      for (int u = 0; u < 32; ++u) r[u] = (uint32_t)__float2int_rn(__uint_as_float(r[u]));
    }
#pragma unroll 1
    for (int u = 0; u < 32; ++u) {
      if (col0 + u >= ncols) break;
      const int k = k0 + a;
      if ((redo >> a) & 1u) dn_cell_add(cell, j, (double)(int)r[u], 1.0 / prm.etab[(int64_t)k * I + j], rf);
      if (++j == I) {
        if ((redo >> a) & 1u) {
          const bool empty = prm.annoptr[k + 1] == prm.annoptr[k];
          double devv, zv;
          dn_cell_finish(cell, prm.etab[(int64_t)k * I], fpsv, prm.B, empty, devv, zv);
          if (live) {
            if ((redo_mine >> a) & 1u) {
              prm.dev_rm[(int64_t)k * prm.N + c] = devv;
              prm.z_rm[(int64_t)k * prm.N + c] = zv;
            } else {
              zv = prm.z_rm[(int64_t)k * prm.N + c];           // this thread's own value of the second sweep
            }
          }
          dn_warp_variability(zv, live, lane, s_wred + (a * 4 + q4) * 4);
        }
        j = 0;
        ++a;
      }
    }
  }
}

// merge the 4 epilogue warps' variability partials of every annotation of the group
template <int NT = 128>
__device__ __forceinline__ void dn_epi_merge(const DenseParams& prm, int cb_local, int k0, int nk,
                                             const double* s_wred, int et) {
  dn_epi_bar<NT>();
  if (et < nk) {
    DWelford tot = {0.0, 0.0, 0.0};
    double nf = 0.0;
    for (int w4 = 0; w4 < 4; ++w4) {
      const double* w = s_wred + (et * 4 + w4) * 4;
      DWelford b = {w[0], w[1], w[2]};
      tot = dw_merge(tot, b);
      nf += w[3];
    }
    double* vp = prm.varpart + ((int64_t)(k0 + et) * prm.ncb_total + prm.cb0 + cb_local) * 4;
    vp[0] = tot.n; vp[1] = tot.mean; vp[2] = tot.m2; vp[3] = nf;
  }
}

