// K5a: the dense tensor-core path of compute_deviations_core.
//
// The observed sums and the B background sums of R/compute_deviations.R:486,501 are one integer
// GEMM once the background gather is folded into the annotation operand:
//
//     S[c, (k, i)] = sum_q  Cd[c, q] * M[(k, i), q]
//       Cd[c, q]      = counts of peak q in cell c                      (u8, cells x peaks, K-major)
//       M[(k,0), q]   = multiplicity of q in annotation k               (tf_vec, :480-484)
//       M[(k,i), q]   = #{p in set_k : bg[p, i] = q}                    (sample_mat, :492-499)
//
// M depends only on the annotations and the background matrix, so it is built once per call
// (a scatter of (B+1) * nnzA increments) and the inner loop is a plain u8 x u8 -> s32 GEMM on
// tcgen05 with no gather: TMA (SWIZZLE_128B) stages both operands, one elected thread issues
// tcgen05.mma.kind::i8, accumulators live in TMEM.  The stacked rows of M are grouped so that
// one 256-column accumulator tile holds floor(256 / (B+1)) whole annotations x all B+1
// iterations (5 x 51 = 255 for the default 50 iterations): the epilogue thread that owns a cell
// (one TMEM lane) therefore sees, in its own registers, every sum it needs for that cell, and
// the fused fp64 epilogue (expected, raw deviations, mean / sd over iterations, bias-corrected
// deviation, z, NA mask, variability partials; :488-520) never writes an intermediate to HBM.
//
// Warp roles (256 threads, 1 CTA / SM, persistent over tiles):
//   warp 0  TMA producer          warp 1  MMA issuer        warp 2  TMEM allocator
//   warps 4-7  epilogue (TMEM lane quadrant = warp % 4)
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue, two
// 256-column accumulator stages so the epilogue of tile n overlaps the main loop of tile n+1).
#include <cuda.h>

#include <algorithm>

#include "cv_internal.cuh"

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
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
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

// ---- the GEMM + fused epilogue -----------------------------------------------------------------------
struct DenseParams {
  int N, K, B, G;            // cells, annotations, iterations, annotations per 256-row group
  int n_cell_blocks, n_groups, n_kblocks;
  int64_t n_tiles;
  int band;                  // groups per raster band
  const int64_t* annoptr;
  const double* etab;        // [K][B+1]
  const double* fps;         // [N]
  double* dev_rm;            // [K][N]
  double* z_rm;
  double* varpart;           // [K][n_cell_blocks][4]
  long long* obs;            // [K][N] or null
  long long* samp;           // [K][B][N] or null
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

__device__ __forceinline__ void dn_decode_tile(const DenseParams& p, int64_t tile, int& cb, int& g) {
  // raster: bands of `band` groups; inside a band cell blocks are the slow index so that the
  // ~148 concurrently running tiles cover about band groups x 148/band cell blocks (L2 reuse)
  const int64_t per_band = (int64_t)p.band * p.n_cell_blocks;
  const int bnd = (int)(tile / per_band);
  const int64_t r = tile - (int64_t)bnd * per_band;
  const int g0 = bnd * p.band;
  const int gw = min(p.band, p.n_groups - g0);
  cb = (int)(r / gw);
  g = g0 + (int)(r % gw);
}

__global__ void __launch_bounds__(DN_THREADS, 1)
k_dense_contract(const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_m,
                 const DenseParams prm) {
  extern __shared__ unsigned char dn_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dn_smem_raw) + 1023) &
                                                         ~(uintptr_t)1023);
  unsigned char* smem_a = smem;                                   // [STAGES][16 KB]
  unsigned char* smem_b = smem + DN_STAGES * DN_A_BYTES;          // [STAGES][32 KB]
  double* s_winv = reinterpret_cast<double*>(smem + DN_STAGES * DN_STAGE_BYTES);            // [256]
  double* s_wred = s_winv + 256;                                                            // [4][4][...]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DN_STAGES * DN_STAGE_BYTES + 4096);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + DN_STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * DN_STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * DN_STAGES + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * DN_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int I = prm.B + 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_m) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < DN_STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar + s, 1); mbar_init(tempty_bar + s, 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int64_t tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x) {
        int cb, g;
        dn_decode_tile(prm, tile, cb, g);
        for (int kb = 0; kb < prm.n_kblocks; ++kb) {
          mbar_wait(empty_bar + s, ph ^ 1u);
          mbar_expect_tx(full_bar + s, DN_STAGE_BYTES);
          tma_load_2d(&map_c, full_bar + s, smem_a + s * DN_A_BYTES, kb * DN_BLOCK_K, cb * DN_BLOCK_M);
          tma_load_2d(&map_m, full_bar + s, smem_b + s * DN_B_BYTES, kb * DN_BLOCK_K, g * DN_BLOCK_N);
          if (++s == DN_STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_u8(DN_BLOCK_M, DN_BLOCK_N);
      int s = 0;
      uint32_t ph = 0;
      int64_t it = 0;
      for (int64_t tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x, ++it) {
        const int as = (int)(it & 1);
        mbar_wait(tempty_bar + as, (uint32_t)((it >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * DN_BLOCK_N);
        for (int kb = 0; kb < prm.n_kblocks; ++kb) {
          mbar_wait(full_bar + s, ph);
          tc_fence_after();
          const uint64_t da = make_smem_desc(smem_u32(smem_a + s * DN_A_BYTES));
          const uint64_t db = make_smem_desc(smem_u32(smem_b + s * DN_B_BYTES));
#pragma unroll
          for (int k = 0; k < DN_BLOCK_K / DN_UMMA_K; ++k)
            umma_i8(tmem_d, da + (uint64_t)(k * DN_UMMA_K / 16), db + (uint64_t)(k * DN_UMMA_K / 16), idesc,
                    (kb | k) ? 1u : 0u);
          tc_commit(empty_bar + s);            // frees the smem stage when these MMAs retire
          if (++s == DN_STAGES) { s = 0; ph ^= 1u; }
        }
        tc_commit(tfull_bar + as);             // accumulator tile complete
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: one thread = one cell (TMEM lane) =====
    const int q4 = warp & 3;                   // TMEM lane quadrant this warp may read
    const int et = threadIdx.x - 128;          // 0..127
    const unsigned full = 0xffffffffu;
    int64_t it = 0;
    for (int64_t tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x, ++it) {
      int cb, g;
      dn_decode_tile(prm, tile, cb, g);
      const int as = (int)(it & 1);
      const int k0 = g * prm.G;
      const int nk = min(prm.G, prm.K - k0);
      // 1 / E[k][j] for the annotations of this group
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int i = et; i < nk * I; i += 128) s_winv[i] = 1.0 / prm.etab[(int64_t)k0 * I + i];
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int c = cb * DN_BLOCK_M + q4 * 32 + lane;
      const bool live = c < prm.N;
      const double fpsv = live ? prm.fps[c] : 1.0;
      const double rf = 1.0 / fpsv;

      mbar_wait(tfull_bar + as, (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(as * DN_BLOCK_N);

      // stream the 256 columns; (a, j) walks (annotation in group, iteration)
      int a = 0, j = 0;
      double obs_v = 0.0, pivot = 0.0, sum = 0.0, sumsq = 0.0;
      const int ncols = nk * I;
      for (int col0 = 0; col0 < ncols; col0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)col0, r);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          if (col0 + u < ncols) {
            const double sv = (double)(int)r[u];
            const int k = k0 + a;
            if (j == 0) {
              obs_v = sv;
              sum = 0.0; sumsq = 0.0;
              if (prm.obs && live) prm.obs[(int64_t)k * prm.N + c] = (long long)(int)r[u];
            } else {
              // (sampled - sampled_expected) / sampled_expected with sampled_expected = E_j * fps
              // accumulated relative to the first iteration's value (exact zeros when all agree)
              const double d = sv * s_winv[a * I + j] * rf - 1.0;
              if (j == 1) pivot = d;
              const double dd = d - pivot;
              sum += dd;
              sumsq += dd * dd;
              if (prm.samp && live) prm.samp[((int64_t)k * prm.B + (j - 1)) * prm.N + c] = (long long)(int)r[u];
            }
            if (++j == I) {
              // ---- finish annotation k for this cell (R/compute_deviations.R:488-520) ----
              const bool empty = prm.annoptr[k + 1] == prm.annoptr[k];
              const double E0 = prm.etab[(int64_t)k * I];
              const double expected = E0 * fpsv;
              const double obs_dev = (obs_v - expected) / expected;
              const double Bd = (double)prm.B;
              const double mshift = sum / Bd;
              const double mean = pivot + mshift;
              const double var = (sumsq - sum * mshift) / (Bd - 1.0);
              const double sd = sqrt(var > 0.0 ? var : (var == var ? 0.0 : var));
              double devv = obs_dev - mean, zv = devv / sd;
              if (empty || expected < 1.0) devv = zv = cv_na_real();
              DWelford wf = {0.0, 0.0, 0.0};
              double nonfinite = 0.0;
              if (live) {
                prm.dev_rm[(int64_t)k * prm.N + c] = devv;
                prm.z_rm[(int64_t)k * prm.N + c] = zv;
                if (isfinite(zv)) { wf.n = 1.0; wf.mean = zv; } else nonfinite = 1.0;
              }
              // variability partial of (k, cell block): warp tree, then the 4 epilogue warps
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                DWelford b;
                b.n = __shfl_xor_sync(full, wf.n, o);
                b.mean = __shfl_xor_sync(full, wf.mean, o);
                b.m2 = __shfl_xor_sync(full, wf.m2, o);
                wf = (lane & o) ? dw_merge(b, wf) : dw_merge(wf, b);
                nonfinite += __shfl_xor_sync(full, nonfinite, o);
              }
              if (lane == 0) {
                double* w = s_wred + (a * 4 + q4) * 4;
                w[0] = wf.n; w[1] = wf.mean; w[2] = wf.m2; w[3] = nonfinite;
              }
              j = 0;
              ++a;
            }
          }
        }
      }
      // every TMEM column of this stage has been read: hand the accumulator back to the MMA warp
      tc_fence_before();
      mbar_arrive(tempty_bar + as);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et < nk) {
        DWelford tot = {0.0, 0.0, 0.0};
        double nf = 0.0;
        for (int w4 = 0; w4 < 4; ++w4) {
          const double* w = s_wred + (et * 4 + w4) * 4;
          DWelford b = {w[0], w[1], w[2]};
          tot = dw_merge(tot, b);
          nf += w[3];
        }
        double* vp = prm.varpart + ((int64_t)(k0 + et) * prm.n_cell_blocks + cb) * 4;
        vp[0] = tot.n; vp[1] = tot.mean; vp[2] = tot.m2; vp[3] = nf;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- host side --------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map_u8(cv_handle* h, CUtensorMap* map, void* base, uint64_t rows, uint64_t row_bytes,
                       uint32_t box_rows) {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || !p) CV_FAIL(h, CV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    fn = (PFN_encodeTiled)p;
  }
  cuuint64_t gdim[2] = {row_bytes, rows};
  cuuint64_t gstr[1] = {row_bytes};
  cuuint32_t box[2] = {DN_BLOCK_K, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) CV_FAIL(h, CV_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return CV_OK;
}

// can the dense path represent this problem exactly?  (reason filled when it cannot)
int cv_dense_feasible(cv_handle* h, std::string* reason) {
  const int I = h->B + 1;
  if (I > DN_BLOCK_N) { *reason = "niterations + 1 > 256"; return 0; }
  if (h->max_val > 255) { *reason = "counts above 255 (u8 operand)"; return 0; }
  const int G = std::min(DN_BLOCK_N / I, DN_MAX_G);
  const int64_t Ppad = ((h->P + DN_BLOCK_K - 1) / DN_BLOCK_K) * DN_BLOCK_K;
  const int64_t n_groups = (h->K + G - 1) / G;
  const int64_t ncb = (h->N + DN_BLOCK_M - 1) / DN_BLOCK_M;
  const double bytes = (double)Ppad * ((double)n_groups * DN_BLOCK_N + (double)ncb * DN_BLOCK_M);
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  const double avail = (double)free_b + (double)h->dn_M.cap + (double)h->dn_C.cap;
  if (bytes > 0.85 * avail) { *reason = "dense operands do not fit in device memory"; return 0; }
  return 1;
}

double cv_dense_cost_ms(cv_handle* h) {
  const int I = h->B + 1;
  const int G = std::min(DN_BLOCK_N / I, DN_MAX_G);
  const double Ppad = (double)(((h->P + DN_BLOCK_K - 1) / DN_BLOCK_K) * DN_BLOCK_K);
  const double rows = (double)((h->K + G - 1) / G) * DN_BLOCK_N;
  const double cells = (double)((h->N + DN_BLOCK_M - 1) / DN_BLOCK_M) * DN_BLOCK_M;
  return 2.0 * Ppad * rows * cells / 2.0e15 * 1e3 + 0.3;     // ~2 POP/s sustained + fixed cost
}

// Builds the operands, runs the GEMM + fused epilogue.  On return dev_rm / z_rm / varpart (with
// *T_out cell blocks per annotation) / obs / samp are filled exactly like k_contract fills them.
// Returns CV_ERR_UNSUPPORTED (with *soft_fail = 1) when the exact-multiplicity check fails so the
// caller can fall back to the sparse kernel.
int cv_dense_contract(cv_handle* h, int keep_sums, int* T_out, int* soft_fail) {
  *soft_fail = 0;
  const int64_t P = h->P, N = h->N, K = h->K;
  const int B = h->B, I = B + 1;
  const int G = std::min(DN_BLOCK_N / I, DN_MAX_G);
  const int64_t Ppad = ((P + DN_BLOCK_K - 1) / DN_BLOCK_K) * DN_BLOCK_K;
  const int n_groups = (int)((K + G - 1) / G);
  const int ncb = (int)((N + DN_BLOCK_M - 1) / DN_BLOCK_M);
  const int64_t rowsM = (int64_t)n_groups * DN_BLOCK_N, rowsC = (int64_t)ncb * DN_BLOCK_M;

  // ---- operands, per-annotation tables, exactness bounds (cv_dense_build.cu) -------------------
  CV_TRY(cv_dense_build_operands(h, G, Ppad, rowsM, rowsC, soft_fail));

  // ---- GEMM + fused epilogue --------------------------------------------------------------------------
  CUtensorMap map_c, map_m;
  CV_TRY(make_map_u8(h, &map_c, h->dn_C.p, (uint64_t)rowsC, (uint64_t)Ppad, DN_BLOCK_M));
  CV_TRY(make_map_u8(h, &map_m, h->dn_M.p, (uint64_t)rowsM, (uint64_t)Ppad, DN_BLOCK_N));
  CV_TRY(cv_ensure(h, h->varpart, (size_t)K * ncb * 4 * 8));
  DenseParams prm;
  prm.N = (int)N; prm.K = (int)K; prm.B = B; prm.G = G;
  prm.n_cell_blocks = ncb; prm.n_groups = n_groups; prm.n_kblocks = (int)(Ppad / DN_BLOCK_K);
  prm.n_tiles = (int64_t)ncb * n_groups;
  prm.band = 12;
  prm.annoptr = h->annoptr.as<int64_t>();
  prm.etab = h->etab.as<double>();
  prm.fps = h->fps.as<double>();
  prm.dev_rm = h->dev_rm.as<double>();
  prm.z_rm = h->z_rm.as<double>();
  prm.varpart = h->varpart.as<double>();
  prm.obs = keep_sums ? h->obs.as<long long>() : nullptr;
  prm.samp = keep_sums ? h->samp.as<long long>() : nullptr;
  CV_CUDA(h, cudaFuncSetAttribute(k_dense_contract, cudaFuncAttributeMaxDynamicSharedMemorySize, DN_SMEM_BYTES));
  const int grid = (int)std::min<int64_t>(prm.n_tiles, h->num_sms);
  CV_CUDA(h, cudaEventRecord(h->ev[3], h->stream));
  CV_LAUNCH(h, k_dense_contract, grid, DN_THREADS, DN_SMEM_BYTES, map_c, map_m, prm);
  CV_CUDA(h, cudaEventRecord(h->ev[4], h->stream));
  *T_out = ncb;
  h->st.contract_updates = 0;
  h->st.contract_bytes = 0;
  h->dn_flops = 2.0 * (double)rowsM * (double)rowsC * (double)Ppad;
  return CV_OK;
}
