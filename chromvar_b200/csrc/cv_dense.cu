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
// Two kernels share the epilogue: k_dense_contract (one CTA per 128-cell x 256-row tile) and
// k_dense_contract_pair (tcgen05 cta_group::2: a CTA pair per 256-cell x 256-row tile; default).
// Counts above 255 run as u8 digit planes (one GEMM per plane, int64 sums combined in HBM scratch
// and finished by k_dense_finalize_sums); cells are processed in chunks so the densified operand
// never exceeds a fixed budget and host copies can overlap the GEMMs.
//
// Warp roles (256 threads, 1 CTA / SM, persistent over tiles):
//   warp 0  TMA producer          warp 1  MMA issuer        warp 2  TMEM allocator
//   warps 4-7  epilogue (TMEM lane quadrant = warp % 4)
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue, two
// 256-column accumulator stages so the epilogue of tile n overlaps the main loop of tile n+1).
#include <cuda.h>

#include <algorithm>

#include "cv_internal.cuh"

#include "cv_dense_dev.cuh"

// ---- single-CTA kernel: tile = 128 cells x 256 stacked rows ----------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(DN_THREADS, 1)
k_dense_contract(const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_m,
                 const DenseParams prm) {
  extern __shared__ unsigned char dn_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dn_smem_raw) + 1023) &
                                                         ~(uintptr_t)1023);
  unsigned char* smem_a = smem;                                   // [STAGES][16 KB]
  unsigned char* smem_b = smem + DN_STAGES * DN_A_BYTES;          // [STAGES][32 KB]
  double* s_winv = reinterpret_cast<double*>(smem + DN_STAGES * DN_STAGE_BYTES);            // [256]
  double* s_wred = s_winv + 256;                                                            // [G][4][4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DN_STAGES * DN_STAGE_BYTES + 4096);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + DN_STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * DN_STAGES;   // [2]
  uint64_t* tempty_bar = bars + 2 * DN_STAGES + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * DN_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

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
        dn_decode_tile(prm.band, prm.n_groups, prm.n_cell_blocks, tile, cb, g);
        for (int kb = 0; kb < prm.n_kblocks; ++kb) {
          mbar_wait_relaxed(empty_bar + s, ph ^ 1u, 32);
          if (prm.debug_no_tma) { mbar_arrive(full_bar + s); if (++s == DN_STAGES) { s = 0; ph ^= 1u; } continue; }
          mbar_expect_tx(full_bar + s, DN_STAGE_BYTES);
          tma_load_2d(&map_c, full_bar + s, smem_a + s * DN_A_BYTES, kb * DN_BLOCK_K, cb * DN_BLOCK_M);
          tma_load_2d(&map_m, full_bar + s, smem_b + s * DN_B_BYTES, kb * DN_BLOCK_K, g * DN_BLOCK_N);
          if (++s == DN_STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the pipeline, one elected lane issues =====
    {
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
          if (prm.debug_no_tma > 1 && prm.debug_no_tma < 7) {
            // issue-pattern experiments (garbage results; bench.py --debug-no-tma V)
            const int v = prm.debug_no_tma;
            if (elect_one()) {
            if (v == 2) {          // alternate two accumulators: no back-to-back dependent MMAs
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_i8(tmem_base + (uint32_t)((k & 1) * DN_BLOCK_N), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1u);
            } else if (v == 3) {   // two independent N = 128 halves, alternating
              constexpr uint32_t idesc_h = make_idesc_u8(DN_BLOCK_M, DN_BLOCK_N / 2);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_i8(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc_h, 1u);
                umma_i8(tmem_d + 128u, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2 + 1024), idesc_h, 1u);
              }
            } else if (v == 4) {   // overwrite instead of accumulate: no read of D
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_i8(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 0u);
            } else if (v == 6) {   // A operand from tensor memory (garbage columns of the other stage)
              const uint32_t ta = tmem_base + (uint32_t)((as ^ 1) * DN_BLOCK_N);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_i8_ts(tmem_d, ta + (uint32_t)(k * 8), db + (uint64_t)(k * 2), idesc, 1u);
            } else {               // v == 5: same operand bytes every time (k offset 0): smem bank / swizzle effects
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_i8(tmem_d, da, db, idesc, 1u);
            }
            tc_commit(empty_bar + s);
            }
            __syncwarp();
            if (++s == DN_STAGES) { s = 0; ph ^= 1u; }
            continue;
          }
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < DN_BLOCK_K / DN_UMMA_K; ++k)
              umma_i8(tmem_d, da + (uint64_t)(k * DN_UMMA_K / 16), db + (uint64_t)(k * DN_UMMA_K / 16), idesc,
                      (kb | k) ? 1u : 0u);
            tc_commit(empty_bar + s);          // frees the smem stage when these MMAs retire
          }
          __syncwarp();
          if (++s == DN_STAGES) { s = 0; ph ^= 1u; }
        }
        if (elect_one()) tc_commit(tfull_bar + as);   // accumulator tile complete
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: one thread = one cell (TMEM lane) =====
    const int q4 = warp & 3;                   // TMEM lane quadrant this warp may read
    const int et = threadIdx.x - 128;          // 0..127
    int64_t it = 0;
    for (int64_t tile = blockIdx.x; tile < prm.n_tiles; tile += gridDim.x, ++it) {
      int cb, g;
      dn_decode_tile(prm.band, prm.n_groups, prm.n_cell_blocks, tile, cb, g);
      const int as = (int)(it & 1);
      const int k0 = g * prm.G, nk = min(prm.G, prm.K - k0);
      if (MODE != DN_EPI_SUMS) dn_epi_stage_tables<MODE>(prm, k0, nk, s_winv, et);
      mbar_wait_relaxed(tfull_bar + as, (uint32_t)((it >> 1) & 1), 256);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(as * DN_BLOCK_N);
      if (prm.debug_no_tma == 7) {
        // experiment: no epilogue work at all
      } else if (prm.debug_no_tma == 8) {
        // experiment: TMEM loads only
        uint32_t r[32];
        uint32_t acc = 0;
        for (int col0 = 0; col0 < 256; col0 += 32) {
          tmem_ld32(taddr + (uint32_t)col0, r);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 32; ++u) acc ^= r[u];
        }
        if (acc == 0xdeadbeefu) prm.z_rm[0] = 0.0;
      } else
      dn_epi_consume<MODE>(prm, cb, k0, nk, taddr, s_winv, s_wred, q4, lane);
      // every TMEM column of this stage has been read: hand the accumulator back to the MMA warp
      tc_fence_before();
      mbar_arrive(tempty_bar + as);
      if (MODE != DN_EPI_SUMS) dn_epi_merge(prm, cb, k0, nk, s_wred, et);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- CTA-pair kernel: tile = 256 cells x 256 stacked rows on two SMs (tcgen05 cta_group::2) -------------
// Each CTA of the pair stages its own 128 cells (A) and one half of the 256 stacked rows (B), so
// the L2 -> SM operand stream per SM drops by a third against the single-CTA tile and the ring is
// 6 stages of 32 KB.  The leader CTA (cluster rank 0) issues every MMA; its commits are multicast
// to the barriers of both CTAs.  Both CTAs run the same epilogue on their own TMEM lanes.
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(DN_THREADS, 1)
k_dense_contract_pair(const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_m,
                      const DenseParams prm) {
  extern __shared__ unsigned char dn_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dn_smem_raw) + 1023) &
                                                         ~(uintptr_t)1023);
  unsigned char* smem_a = smem;                                   // [STAGES2][16 KB] own cells
  unsigned char* smem_b = smem + DN_STAGES2 * DN_A_BYTES;         // [STAGES2][16 KB] own half of the rows
  double* s_winv = reinterpret_cast<double*>(smem + DN_STAGES2 * DN_STAGE2_BYTES);          // [256]
  double* s_wred = s_winv + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DN_STAGES2 * DN_STAGE2_BYTES + 4096);
  uint64_t* full_bar = bars;                     // [STAGES2]  (the leader's are the ones waited on)
  uint64_t* empty_bar = bars + DN_STAGES2;       // [STAGES2]
  uint64_t* tfull_bar = bars + 2 * DN_STAGES2;   // [2]
  uint64_t* tempty_bar = bars + 2 * DN_STAGES2 + 2;   // [2]  (leader's: 256 arrivals, both CTAs' epilogues)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * DN_STAGES2 + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_units = (prm.n_cell_blocks + 1) >> 1;               // 256-cell units of this launch

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_m) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < DN_STAGES2; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar + s, 1); mbar_init(tempty_bar + s, 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own cells + own half of the stacked rows =====
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int64_t tile = pair_id; tile < prm.n_tiles; tile += n_pairs) {
        int unit, g;
        dn_decode_tile(prm.band, prm.n_groups, n_units, tile, unit, g);
        const int yc = (2 * unit + (int)rank) * DN_BLOCK_M;
        const int ym = g * DN_BLOCK_N + (int)rank * (DN_BLOCK_N / 2);
        for (int kb = 0; kb < prm.n_kblocks; ++kb) {
          mbar_wait_relaxed(empty_bar + s, ph ^ 1u, 32);
          if (prm.debug_no_tma) {
            if (rank == 0) mbar_arrive(full_bar + s);
            if (++s == DN_STAGES2) { s = 0; ph ^= 1u; }
            continue;
          }
          if (rank == 0) mbar_expect_tx(full_bar + s, 2 * DN_STAGE2_BYTES);
          const uint32_t fb = mapa_u32(smem_u32(full_bar + s), 0);
          tma_load_2d_pair(&map_c, fb, smem_a + s * DN_A_BYTES, kb * DN_BLOCK_K, yc);
          tma_load_2d_pair(&map_m, fb, smem_b + s * DN_A_BYTES, kb * DN_BLOCK_K, ym);
          if (++s == DN_STAGES2) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_u8(2 * DN_BLOCK_M, DN_BLOCK_N);
      int s = 0;
      uint32_t ph = 0;
      int64_t it = 0;
      for (int64_t tile = pair_id; tile < prm.n_tiles; tile += n_pairs, ++it) {
        const int as = (int)(it & 1);
        mbar_wait(tempty_bar + as, (uint32_t)((it >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * DN_BLOCK_N);
        for (int kb = 0; kb < prm.n_kblocks; ++kb) {
          mbar_wait(full_bar + s, ph);
          tc_fence_after();
          const uint64_t da = make_smem_desc(smem_u32(smem_a + s * DN_A_BYTES));
          const uint64_t db = make_smem_desc(smem_u32(smem_b + s * DN_A_BYTES));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < DN_BLOCK_K / DN_UMMA_K; ++k)
              umma_i8_pair(tmem_d, da + (uint64_t)(k * DN_UMMA_K / 16), db + (uint64_t)(k * DN_UMMA_K / 16), idesc,
                           (kb | k) ? 1u : 0u);
            tc_commit_pair(empty_bar + s, 3);  // frees the stage in both CTAs when these MMAs retire
          }
          __syncwarp();
          if (++s == DN_STAGES2) { s = 0; ph ^= 1u; }
        }
        if (elect_one()) tc_commit_pair(tfull_bar + as, 3);   // accumulator tile complete: both epilogues may start
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs): one thread = one cell (TMEM lane of this CTA) =====
    const int q4 = warp & 3;
    const int et = threadIdx.x - 128;
    int64_t it = 0;
    for (int64_t tile = pair_id; tile < prm.n_tiles; tile += n_pairs, ++it) {
      int unit, g;
      dn_decode_tile(prm.band, prm.n_groups, n_units, tile, unit, g);
      const int cb = 2 * unit + (int)rank;
      const int as = (int)(it & 1);
      const int k0 = g * prm.G, nk = min(prm.G, prm.K - k0);
      if (MODE != DN_EPI_SUMS) dn_epi_stage_tables<MODE>(prm, k0, nk, s_winv, et);
      mbar_wait_relaxed(tfull_bar + as, (uint32_t)((it >> 1) & 1), 256);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(as * DN_BLOCK_N);
      dn_epi_consume<MODE>(prm, cb, k0, nk, taddr, s_winv, s_wred, q4, lane);
      tc_fence_before();
      mbar_arrive_cluster(mapa_u32(smem_u32(tempty_bar + as), 0));
      if (MODE != DN_EPI_SUMS && cb < prm.n_cell_blocks) dn_epi_merge(prm, cb, k0, nk, s_wred, et);
      else if (MODE != DN_EPI_SUMS) asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  }
  // no CTA may leave (or free TMEM) while its peer can still read its shared memory / signal it
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- digit-plane finish: int64 sums -> deviations / z / variability partials -------------------------------
// grid (cell blocks of the chunk, K); 128 threads = the cells of one block.  Same arithmetic as the
// fused epilogue (dn_cell_add / dn_cell_finish), so results are bit-identical to the one-plane path.
__global__ void __launch_bounds__(128)
k_dense_finalize_sums(const DenseParams prm) {
  __shared__ double s_w[4][4];
  const int I = prm.B + 1;
  const int k = blockIdx.y, cb_local = blockIdx.x;
  const int lane = threadIdx.x & 31, w4 = threadIdx.x >> 5;
  const int c = (prm.cb0 + cb_local) * DN_BLOCK_M + threadIdx.x;
  const bool live = c < prm.c_end;
  const double fpsv = live ? prm.fps[c] : 1.0;
  const double rf = 1.0 / fpsv;
  const double* et = prm.etab + (int64_t)k * I;
  DnCell cell = {0.0, 0.0, 0.0, 0.0};
  for (int j = 0; j < I; ++j) {
    const long long s = live ? prm.sums[((int64_t)k * I + j) * prm.sums_ld + (c - prm.sums_c0)] : 0ll;
    dn_cell_add(cell, j, (double)s, 1.0 / et[j], rf);
    if (live) {
      if (j == 0) { if (prm.obs) prm.obs[(int64_t)k * prm.N + c] = s; }
      else if (prm.samp) prm.samp[((int64_t)k * prm.B + (j - 1)) * prm.N + c] = s;
    }
  }
  const bool empty = prm.annoptr[k + 1] == prm.annoptr[k];
  double devv, zv;
  dn_cell_finish(cell, et[0], fpsv, prm.B, empty, devv, zv);
  if (live) {
    prm.dev_rm[(int64_t)k * prm.N + c] = devv;
    prm.z_rm[(int64_t)k * prm.N + c] = zv;
  }
  dn_warp_variability(zv, live, lane, &s_w[w4][0]);
  __syncthreads();
  if (threadIdx.x == 0) {
    DWelford tot = {0.0, 0.0, 0.0};
    double nf = 0.0;
    for (int w = 0; w < 4; ++w) {
      DWelford b = {s_w[w][0], s_w[w][1], s_w[w][2]};
      tot = dw_merge(tot, b);
      nf += s_w[w][3];
    }
    double* vp = prm.varpart + ((int64_t)k * prm.ncb_total + prm.cb0 + cb_local) * 4;
    vp[0] = tot.n; vp[1] = tot.mean; vp[2] = tot.m2; vp[3] = nf;
  }
}

// launch of k_dense_finalize_sums for the other translation units (cv_dense_gather.cu)
int cv_dense_finalize_sums(cv_handle* h, const DenseParams& prm) {
  dim3 grid((unsigned)prm.n_cell_blocks, (unsigned)prm.K);
  CV_LAUNCH(h, k_dense_finalize_sums, grid, 128, 0, prm);
  return CV_OK;
}

// Fixed-point table of the integer epilogue: W[k][j] = round(2^F_k / E[k][j]) with F_k chosen so
// that the largest entry of annotation k lies in [2^31, 2^32).  An annotation whose table holds a
// zero or non-finite expectation sum cannot be scaled: flagged, the call then uses the fp64 epilogue.
__global__ void k_winv_fixed(const double* __restrict__ etab, int K, int I, uint32_t* __restrict__ W,
                             double* __restrict__ wscale, uint32_t* __restrict__ dflags) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const double* e = etab + (int64_t)k * I;
  double wmax = 0.0;
  bool bad = false;
  for (int j = 0; j < I; ++j) {
    const double w = 1.0 / e[j];
    if (!(w > 0.0) || !isfinite(w)) bad = true;
    else wmax = fmax(wmax, w);
  }
  int ex = 0;
  if (!bad) frexp(wmax, &ex);                       // wmax = m * 2^ex, m in [0.5, 1)
  const int F = 32 - ex;
  // The rounding of the table entries, delta_j = W_j - 2^F / E_j, shifts the mean over iterations by
  // sum_j s_j delta_j / (B 2^F fps).  Where that matters -- sums s_j nearly equal, i.e. a small spread of
  // the background deviations, which then divides it -- it is (mean + 1) * rho with
  // rho = sum_j delta_j / sum_j W_j, a constant of the annotation: folded into the mean's scale below
  // (the remainder is of relative size spread x 2^-32 / sqrt(B)).
  double dsum = 0.0, wsum = 0.0;
  for (int j = 0; j < I; ++j) {
    const double exact = bad ? 0.0 : ldexp(1.0 / e[j], F);
    double v = rint(exact);
    if (v > 4294967295.0) v = 4294967295.0;
    W[(int64_t)k * I + j] = (uint32_t)v;
    if (j > 0) { dsum += v - exact; wsum += v; }
  }
  const double rho = wsum > 0.0 ? dsum / wsum : 0.0;
  // per-annotation constants of the fixed-point epilogue: 2^-F / B (mean), 2^F (1 / sd), 1 / E_0, E_0
  wscale[4 * k + 0] = bad ? 0.0 : ldexp(1.0, -F) / (double)(I - 1) * (1.0 - rho);
  wscale[4 * k + 1] = bad ? 0.0 : ldexp(1.0, F);
  wscale[4 * k + 2] = 1.0 / e[0];
  wscale[4 * k + 3] = e[0];
  // an empty annotation has an all-zero table and all-NA results whatever the arithmetic
  if (bad && e[0] != 0.0) atomicOr(dflags + 4, 1u);
}

// ---- host side --------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D u8 map, SWIZZLE_128B boxes of 128 bytes x box_rows; rows are row_stride bytes apart and
// row_bytes of each are addressable
int cv_make_map_u8(cv_handle* h, CUtensorMap* map, void* base, uint64_t rows, uint64_t row_bytes,
                   uint64_t row_stride, uint32_t box_rows) {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || !p) CV_FAIL(h, CV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    fn = (PFN_encodeTiled)p;
  }
  cuuint64_t gdim[2] = {row_bytes, rows};
  cuuint64_t gstr[1] = {row_stride};
  cuuint32_t box[2] = {DN_BLOCK_K, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) CV_FAIL(h, CV_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return CV_OK;
}

static int make_map_u8(cv_handle* h, CUtensorMap* map, void* base, uint64_t rows, uint64_t row_bytes,
                       uint32_t box_rows) {
  return cv_make_map_u8(h, map, base, rows, row_bytes, row_bytes, box_rows);
}

static inline int64_t round_up(int64_t v, int64_t m) { return ((v + m - 1) / m) * m; }

// Can the dense path represent this problem exactly, and in which shape?  max_val = largest count
// (0: not known yet -- streaming call -- assume one u8 plane and verify afterwards).
// want_chunk: cells per GEMM launch requested by the caller (0 = as many as memory allows).
int cv_dense_plan(cv_handle* h, uint32_t max_val, int64_t want_chunk, DensePlan* pl, std::string* why) {
  const int I = h->B + 1;
  if (I > DN_BLOCK_N) { *why = "niterations + 1 > 256"; return 0; }
  pl->I = I;
  pl->G = std::min(DN_BLOCK_N / I, DN_MAX_G);
  pl->Ppad = round_up(h->P, DN_BLOCK_K);
  pl->n_groups = (int)((h->K + pl->G - 1) / pl->G);
  pl->rowsM = (int64_t)pl->n_groups * DN_BLOCK_N;
  pl->planes = 1;
  for (uint32_t v = max_val >> 8; v; v >>= 8) pl->planes++;
  // the CTA pair moves a third less through the L2 -> SM crossbar; once the epilogue stopped
  // being the limiter it is the faster kernel in every configuration measured (profiles/r01h)
  pl->pair = h->opt_pair < 0 ? 1 : (h->opt_pair ? 1 : 0);
  pl->m_bytes = pl->rowsM * pl->Ppad;
  // per cell: one u8 row of the operand (+ the int64 sums scratch when digit planes are combined)
  const double per_cell = (double)pl->Ppad + (pl->planes > 1 ? (double)h->K * I * 8.0 : 0.0);
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  const double avail = 0.85 * ((double)free_b + (double)h->dn_M.cap + (double)h->dn_C.cap + (double)h->dn_S.cap);
  double budget = avail - (double)pl->m_bytes;
  if (budget > 24e9) budget = 24e9;                       // operand chunk beyond this buys nothing
  int64_t chunk = (int64_t)(budget / per_cell / 256.0) * 256;
  if (chunk < 256) { *why = "dense operands do not fit in device memory"; return 0; }
  const int64_t n256 = round_up(h->N, 256);
  if (h->opt_chunk > 0) chunk = std::min<int64_t>(chunk, round_up(h->opt_chunk, 256));   // "cell_chunk" overrides the caller's wish
  else if (want_chunk > 0) chunk = std::min<int64_t>(chunk, round_up(want_chunk, 256));
  pl->chunk_cells = std::min(chunk, n256);
  pl->c_bytes = pl->chunk_cells * pl->Ppad;
  pl->sums_bytes = pl->planes > 1 ? pl->chunk_cells * (int64_t)h->K * I * 8 : 0;
  pl->m_bytes_u8 = pl->m_bytes;
  pl->c_bytes_u8 = pl->c_bytes;
  // e2m1 operands (cv_dense_fp4.cu): single-plane counts, fixed-point epilogue, gather-built rows
  pl->fp4 = 0;
  if (h->opt_fp4 != 0 && !h->f4_veto && pl->planes == 1 && pl->pair && !h->opt_rows_scatter &&
      !h->opt_debug_no_tma && cv_f4_plan(h, pl))
    cv_dense_plan_set_fp4(pl, 1);
  return 1;
}

void cv_dense_plan_set_fp4(DensePlan* pl, int on) {
  pl->fp4 = on;
  pl->m_bytes = on ? pl->rowsM4 * pl->rowb : pl->m_bytes_u8;
  pl->c_bytes = on ? pl->chunk_cells * pl->rowb : pl->c_bytes_u8;
}

double cv_dense_cost_ms(cv_handle* h, const DensePlan& pl) {
  const double cells = (double)round_up(h->N, DN_BLOCK_M);
  const double gemm = pl.fp4 ? 2.0 * (double)pl.Ppad4 * (double)pl.rowsM4 * cells / 5.0e15 * 1e3
                             : 2.0 * (double)pl.Ppad * (double)pl.rowsM * cells / 2.5e15 * 1e3 * pl.planes;
  const double build = (double)pl.m_bytes / 1.0e12 * 1e3;           // operand build ~1 TB/s
  return gemm + build + 0.3;
}

template <int MODE>
static int launch_gemm(cv_handle* h, const DensePlan& pl, const CUtensorMap& map_c, const CUtensorMap& map_m,
                       const DenseParams& prm_in) {
  DenseParams prm = prm_in;
  if (pl.pair) {
    const int n_units = (prm.n_cell_blocks + 1) / 2;
    prm.n_tiles = (int64_t)n_units * prm.n_groups;
    CV_CUDA(h, cudaFuncSetAttribute(k_dense_contract_pair<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    DN_SMEM2_BYTES));
    int pairs = h->num_sms / 2;
    if (prm.n_tiles < pairs) pairs = (int)prm.n_tiles;
    k_dense_contract_pair<MODE><<<2 * pairs, DN_THREADS, DN_SMEM2_BYTES, h->stream>>>(map_c, map_m, prm);
  } else {
    prm.n_tiles = (int64_t)prm.n_cell_blocks * prm.n_groups;
    CV_CUDA(h, cudaFuncSetAttribute(k_dense_contract<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    DN_SMEM_BYTES));
    const int grid = (int)std::min<int64_t>(prm.n_tiles, h->num_sms);
    k_dense_contract<MODE><<<grid, DN_THREADS, DN_SMEM_BYTES, h->stream>>>(map_c, map_m, prm);
  }
  h->launches++;
  CV_CUDA(h, cudaGetLastError());
  if (h->opt_sync_launch) CV_TRY(cv_sync_after_launch(h, "k_dense_contract / k_dense_contract_pair"));
  return CV_OK;
}

// Operand M, per-annotation tables and the exactness counters: everything that does not depend on
// the cells.  Asynchronous on the handle's stream.
int cv_dense_prepare(cv_handle* h, const DensePlan& pl, int scatter_rows) {
  CV_TRY(cv_ensure(h, h->dn_M, (size_t)pl.m_bytes));
  CV_TRY(cv_ensure(h, h->dn_C, (size_t)pl.c_bytes));
  if (pl.sums_bytes) CV_TRY(cv_ensure(h, h->dn_S, (size_t)pl.sums_bytes));
  const int ncb_total = (int)((h->N + DN_BLOCK_M - 1) / DN_BLOCK_M);
  CV_TRY(cv_ensure(h, h->varpart, (size_t)h->K * ncb_total * 4 * 8));
  cv_span_begin(h, CV_SPAN_LAYOUT);
  CV_TRY(cv_dense_build_rows(h, pl, scatter_rows));
  CV_TRY(cv_ensure(h, h->dn_winv, (size_t)h->K * pl.I * 4 + (size_t)h->K * 32 + 16));
  CV_LAUNCH(h, k_winv_fixed, (unsigned)((h->K + 127) / 128), 128, 0, h->etab.as<double>(), (int)h->K, pl.I,
            h->dn_winv.as<uint32_t>() + 8 * h->K, reinterpret_cast<double*>(h->dn_winv.p),
            h->flags.as<uint32_t>() + CV_DFLAGS);
  cv_span_end(h);
  h->dn_flops = 0;
  return CV_OK;
}

// Cells [c0, c1) of this shard: densify (one u8 digit plane at a time), GEMM with fused epilogue
// (or raw sums + finish when the counts need more than one plane).  Asynchronous.
int cv_dense_chunk(cv_handle* h, const DensePlan& pl, int64_t c0, int64_t c1, int keep_sums) {
  const int64_t N = h->N, K = h->K;
  const int B = h->B;
  if (c0 % 256 || c1 <= c0 || c1 - c0 > pl.chunk_cells)
    CV_FAIL(h, CV_ERR_INVALID, "dense chunk [%lld, %lld) is not aligned / too large", (long long)c0, (long long)c1);
  const int64_t rows = round_up(c1 - c0, 256);
  if (pl.planes > 1) CV_TRY(cv_ensure(h, h->dn_S, (size_t)pl.chunk_cells * (size_t)K * (B + 1) * 8));
  CUtensorMap map_c, map_m;
  if (!pl.fp4) {
    CV_TRY(make_map_u8(h, &map_c, h->dn_C.p, (uint64_t)rows, (uint64_t)pl.Ppad, DN_BLOCK_M));
    CV_TRY(make_map_u8(h, &map_m, h->dn_M.p, (uint64_t)pl.rowsM, (uint64_t)pl.Ppad, pl.pair ? DN_BLOCK_N / 2 : DN_BLOCK_N));
  }
  DenseParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.N = (int)N; prm.K = (int)K; prm.B = B; prm.G = pl.G;
  prm.cb0 = (int)(c0 / DN_BLOCK_M);
  prm.c_end = (int)std::min<int64_t>(c1, N);
  prm.n_cell_blocks = (int)((std::min<int64_t>(c1, N) - c0 + DN_BLOCK_M - 1) / DN_BLOCK_M);
  prm.ncb_total = (int)((N + DN_BLOCK_M - 1) / DN_BLOCK_M);
  prm.n_groups = pl.n_groups; prm.n_kblocks = (int)(pl.Ppad / DN_BLOCK_K);
  prm.band = h->opt_band > 0 ? h->opt_band : 12;
  prm.debug_no_tma = h->opt_debug_no_tma;
  prm.annoptr = h->annoptr.as<int64_t>();
  prm.etab = h->etab.as<double>();
  prm.fps = h->fps.as<double>();
  prm.dev_rm = h->dev_rm.as<double>();
  prm.z_rm = h->z_rm.as<double>();
  prm.varpart = h->varpart.as<double>();
  prm.obs = keep_sums ? h->obs.as<long long>() : nullptr;
  prm.samp = keep_sums ? h->samp.as<long long>() : nullptr;
  prm.wscale = reinterpret_cast<const double*>(h->dn_winv.p);            // [K][4] doubles, then [K][I] u32
  prm.winv_fx = h->dn_winv.as<uint32_t>() + 8 * K;
  prm.sums = h->dn_S.as<long long>();
  prm.sums_ld = (int)pl.chunk_cells;
  prm.sums_c0 = (int)c0;
  if (pl.fp4) {
    // e2m1 operands: exact while every sum fits a float (checked like the fixed-point epilogue)
    if (!((double)pl.cell_tot_bound * (double)h->dn_maxmult < 16777216.0)) {
      h->f4_veto = true;
      CV_FAIL(h, CV_ERR_UNSUPPORTED, "sums of this cell chunk may exceed 2^24: not exact on the e2m1 path");
    }
    const int64_t tiles = (int64_t)prm.n_cell_blocks * prm.n_groups;
    const bool want32 = h->opt_epilogue == 1 ||
                        (h->opt_epilogue == 0 && pl.Ppad4 / 256 < 1100 && tiles >= 4 * (int64_t)h->num_sms);
    const int fx = want32 && h->dn_fixed_ok;
    CV_TRY(cv_f4_chunk(h, pl, prm, c0, std::min<int64_t>(c1, N), rows, fx, c0 == 0 && !h->replaying));
    h->st.epilogue = fx ? 32 : 64;
    h->dn_flops += 2.0 * (double)pl.rowsM4 * (double)rows * (double)pl.Ppad4;
    return CV_OK;
  }
  for (int plane = 0; plane < pl.planes; ++plane) {
    cv_span_begin(h, CV_SPAN_LAYOUT);
    CV_TRY(cv_dense_build_cells(h, pl, c0, rows, 8 * plane));
    cv_span_end(h);
    cv_span_begin(h, CV_SPAN_CONTRACT);
    if (pl.planes == 1) {
      // fp64 epilogue when it hides behind the next tile's MMAs anyway (long tiles) or when the sums
      // may not fit a float exactly; the float-pair epilogue otherwise ("dense_epilogue": 1 / 2 force)
      const bool exact32 = h->dn_fixed_ok && (double)pl.cell_tot_bound * (double)h->dn_maxmult < 16777216.0;
      const int64_t tiles = (int64_t)prm.n_cell_blocks * prm.n_groups;
      const bool want32 = h->opt_epilogue == 1 ||
                          (h->opt_epilogue == 0 && prm.n_kblocks < 1100 && tiles >= 4 * (int64_t)h->num_sms);
      if (want32 && exact32) CV_TRY(launch_gemm<DN_EPI_FUSED32>(h, pl, map_c, map_m, prm));
      else CV_TRY(launch_gemm<DN_EPI_FUSED>(h, pl, map_c, map_m, prm));
      h->st.epilogue = (want32 && exact32) ? 32 : 64;
    } else {
      h->st.epilogue = 64;
      prm.sums_shift = 8 * plane;
      prm.sums_acc = plane > 0;
      CV_TRY(launch_gemm<DN_EPI_SUMS>(h, pl, map_c, map_m, prm));
    }
    cv_span_end(h);
    h->dn_flops += 2.0 * (double)pl.rowsM * (double)rows * (double)pl.Ppad;
  }
  if (pl.planes > 1) {
    cv_span_begin(h, CV_SPAN_CONTRACT);
    dim3 grid((unsigned)prm.n_cell_blocks, (unsigned)K);
    CV_LAUNCH(h, k_dense_finalize_sums, grid, 128, 0, prm);
    cv_span_end(h);
  }
  return CV_OK;
}

// Exactness bounds of the u8 operands / s32 accumulators, from the device-side counters
// (hflags = the handle's 16 flag words copied to the host after a stream sync).  with_counts = 0
// checks only what is known before the counts have arrived (streamed call).
int cv_dense_check(cv_handle* h, const DensePlan& pl, const uint32_t* hflags, int with_counts, int* soft_fail) {
  *soft_fail = 0;
  const uint32_t* df = hflags + CV_DFLAGS;
  const uint64_t maxmult = (uint64_t)std::max(df[1], 1u) * std::max(df[2], 1u);
  if (df[2] >= 0xFFFFu || maxmult > 255) {
    *soft_fail = 1;
    CV_FAIL(h, CV_ERR_UNSUPPORTED, "annotation multiplicity bound %llu exceeds the u8 operand",
            (unsigned long long)maxmult);
  }
  h->dn_maxmult = maxmult;
  h->dn_fixed_ok = df[4] == 0;                 // every expectation sum could be scaled to 32 bits
  if (!with_counts) return CV_OK;
  // every digit plane's sum is bounded by (cell total) x multiplicity
  if ((double)h->max_cell_tot * (double)maxmult >= 2147483647.0) {
    *soft_fail = 1;
    CV_FAIL(h, CV_ERR_UNSUPPORTED, "s32 accumulator bound exceeded");
  }
  return CV_OK;
}
