// cv_dense_gather.cu -- the contraction for FEW cells (bulk ATAC / DNase: a few hundred samples).
//
// With one 256-cell unit the materialised-operand GEMM of cv_dense.cu moves ~100 bytes per useful MAC:
// the stacked operand M (annotation x iteration rows, 22 GB at config 4) is written once and
// streamed once per digit plane.  Here nothing is stacked.  Per iteration j the sums are
//
//     S_j[k, c] = sum_p A[k, p] * C[bg_j(p), c]          (bg_0 = identity: the observed sums)
//
// i.e. a GEMM of the annotation membership matrix A (K x P, u8 0/1, K-major, loaded by TMA) with the
// rows of the counts GATHERED by the background indices of that iteration: producer warps fetch row
// bg_j(p) of the cell-major counts planes (256 contiguous bytes per peak and digit plane) with
// cp.async straight into the shared-memory tile of an MN-major tcgen05 operand (cells contiguous,
// peaks strided, 128-byte swizzle applied by address arithmetic).  tcgen05.mma kind::i8,
// cta_group::2: a CTA pair owns 256 annotations x 256 cells; both u8 digit planes of the counts
// accumulate side by side in the 512 TMEM columns, so A is read once per iteration and the planes are
// combined in registers (s0 + 256 s1, exact in int64).  Work item = (256-annotation tile, iteration);
// items of one iteration run next to each other, so the gathered rows are shared through L2.
// The raw sums go to the int64 scratch of the digit-plane path and k_dense_finalize_sums applies the
// reference arithmetic (R/compute_deviations.R:488-520), so results are bit-identical to that path.
//
// Applies when: N <= 256 cells, counts < 65536, no annotation lists a peak twice.
#include <cuda.h>

#include <algorithm>

#include "cv_internal.cuh"

#include "cv_dense_dev.cuh"

#define GA_THREADS 384          // warp 0 TMA (A), 1 MMA, 2 TMEM allocation, 4-7 epilogue, 8-11 gather producers
#define GA_STAGES 4
#define GA_A_BYTES (128 * 128)  // 128 annotations x 128 peaks (K-major, SWIZZLE_128B)
#define GA_B_BYTES (128 * 128)  // 128 peaks x 128 cells of one digit plane (MN-major, SWIZZLE_128B)
#define GA_STAGE_BYTES (GA_A_BYTES + 2 * GA_B_BYTES)
#define GA_SMEM_BYTES (GA_STAGES * GA_STAGE_BYTES + 1024 + 512)
#define GA_CELLS 256
#define GA_WATCHDOG (1u << 24)  // polls of one mbarrier wait before the kernel traps instead of hanging

struct GatherParams {
  int K, P, I, N;                  // annotations, peaks, iterations + 1, cells (<= 256)
  int n_store;                     // cell columns whose sums are kept: N (+ the expectation's digit columns)
  int n_kblocks;                   // 128-peak blocks (peak axis padded)
  int n_tiles;                     // 256-annotation tiles
  int64_t ppad;                    // padded peaks = row stride of bgx
  int64_t rows_cd;                 // rows of one counts plane (ppad + 1: the last one is all zero)
  const int32_t* bgx;              // [I][ppad] source peak of (iteration, peak); pad peaks -> the zero row
  const uint8_t* cdT;              // [2][rows_cd][256] counts digit planes, cells contiguous
  long long* sums;                 // [K][I][256]
  int debug;                       // experiments only (garbage results): 1 no gathers, 2 no MMAs, 3 no A loads
};

__device__ __forceinline__ void ga_wait(uint64_t* bar, uint32_t parity, unsigned ns) {
  uint32_t done, polls = 0;
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
    if (ns) __nanosleep(ns);
    if (++polls > GA_WATCHDOG) __trap();
  }
}

// MN-major operand tile, SWIZZLE_128B: 16-byte units, ((8, n), (8, k)) : ((1, LBO), (8, SBO)) -- 128 bytes
// of the MN axis contiguous, 8 K-rows 128 bytes apart form the 1 KB swizzle atom, K-groups SBO apart.
// One atom along MN per CTA here (128 cells), so LBO is unused.
__device__ __forceinline__ uint64_t ga_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 16;                     // leading byte offset (next MN atom; unused)
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset: next group of 8 K-rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
  return d;
}
// u8 x u8 -> s32, A K-major, B MN-major (bit 16)
__host__ __device__ constexpr uint32_t ga_idesc(int M, int N) {
  return (2u << 4) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GA_THREADS, 1)
k_gather_contract(const __grid_constant__ CUtensorMap map_a, const GatherParams prm) {
  extern __shared__ unsigned char ga_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ga_smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* smem_a = smem;                                        // [STAGES][16 KB]
  unsigned char* smem_b = smem + GA_STAGES * GA_A_BYTES;               // [STAGES][2 planes][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GA_STAGES * GA_STAGE_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES] leader's: 1 (TMA of both CTAs, with tx bytes) + 1 (the peer's relay)
  uint64_t* empty_bar = bars + GA_STAGES;       // [STAGES] 1 (tcgen05.commit, multicast to both CTAs)
  uint64_t* gfull_bar = bars + 2 * GA_STAGES;   // [STAGES] this CTA's gathered rows: 128 asynchronous cp.async arrivals
  uint64_t* tfull_bar = bars + 3 * GA_STAGES;   // accumulators complete
  uint64_t* tempty_bar = bars + 3 * GA_STAGES + 1;   // leader's: 256 epilogue threads
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * GA_STAGES + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_items = prm.n_tiles * prm.I;      // item = j * n_tiles + tile: items of one iteration are neighbours

  if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < GA_STAGES; ++s) { mbar_init(full_bar + s, 2); mbar_init(empty_bar + s, 1); mbar_init(gfull_bar + s, 128); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own 128 annotations x 128 peaks of A =====
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int item = pair_id; item < n_items; item += n_pairs) {
        const int tile = item % prm.n_tiles;
        const int rt = tile * 2 + (int)rank;               // 128-annotation row tile of this CTA
        for (int kb = 0; kb < prm.n_kblocks; ++kb) {
          ga_wait(empty_bar + s, ph ^ 1u, 32);
          if (prm.debug == 3) {
            if (rank == 0) mbar_arrive(full_bar + s);
          } else {
            if (rank == 0) mbar_expect_tx(full_bar + s, 2 * GA_A_BYTES);
            tma_load_2d_pair(&map_a, mapa_u32(smem_u32(full_bar + s), 0), smem_a + s * GA_A_BYTES, 0,
                             (rt * prm.n_kblocks + kb) * 128);
          }
          if (++s == GA_STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA): both digit planes per k-block, accumulators side by side =====
    if (rank == 0) {
      constexpr uint32_t idesc = ga_idesc(256, GA_CELLS);
      int s = 0;
      uint32_t ph = 0, it = 0;
      for (int item = pair_id; item < n_items; item += n_pairs, ++it) {
        ga_wait(tempty_bar, (it & 1u) ^ 1u, 0);
        tc_fence_after();
        for (int kb = 0; kb < prm.n_kblocks; ++kb) {
          ga_wait(gfull_bar + s, ph, 0);                 // own rows have landed (cp.async completions)
          ga_wait(full_bar + s, ph, 0);                  // A tiles of both CTAs (TMA) + the peer's rows (its relay)
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // rows written through the generic proxy -> tensor core
          tc_fence_after();
          const uint64_t da = make_smem_desc(smem_u32(smem_a + s * GA_A_BYTES));
          const uint32_t b0 = smem_u32(smem_b + s * 2 * GA_B_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int pl = 0; pl < 2; ++pl)
#pragma unroll
              for (int k = 0; k < 4; ++k)          // 32 peaks per instruction: 32 bytes of A's row, 4 K-groups of B
                if (prm.debug != 2)
                umma_i8_pair(tmem_base + (uint32_t)(pl * GA_CELLS), da + (uint64_t)(k * 2),
                             ga_desc_mn(b0 + (uint32_t)(pl * GA_B_BYTES + k * 4096)), idesc, (kb | k) ? 1u : 0u);
            tc_commit_pair(empty_bar + s, 3);
          }
          __syncwarp();
          if (++s == GA_STAGES) { s = 0; ph ^= 1u; }
        }
        if (elect_one()) tc_commit_pair(tfull_bar, 3);
        __syncwarp();
      }
    } else if (lane == 0) {
      // ===== the peer's relay: its gathered rows of a stage have landed -> one arrival on the leader's barrier =====
      int s = 0;
      uint32_t ph = 0;
      for (int item = pair_id; item < n_items; item += n_pairs) {
        for (int kb = 0; kb < prm.n_kblocks; ++kb) {
          ga_wait(gfull_bar + s, ph, 0);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive_cluster(mapa_u32(smem_u32(full_bar + s), 0));
          if (++s == GA_STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp >= 8) {
    // ===== gather producers (both CTAs): this CTA's 128 cells of both planes for the 128 peaks of a k-block =====
    // Random 128-byte rows come back after microseconds, so a thread never waits for its own copies: it
    // fills every free stage ahead and lets the copies count themselves in on the stage's barrier when
    // they land (cp.async.mbarrier.arrive.noinc).
    // Thread t copies 16-byte chunk (t & 7) of the rows (t >> 3) + 16 i, i = 0..7: eight consecutive lanes
    // fetch one whole 128-byte row half (four full sectors, one line), a warp instruction four rows -- not 32
    // chunks of 32 different lines, which doubles the L2 -> SM bytes and floods the LSU with requests
    // (capture r02u: 131 GB through the crossbar for 52 GB of rows, stall_lg on every LDGSTS).
    const int t = threadIdx.x - 256;
    const int ch = t & 7, r0 = t >> 3;                 // r0 in 0..15
    // row r = r0 + 16 i: atom (r >> 3) = (r0 >> 3) + 2 i, row in atom = r0 & 7
    const uint32_t row_off = (uint32_t)((r0 >> 3) * 1024 + (r0 & 7) * 128 + ((ch ^ (r0 & 7)) << 4));
    int s = 0;
    uint32_t ph = 0;
    for (int item = pair_id; item < n_items; item += n_pairs) {
      const int j = item / prm.n_tiles;
      const int32_t* __restrict__ bgj = prm.bgx + (int64_t)j * prm.ppad + r0;
      int32_t next_row[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) next_row[i] = __ldg(bgj + 16 * i);
      for (int kb = 0; kb < prm.n_kblocks; ++kb) {
        int32_t src_row[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) src_row[i] = next_row[i];
        if (kb + 1 < prm.n_kblocks) {                      // one k-block ahead of its use
#pragma unroll
          for (int i = 0; i < 8; ++i) next_row[i] = __ldg(bgj + (int64_t)(kb + 1) * 128 + 16 * i);
        }
        ga_wait(empty_bar + s, ph ^ 1u, 20);
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
          const unsigned char* plane = prm.cdT + (int64_t)pl * prm.rows_cd * GA_CELLS + rank * 128 + ch * 16;
          const uint32_t dst = smem_u32(smem_b + (s * 2 + pl) * GA_B_BYTES) + row_off;
          if (prm.debug != 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)(i * 2048)),
                           "l"(plane + (int64_t)src_row[i] * GA_CELLS) : "memory");
          }
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(gfull_bar + s)) : "memory");
        if (++s == GA_STAGES) { s = 0; ph ^= 1u; }
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp >= 4 && warp < 8) {
    // ===== epilogue (both CTAs): thread = annotation (TMEM lane); s0 + 256 s1 -> int64 scratch =====
    const int q4 = warp & 3;
    uint32_t it = 0;
    for (int item = pair_id; item < n_items; item += n_pairs, ++it) {
      const int tile = item % prm.n_tiles, j = item / prm.n_tiles;
      const int k = tile * 256 + (int)rank * 128 + q4 * 32 + lane;
      ga_wait(tfull_bar, it & 1u, 128);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16);
      long long* dst = prm.sums + ((int64_t)k * prm.I + j) * GA_CELLS;
      for (int c0 = 0; c0 < GA_CELLS; c0 += 32) {
        uint32_t r0[32], r1[32];
        tmem_ld32(taddr + (uint32_t)c0, r0);
        tmem_ld32(taddr + (uint32_t)(GA_CELLS + c0), r1);
        tmem_ld_wait();
        if (k < prm.K) {
#pragma unroll
          for (int u = 0; u < 32; ++u)
            if (c0 + u < prm.n_store) dst[c0 + u] = (long long)(int)r0[u] + ((long long)(int)r1[u] << 8);
        }
      }
      tc_fence_before();
      mbar_arrive_cluster(mapa_u32(smem_u32(tempty_bar), 0));
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- operand builders ----------------------------------------------------------------------------------------
// A8, tiled: the 128 annotations x 128 peaks of (row tile rt, k-block kb) are one contiguous 16 KB block
// at ((rt * n_kb + kb) * 128 + k % 128) * 128 + p % 128, so a TMA box is one sequential read (rows of the
// plain K x P layout are 500 KB apart: 128 DRAM pages per box and microseconds of latency).
// Entry = bit k of Abits[k / 32][p]  (u8 0/1; rows >= K and peaks >= P stay zero)
__global__ void __launch_bounds__(256)
k_ga_build_a(const uint32_t* __restrict__ Abits, int K, int P, int64_t ppad, uint8_t* __restrict__ A8) {
  const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int w = blockIdx.y;
  if (p >= ppad) return;
  const uint32_t word = p < P ? Abits[(int64_t)w * P + p] : 0u;
  const int64_t n_kb = ppad / 128;
  const int64_t rt = (w * 32) / 128, kb = p / 128;
  uint8_t* blk = A8 + ((rt * n_kb + kb) * 128 + (w * 32) % 128) * 128 + (p % 128);
#pragma unroll 4
  for (int b = 0; b < 32; ++b) blk[b * 128] = (uint8_t)((word >> b) & 1u);
}

// counts planes, cells contiguous: cdT[plane][peak][cell] = byte `plane` of the count; one CTA per cell
// streams its column (coalesced reads, one byte per entry scattered into the 256-byte rows; a tile-wise
// variant with binary searches into the columns was slower: 3.1 vs 1.3 ms at config 4)
__global__ void __launch_bounds__(512)
k_ga_build_cd(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowidx, const uint32_t* __restrict__ val,
              int N, int64_t rows_cd, uint8_t* __restrict__ cdT) {
  const int c = blockIdx.x;
  if (c >= N) return;
  for (int64_t e = colptr[c] + threadIdx.x; e < colptr[c + 1]; e += 512) {
    const uint32_t v = val[e];
    const int64_t q = rowidx[e];
    // (a column listing a peak twice would need the sum: such input takes the other paths, K1's verdict)
    cdT[q * GA_CELLS + c] = (uint8_t)(v & 0xFFu);
    if (v >> 8) cdT[(rows_cd + q) * GA_CELLS + c] = (uint8_t)((v >> 8) & 0xFFu);
  }
}

// The expectation rides on the GEMM: E[k][j] = sum_p A[k, p] e[bg_j(p)] is one more "cell" of the same
// contraction.  e[q] is written in fixed point, GA_EDIGITS base-256 digits with the unit 2^-GA_EUNIT (every
// double in [2^-GA_EUNIT+52, 1) is represented exactly: 2^-40 <= e covers matrices up to 1e12 fragments), one
// digit per spare cell column of digit plane 0; the integer sums of the digits are exact, so the table is
// the correctly rounded sum of the exact values -- no fp64 summation order at all.
#define GA_EDIGITS 12
#define GA_EUNIT 92
__global__ void __launch_bounds__(256)
k_ga_build_edigits(const double* __restrict__ e, int P, int N, uint8_t* __restrict__ cdT, uint32_t* __restrict__ dflags) {
  const int q = blockIdx.x * 256 + threadIdx.x;
  if (q >= P) return;
  const double v = e[q];
  uint8_t* row = cdT + (int64_t)q * GA_CELLS + N;
  if (!(v >= 0.0) || !(v < 1.0)) { atomicOr(dflags + 4, 2u); return; }      // outside the fixed-point range: the table kernel instead
  int ex;
  const double m = frexp(v, &ex);                              // v = m 2^ex, m in [0.5, 1)
  unsigned long long mant = (unsigned long long)ldexp(m, 53);  // 53-bit integer
  int sh = ex - 53 + GA_EUNIT;                                 // v = mant 2^(ex - 53) = mant 2^sh units
  if (v != 0.0 && sh < 0) {
    if (mant & ((1ull << min(-sh, 63)) - 1ull)) atomicOr(dflags + 4, 2u);     // bits below the unit: not exact
    mant = -sh >= 64 ? 0ull : mant >> -sh;
    sh = 0;
  }
  // mant << sh as GA_EDIGITS bytes
  for (int d = 0; d < GA_EDIGITS; ++d) {
    const int bit = 8 * d - sh;                                // bit position inside mant of this digit's LSB
    unsigned long long piece = 0ull;
    if (bit > -8 && bit < 64) piece = bit >= 0 ? (mant >> bit) : (mant << -bit);
    row[d] = (uint8_t)(piece & 0xFFull);
  }
}

// etab[k][j] from the digit sums S[k][j][N + d] (each < 2^31): value = sum_d S_d 256^d units, rounded once
__global__ void __launch_bounds__(256)
k_ga_etab_from_sums(const long long* __restrict__ sums, int K, int I, int N, double* __restrict__ etab) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= (int64_t)K * I) return;
  const long long* s = sums + i * GA_CELLS + N;
  unsigned long long lo = 0ull, hi = 0ull;                     // 128-bit accumulator
  for (int d = 0; d < GA_EDIGITS; ++d) {
    const unsigned long long v = (unsigned long long)s[d];
    const int sh = 8 * d;
    unsigned long long add_lo, add_hi;
    if (sh < 64) { add_lo = v << sh; add_hi = sh ? v >> (64 - sh) : 0ull; }
    else { add_lo = 0ull; add_hi = v << (sh - 64); }
    const unsigned long long nl = lo + add_lo;
    hi += add_hi + (nl < lo ? 1ull : 0ull);
    lo = nl;
  }
  // hi 2^64 + lo -> double with a single rounding: fold the low word's sticky bit in when the high word is set
  double r;
  if (hi == 0ull) r = (double)lo;
  else {
    const int lz = __clzll((long long)hi);
    const unsigned long long top = lz ? (hi << lz) | (lo >> (64 - lz)) : hi;     // the 64 leading bits
    const unsigned long long rest = lz ? lo << lz : lo;
    r = ldexp((double)(top | (rest ? 1ull : 0ull)), 64 - lz);
  }
  etab[i] = ldexp(r, -GA_EUNIT);
}

// bgx[j][p]: j = 0 the peak itself, j >= 1 its background peak of iteration j; pad peaks -> the zero row
__global__ void __launch_bounds__(256)
k_ga_build_idx(const int32_t* __restrict__ bg_raw, int base, int P, int I, int64_t ppad, int32_t zero_row,
               int32_t* __restrict__ bgx) {
  const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int j = blockIdx.y;
  if (p >= ppad) return;
  int32_t v = zero_row;
  if (p < P) {
    if (j == 0) v = (int32_t)p;
    else {
      const int q = bg_raw[(int64_t)(j - 1) * P + p] - base;
      v = (q >= 0 && q < P) ? q : zero_row;          // out-of-range indices are flagged by the table kernels
    }
  }
  bgx[(int64_t)j * ppad + p] = v;
}

// ---- host side -------------------------------------------------------------------------------------------------
// Is the few-cells path exact for the resident problem?  (duplicates inside a column of the counts or
// inside an annotation list are found on the device by the callers' flag words)
int cv_gather_applicable(const cv_handle* h) {
  if (h->opt_gather == 0 || h->opt_path == 1) return 0;
  if (h->N > GA_CELLS || h->N < 1 || h->max_val >= 65536u) return 0;
  if (h->B + 1 > 1024 || h->K < 1) return 0;
  // s32 accumulators per digit plane: at most 255 x (largest list)
  int64_t set_max = 0;
  for (int64_t k = 0; k < h->K; ++k) set_max = std::max(set_max, h->h_annoptr[(size_t)k + 1] - h->h_annoptr[(size_t)k]);
  if ((double)set_max * 255.0 >= 2147483647.0) return 0;
  // worth it only when the stacked operand would be large: its build alone writes rows x P bytes
  const double m_bytes = (double)h->K * (h->B + 1) * (double)h->P;
  return h->opt_gather > 0 || m_bytes > 1.0e9;
}

// Can the expectation table ride on the GEMM (spare cell columns for its fixed-point digits)?
int cv_gather_etab_in_gemm(const cv_handle* h) { return h->N + GA_EDIGITS <= GA_CELLS; }

// Operands of the few-cells path: A8 (tiled), counts planes (+ expectation digits), index table.  The bit
// columns must have been enqueued (cv_dense_build_rows with tables_only); asynchronous.
int cv_gather_prepare(cv_handle* h) {
  const int64_t P = h->P, N = h->N, K = h->K;
  const int I = h->B + 1;
  const int64_t ppad = ((P + 127) / 128) * 128, rows_cd = ppad + 1;
  const int64_t kpad = ((K + 255) / 256) * 256;
  const int64_t Kw = (K + 31) / 32;
  if ((kpad / 128) * (ppad / 128) * 128 > 0x7fffffffll) CV_FAIL(h, CV_ERR_UNSUPPORTED, "annotation operand too large for the few-cells path");
  CV_TRY(cv_ensure(h, h->dn_M, (size_t)kpad * ppad));                       // A8
  CV_TRY(cv_ensure(h, h->dn_C, (size_t)2 * rows_cd * GA_CELLS));            // counts planes
  CV_TRY(cv_ensure(h, h->dn_invsrc, (size_t)I * ppad * 4));                 // index table
  CV_TRY(cv_ensure(h, h->dn_S, (size_t)kpad * I * GA_CELLS * 8));           // raw sums
  cv_span_begin(h, CV_SPAN_LAYOUT);
  CV_CUDA(h, cudaMemsetAsync(h->dn_M.p, 0, (size_t)kpad * ppad, h->stream));
  CV_LAUNCH(h, k_ga_build_a, dim3((unsigned)((ppad + 255) / 256), (unsigned)Kw), 256, 0, h->dn_abits.as<uint32_t>(), (int)K,
            (int)P, ppad, h->dn_M.as<uint8_t>());
  CV_CUDA(h, cudaMemsetAsync(h->dn_C.p, 0, (size_t)2 * rows_cd * GA_CELLS, h->stream));
  CV_LAUNCH(h, k_ga_build_cd, (unsigned)N, 512, 0, h->colptr.as<int64_t>(), h->rowidx.as<int32_t>(), h->val.as<uint32_t>(),
            (int)N, rows_cd, h->dn_C.as<uint8_t>());
  if (cv_gather_etab_in_gemm(h))
    CV_LAUNCH(h, k_ga_build_edigits, (unsigned)((P + 255) / 256), 256, 0, h->expect.as<double>(), (int)P, (int)N,
              h->dn_C.as<uint8_t>(), h->flags.as<uint32_t>() + CV_DFLAGS);
  CV_LAUNCH(h, k_ga_build_idx, dim3((unsigned)((ppad + 255) / 256), (unsigned)I), 256, 0, h->bg_raw.as<int32_t>(),
            h->index_base, (int)P, I, ppad, (int32_t)ppad, h->dn_invsrc.as<int32_t>());
  cv_span_end(h);
  return CV_OK;
}

// GEMM + finish on the operands of cv_gather_prepare.  The overlap / matches vectors (and the expectation
// table when it does not ride on the GEMM) come from the auxiliary stream; asynchronous.
int cv_gather_run(cv_handle* h, int keep_sums) {
  const int64_t P = h->P, N = h->N, K = h->K;
  const int I = h->B + 1;
  const int64_t ppad = ((P + 127) / 128) * 128, rows_cd = ppad + 1;
  const int64_t kpad = ((K + 255) / 256) * 256;
  const int ncb = (int)((N + DN_BLOCK_M - 1) / DN_BLOCK_M);
  CV_TRY(cv_ensure(h, h->varpart, (size_t)K * ncb * 4 * 8));
  CUtensorMap map_a;
  // the tiled A8 as a (kpad / 128 * n_kb * 128) x 128-byte array: box = one contiguous 16 KB tile
  CV_TRY(cv_make_map_u8(h, &map_a, h->dn_M.p, (uint64_t)((kpad / 128) * (ppad / 128) * 128), 128, 128, 128));
  GatherParams gp;
  gp.K = (int)K; gp.P = (int)P; gp.I = I; gp.N = (int)N;
  gp.n_store = (int)N + (cv_gather_etab_in_gemm(h) ? GA_EDIGITS : 0);
  gp.n_kblocks = (int)(ppad / 128);
  gp.n_tiles = (int)(kpad / 256);
  gp.ppad = ppad;
  gp.rows_cd = rows_cd;
  gp.bgx = h->dn_invsrc.as<int32_t>();
  gp.cdT = h->dn_C.as<uint8_t>();
  gp.sums = h->dn_S.as<long long>();
  gp.debug = h->opt_debug_no_tma;
  const int n_items = gp.n_tiles * I;
  const int pairs = std::min(n_items, std::max(1, h->num_sms / 2));
  cv_span_begin(h, CV_SPAN_CONTRACT);
  CV_CUDA(h, cudaFuncSetAttribute(k_gather_contract, cudaFuncAttributeMaxDynamicSharedMemorySize, GA_SMEM_BYTES));
  k_gather_contract<<<2 * pairs, GA_THREADS, GA_SMEM_BYTES, h->stream>>>(map_a, gp);
  h->launches++;
  CV_CUDA(h, cudaGetLastError());
  if (h->opt_sync_launch) CV_TRY(cv_sync_after_launch(h, "k_gather_contract"));
  // the expected-fraction table: from the digit sums of the spare cell columns, or (no spare columns)
  // from the auxiliary stream, which has been building it beside the GEMM
  if (cv_gather_etab_in_gemm(h))
    CV_LAUNCH(h, k_ga_etab_from_sums, (unsigned)((K * I + 255) / 256), 256, 0, h->dn_S.as<long long>(), (int)K, I, (int)N,
              h->etab.as<double>());
  else
    CV_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_aux[2], 0));
  // the reference arithmetic on the exact sums (same kernel as the digit-plane path)
  DenseParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.N = (int)N; prm.K = (int)K; prm.B = h->B;
  prm.cb0 = 0; prm.c_end = (int)N; prm.n_cell_blocks = ncb; prm.ncb_total = ncb;
  prm.annoptr = h->annoptr.as<int64_t>();
  prm.etab = h->etab.as<double>();
  prm.fps = h->fps.as<double>();
  prm.dev_rm = h->dev_rm.as<double>();
  prm.z_rm = h->z_rm.as<double>();
  prm.varpart = h->varpart.as<double>();
  prm.obs = keep_sums ? h->obs.as<long long>() : nullptr;
  prm.samp = keep_sums ? h->samp.as<long long>() : nullptr;
  prm.sums = h->dn_S.as<long long>();
  prm.sums_ld = GA_CELLS;
  prm.sums_c0 = 0;
  CV_TRY(cv_dense_finalize_sums(h, prm));
  cv_span_end(h);
  h->dn_flops = 2.0 * (double)kpad * GA_CELLS * (double)ppad * I * 2.0;
  return CV_OK;
}
