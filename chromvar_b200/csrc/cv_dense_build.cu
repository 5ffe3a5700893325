// Operand construction for the dense tensor-core path (cv_dense.cu), built in shared memory and
// written out coalesced so that neither operand needs a global memset or global atomics:
//
//   k_dense_rows   one CTA per annotation k: the B+1 stacked rows M[(k, j), :] (u8 multiplicities),
//                  together with everything else that is per (annotation, iteration) and cell
//                  independent: the expected-fraction table E[k][j] (tf_vec %*% expectation,
//                  sample_mat %*% expectation; R/compute_deviations.R:488,502), the background
//                  overlap (:506, :477-478) and fractionMatches (:533).  Row j = 0 (the set's own
//                  multiplicities) stays in shared memory while rows j >= 1 are built, so the
//                  overlap sum_q M[(k,j),q] * M[(k,0),q] costs one shared-memory byte read per draw.
//   k_dense_cells  one CTA per cell: the densified counts row Cd[c, :] (u8) from the CSC column.
//   k_bg_mult_*    exactness bound for the u8 operand: max_{j,q} #{p : bg[p, j] = q}.
//
// Rows longer than the shared-memory budget (P > ~110k peaks) are built in chunks of the peak axis.
#include <algorithm>
#include <vector>

#include "cv_internal.cuh"

#define DB_THREADS 512
#define DB_BLOCK_N 256
#define DB_UNROLL 4

enum { DB_FLAG_BADIDX = 8u };

__device__ __forceinline__ void db_zero(uint32_t* row, int nwords) {
  uint4* r4 = reinterpret_cast<uint4*>(row);
  for (int i = threadIdx.x; i < (nwords >> 2); i += DB_THREADS) r4[i] = make_uint4(0, 0, 0, 0);
}
__device__ __forceinline__ void db_store(const uint32_t* row, uint32_t* __restrict__ dst, int nwords) {
  const uint4* r4 = reinterpret_cast<const uint4*>(row);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  for (int i = threadIdx.x; i < (nwords >> 2); i += DB_THREADS) d4[i] = r4[i];
}

// bg: P x B column-major as uploaded (index_base); annoidx: 0-based, validated.
__global__ void __launch_bounds__(DB_THREADS, 1)
k_dense_rows(const int64_t* __restrict__ annoptr, const int32_t* __restrict__ annoidx,
             const int32_t* __restrict__ bg, int base, const double* __restrict__ e,
             const int32_t* __restrict__ order, int K, int P, int B, int G, int64_t Ppad, int chunk,
             uint8_t* __restrict__ M, double* __restrict__ etab, double* __restrict__ overlap,
             double* __restrict__ matches, uint32_t* __restrict__ work_ctr, uint32_t* __restrict__ flags) {
  extern __shared__ __align__(16) uint32_t db_smem[];
  uint32_t* row0 = db_smem;                       // [chunk / 4] words: the set's own multiplicities
  uint32_t* rowj = db_smem + (chunk >> 2);        // [chunk / 4] words: the row being built
  __shared__ double s_red[DB_THREADS / 32];
  __shared__ unsigned long long s_ov[DB_THREADS / 32];
  __shared__ int s_item;
  const int I = B + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint8_t* row0b = reinterpret_cast<const uint8_t*>(row0);
  for (;;) {
    if (threadIdx.x == 0) s_item = (int)atomicAdd(work_ctr, 1u);
    __syncthreads();
    const int item = s_item;
    if (item >= K) break;
    const int k = order[item];
    const int64_t beg = annoptr[k], end = annoptr[k + 1];
    const int64_t mrow0 = (int64_t)(k / G) * DB_BLOCK_N + (int64_t)(k % G) * I;
    unsigned long long ov_total = 0;               // thread 0 only
    for (int64_t q0 = 0; q0 < Ppad; q0 += chunk) {
      const int cw = (int)min((int64_t)chunk, Ppad - q0) >> 2;     // words in this chunk
      const int qhi = (int)min((int64_t)P, q0 + ((int64_t)cw << 2));
      // ---- row j = 0: multiplicity of each peak in the annotation's list ----
      db_zero(row0, cw);
      __syncthreads();
      double se = 0.0;
      for (int64_t idx = beg + threadIdx.x; idx < end; idx += DB_UNROLL * DB_THREADS) {
        int p[DB_UNROLL];
#pragma unroll
        for (int u = 0; u < DB_UNROLL; ++u) {
          const int64_t i = idx + (int64_t)u * DB_THREADS;
          p[u] = i < end ? __ldg(annoidx + i) : -1;
        }
#pragma unroll
        for (int u = 0; u < DB_UNROLL; ++u) {
          if (p[u] >= q0 && p[u] < qhi) {
            const int off = p[u] - (int)q0;
            atomicAdd(row0 + (off >> 2), 1u << (8 * (off & 3)));
            se += __ldg(e + p[u]);
          }
        }
      }
      __syncthreads();
      db_store(row0, reinterpret_cast<uint32_t*>(M + mrow0 * Ppad + q0), cw);
      se = warp_sum<double>(se);
      if (lane == 0) s_red[warp] = se;
      __syncthreads();
      if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < DB_THREADS / 32; ++w) s += s_red[w];
        double* dst = etab + (int64_t)k * I;
        dst[0] = (q0 == 0) ? s : dst[0] + s;
      }
      // ---- rows j = 1..B: background iteration j ----
      for (int j = 1; j <= B; ++j) {
        db_zero(rowj, cw);
        __syncthreads();                           // also orders the s_red reads above
        const int32_t* __restrict__ bgcol = bg + (int64_t)(j - 1) * P;
        double sj = 0.0;
        unsigned long long ov = 0;
        for (int64_t idx = beg + threadIdx.x; idx < end; idx += DB_UNROLL * DB_THREADS) {
          // three dependent gathers per item (list -> background column -> expectation): keep
          // DB_UNROLL independent chains in flight per thread
          int q[DB_UNROLL];
#pragma unroll
          for (int u = 0; u < DB_UNROLL; ++u) {
            const int64_t i = idx + (int64_t)u * DB_THREADS;
            q[u] = i < end ? __ldg(annoidx + i) : -1;
          }
#pragma unroll
          for (int u = 0; u < DB_UNROLL; ++u) {
            if (q[u] >= 0) {
              int v = __ldg(bgcol + q[u]) - base;
              if (v < 0 || v >= P) { atomicOr(flags, DB_FLAG_BADIDX); v = 0; }
              q[u] = (v >= q0 && v < qhi) ? v : -1;
            }
          }
          double ev[DB_UNROLL];
#pragma unroll
          for (int u = 0; u < DB_UNROLL; ++u) ev[u] = q[u] >= 0 ? __ldg(e + q[u]) : 0.0;
#pragma unroll
          for (int u = 0; u < DB_UNROLL; ++u) {
            if (q[u] >= 0) {
              const int off = q[u] - (int)q0;
              atomicAdd(rowj + (off >> 2), 1u << (8 * (off & 3)));
              sj += ev[u];
              ov += (unsigned long long)row0b[off];
            }
          }
        }
        __syncthreads();
        db_store(rowj, reinterpret_cast<uint32_t*>(M + (mrow0 + j) * Ppad + q0), cw);
        sj = warp_sum<double>(sj);
        ov = warp_sum<unsigned long long>(ov);
        if (lane == 0) { s_red[warp] = sj; s_ov[warp] = ov; }
        __syncthreads();
        if (threadIdx.x == 0) {
          double s = 0.0;
          unsigned long long o = 0;
          for (int w = 0; w < DB_THREADS / 32; ++w) { s += s_red[w]; o += s_ov[w]; }
          double* dst = etab + (int64_t)k * I + j;
          *dst = (q0 == 0) ? s : *dst + s;
          ov_total += o;
        }
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const double cnt = (double)(end - beg);
      matches[k] = cnt / (double)P;
      overlap[k] = (end > beg) ? ((double)ov_total / (double)B) / cnt : cv_na_real();
    }
    __syncthreads();
  }
}

// Cd[c][:] for one cell per CTA iteration.  val: validated counts (<= 255 checked by the host).
// Row r of Cd = cell c0 + r (zero row when that is past the last cell); shift selects the u8 digit plane.
__global__ void __launch_bounds__(DB_THREADS, 1)
k_dense_cells(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowidx,
              const uint32_t* __restrict__ val, int N, int c0, int n_rows_padded, int shift, int64_t Ppad, int chunk,
              uint8_t* __restrict__ Cd, uint32_t* __restrict__ work_ctr) {
  extern __shared__ __align__(16) uint32_t db_smem[];
  __shared__ int s_item;
  for (;;) {
    if (threadIdx.x == 0) s_item = (int)atomicAdd(work_ctr, 1u);
    __syncthreads();
    const int r = s_item;
    if (r >= n_rows_padded) break;
    const int c = c0 + r;
    const int64_t beg = c < N ? colptr[c] : 0, end = c < N ? colptr[c + 1] : 0;
    for (int64_t q0 = 0; q0 < Ppad; q0 += chunk) {
      const int cw = (int)min((int64_t)chunk, Ppad - q0) >> 2;
      const int64_t qhi = q0 + ((int64_t)cw << 2);
      db_zero(db_smem, cw);
      __syncthreads();
      for (int64_t x = beg + threadIdx.x; x < end; x += DB_THREADS) {
        const int q = __ldg(rowidx + x);
        if (q >= q0 && q < qhi) {
          const int off = q - (int)q0;
          // (duplicate row indices inside a column still sum, like Matrix does)
          atomicAdd(db_smem + (off >> 2), ((__ldg(val + x) >> shift) & 0xFFu) << (8 * (off & 3)));
        }
      }
      __syncthreads();
      db_store(db_smem, reinterpret_cast<uint32_t*>(Cd + (int64_t)r * Ppad + q0), cw);
      __syncthreads();
    }
  }
}

// The same for long columns (bulk data: every peak of every sample holds a count): work item = (cell,
// range of DC_RANGE peaks), so a dense 500 000-peak column is built by 16 CTAs at once instead of one
// CTA passing over it several times.  Sorted columns (flags4 == 0, checked by K1) are entered by
// binary search; unsorted ones are scanned whole and filtered.
#define DC_RANGE 32768
__global__ void __launch_bounds__(256)
k_dense_cells_split(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowidx,
                    const uint32_t* __restrict__ val, int N, int c0, int n_rows_padded, int shift, int64_t Ppad,
                    const uint32_t* __restrict__ k1_flags, uint8_t* __restrict__ Cd) {
  __shared__ __align__(16) uint32_t sm[DC_RANGE / 4];
  const int n_ranges = (int)((Ppad + DC_RANGE - 1) / DC_RANGE);
  const bool sorted = k1_flags[4] == 0u;
  for (int64_t item = blockIdx.x; item < (int64_t)n_rows_padded * n_ranges; item += gridDim.x) {
    const int r = (int)(item / n_ranges), rg = (int)(item % n_ranges);
    const int c = c0 + r;
    const int64_t q0 = (int64_t)rg * DC_RANGE, q1 = min(Ppad, q0 + DC_RANGE);
    const int nw = (int)(q1 - q0) >> 2;                  // Ppad is a multiple of 128
    for (int i = threadIdx.x; i < (nw >> 2); i += 256) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    int64_t beg = c < N ? colptr[c] : 0, end = c < N ? colptr[c + 1] : 0;
    if (sorted && end > beg) {
      int64_t lo = beg, hi = end;                        // first entry with row >= q0
      while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (__ldg(rowidx + mid) < q0) lo = mid + 1; else hi = mid; }
      beg = lo;
      hi = end;                                          // first entry with row >= q1
      while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (__ldg(rowidx + mid) < q1) lo = mid + 1; else hi = mid; }
      end = lo;
    }
    for (int64_t x = beg + threadIdx.x; x < end; x += 256) {
      const int q = __ldg(rowidx + x);
      if (q >= q0 && q < q1) {
        const int off = q - (int)q0;
        // (duplicate row indices inside a column still sum, like Matrix does)
        atomicAdd(&sm[off >> 2], ((__ldg(val + x) >> shift) & 0xFFu) << (8 * (off & 3)));
      }
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(Cd + (int64_t)r * Ppad + q0);
    for (int i = threadIdx.x; i < (nw >> 2); i += 256) dst[i] = reinterpret_cast<const uint4*>(sm)[i];
    __syncthreads();
  }
}

// cnt[j][q] = #{p : bg[p, j] = q}; bg column-major, so reads are coalesced
__global__ void k_bg_mult_count_cm(const int32_t* __restrict__ bg, int base, int64_t PB, int P,
                                   uint32_t* __restrict__ cnt, uint32_t* __restrict__ flags) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= PB) return;
  int q = bg[i] - base;
  if (q < 0 || q >= P) { atomicOr(flags, DB_FLAG_BADIDX); return; }
  atomicAdd(cnt + (i / P) * P + q, 1u);
}
__global__ void k_u32_max_db(const uint32_t* __restrict__ x, int64_t n, uint32_t* __restrict__ out) {
  uint32_t m = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = max(m, x[i]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}
// largest byte of the j = 0 rows = largest multiplicity of one peak inside one annotation list.
// A byte that wrapped (a peak listed > 255 times) shows up as a byte sum != list length and is
// reported as an impossible multiplicity.
__global__ void __launch_bounds__(256)
k_row0_max(const uint8_t* __restrict__ M, const int64_t* __restrict__ annoptr, int K, int G, int I,
           int64_t Ppad, uint32_t* __restrict__ out) {
  __shared__ unsigned long long s_sum[8];
  const int k = blockIdx.x;
  if (k >= K) return;
  const uint32_t* row = reinterpret_cast<const uint32_t*>(M + ((int64_t)(k / G) * DB_BLOCK_N + (int64_t)(k % G) * I) * Ppad);
  uint32_t m = 0;
  unsigned long long sum = 0;
  for (int64_t i = threadIdx.x; i < (Ppad >> 2); i += blockDim.x) {
    uint32_t w = row[i];
    m = max(m, max(max(w & 0xffu, (w >> 8) & 0xffu), max((w >> 16) & 0xffu, w >> 24)));
    sum += (w & 0xffu) + ((w >> 8) & 0xffu) + ((w >> 16) & 0xffu) + (w >> 24);
  }
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  sum = warp_sum<unsigned long long>(sum);
  if ((threadIdx.x & 31) == 0) {
    if (m) atomicMax(out, m);
    s_sum[threadIdx.x >> 5] = sum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < 8; ++w) t += s_sum[w];
    if (t != (unsigned long long)(annoptr[k + 1] - annoptr[k])) atomicMax(out, 0xFFFFu);
  }
}

// rows of a 256-row group that no (annotation, iteration) owns must read as zero
__global__ void __launch_bounds__(256)
k_zero_pad_rows(uint8_t* __restrict__ M, int K, int G, int I, int64_t Ppad) {
  const int g = blockIdx.x;
  const int nk = min(G, K - g * G);
  const int first = nk * I;
  uint4* base = reinterpret_cast<uint4*>(M + ((int64_t)g * DB_BLOCK_N + first) * Ppad);
  const int64_t n16 = ((int64_t)(DB_BLOCK_N - first) * Ppad) >> 4;
  for (int64_t i = threadIdx.x; i < n16; i += blockDim.x) base[i] = make_uint4(0, 0, 0, 0);
}

// =====================================================================================================
// Gather formulation of the operand rows (default).  With the annotations as bit columns
// Abits[w][p] (bit b of word w = peak p belongs to annotation 32 w + b) and the inverse background
// map of every iteration (sources p of each destination peak q: bg[p, j] = q),
//     M[(k, j), q] = sum over the sources p of (j, q) of bit_k(p),        M[(k, 0), q] = bit_k(q),
// one thread produces 4 consecutive peaks of one iteration for 32 annotations at once: every
// operand byte is written exactly once, coalesced (128 B per warp and row), with no shared-memory
// staging, no zero fill and no atomics on the operand.  The overlap counts
// sum_q M[(k, j), q] * M[(k, 0), q] (R/compute_deviations.R:506) fall out of the same registers;
// the expected-fraction table is a row-gather sum over E2[p][j] = e[bg[p, j]] (:488, :502).
// Lists that repeat a peak (multiplicity > 1 inside one annotation) cannot be bit columns: the
// per-annotation scatter kernel above handles those inputs.
// =====================================================================================================
#define GR_THREADS 256
#define GR_TQ 1024           // peaks per tile: 4 per thread

__global__ void __launch_bounds__(256)
k_anno_bits(const int64_t* __restrict__ annoptr, const int32_t* __restrict__ annoidx, int K, int P,
            uint32_t* __restrict__ Abits, uint32_t* __restrict__ AbT, uint32_t* __restrict__ dflags) {
  const int k = blockIdx.x;
  const uint32_t bit = 1u << (k & 31);
  const int w = k >> 5;
  uint32_t* __restrict__ col = Abits + (int64_t)w * P;
  // peak-major copy [word group of 32][P][32 words]: the 32 words of a peak are one 128-byte line
  uint32_t* __restrict__ colT = AbT + (int64_t)(w >> 5) * P * 32 + (w & 31);
  const int64_t beg = annoptr[k], end = annoptr[k + 1];
  bool dup = false;
  for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
    const int p = annoidx[i];
    const uint32_t old = atomicOr(col + p, bit);
    atomicOr(colT + (int64_t)p * 32, bit);
    dup |= (old & bit) != 0;
  }
  if (dup) atomicOr(dflags + 3, 1u);
}

// invsrc[invptr[j*P + q] ...] = the peaks p with bg[p, j] = q (cursor = the counts, consumed)
__global__ void __launch_bounds__(256)
k_inv_fill(const int32_t* __restrict__ bg, int base, int64_t PB, int P, const uint32_t* __restrict__ invptr,
           uint32_t* __restrict__ cursor, int32_t* __restrict__ invsrc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= PB) return;
  const int q = bg[i] - base;
  if (q < 0 || q >= P) return;                    // flagged by k_bg_mult_count_cm
  const int64_t j = i / P;
  const int64_t idx = j * P + q;
  const uint32_t old = atomicSub(cursor + idx, 1u);
  invsrc[invptr[idx] + old - 1u] = (int32_t)(i - j * P);
}

// One CTA = one tile of 32 annotations (word w) x GR_TQ peaks of iteration j, built in shared
// memory: thread t owns the 4 consecutive peaks 4t .. 4t+3 (one 32-bit column of every row), so
// the byte increments need no atomics and lane t always hits bank t; only the SET bits of a source
// word are visited (annotations are ~7 % dense).  The tile leaves as 16-byte stores, 1 KB per row.
__global__ void __launch_bounds__(GR_THREADS)
k_dense_rows_gather(const uint32_t* __restrict__ Abits, const uint32_t* __restrict__ invptr,
                    const int32_t* __restrict__ invsrc, int K, int P, int B, int G, int64_t Ppad,
                    uint8_t* __restrict__ M) {
  __shared__ __align__(16) uint8_t tile[32 * GR_TQ];
  __shared__ long long s_row[32];                 // byte offset of row (k, j) for the 32 annotations of this word
  const int j = blockIdx.y, w = blockIdx.z;
  const int I = B + 1;
  if (threadIdx.x < 32) {
    const int k = w * 32 + threadIdx.x;
    s_row[threadIdx.x] = k < K ? ((long long)(k / G) * DB_BLOCK_N + (long long)(k % G) * I + j) * Ppad : -1ll;
  }
  {
    uint4* t4 = reinterpret_cast<uint4*>(tile);
#pragma unroll
    for (int i = 0; i < (32 * GR_TQ / 16) / GR_THREADS; ++i) t4[i * GR_THREADS + threadIdx.x] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  const uint32_t* __restrict__ Aw = Abits + (int64_t)w * P;
  const int64_t q0 = (int64_t)blockIdx.x * GR_TQ + threadIdx.x * 4;
  uint8_t* col = tile + threadIdx.x * 4;
  if (j == 0) {
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) {
      uint32_t wd = (q0 + qi < P) ? __ldg(Aw + q0 + qi) : 0u;
      while (wd) {
        const int b = __ffs(wd) - 1;
        wd &= wd - 1;
        col[b * GR_TQ + qi] = 1;
      }
    }
  } else if (q0 < P) {
    const uint32_t* __restrict__ ip = invptr + (int64_t)(j - 1) * P;
    uint32_t sp[5];
#pragma unroll
    for (int qi = 0; qi < 5; ++qi) sp[qi] = __ldg(ip + min(q0 + qi, (int64_t)P));   // ip[P] = start of the next iteration
    // all sources of the 4 peaks in one loop (less divergence than 4 short loops)
    // all sources of the 4 peaks in one loop (less divergence than 4 short loops)
    for (uint32_t s = sp[0]; s < sp[4]; ++s) {
      uint32_t wd = __ldg(Aw + __ldg(invsrc + s));
      const int qi = (s >= sp[1]) + (s >= sp[2]) + (s >= sp[3]);
      while (wd) {
        const int b = __ffs(wd) - 1;
        wd &= wd - 1;
        col[b * GR_TQ + qi] += 1;
      }
    }
  }
  __syncthreads();
  // rows out: 64 x 16 B per row
  for (int i = threadIdx.x; i < 32 * (GR_TQ / 16); i += GR_THREADS) {
    const int r = i / (GR_TQ / 16), c = i % (GR_TQ / 16);
    const long long ro = s_row[r];
    const int64_t q = (int64_t)blockIdx.x * GR_TQ + c * 16;
    if (ro >= 0 && q < Ppad)
      *reinterpret_cast<uint4*>(M + ro + q) = *reinterpret_cast<const uint4*>(tile + r * GR_TQ + c * 16);
  }
}

// Background overlap counts (R/compute_deviations.R:506): ov[k] = sum_j sum_p bit_k(p) * bit_k(bg[p, j]).
// Bit-parallel over the annotations: lane = one membership word (32 annotations) of a group of 32 words,
// so a (peak, iteration) pair costs the warp two 128-byte loads from the peak-major bit matrix -- the
// words of peak p and of its background peak -- instead of 32 scattered 4-byte gathers per word.  The
// per-annotation counts are bit-sliced (7 carry-save planes per lane, flushed into 32 per-lane counters
// every 127 pairs), then summed through shared memory.
#define OV_PAIRS_PER_WARP 2032          // 16 x 127
__global__ void __launch_bounds__(256)
k_overlap_bits(const uint32_t* __restrict__ AbT, const int32_t* __restrict__ bg, int base, int64_t PB, int P, int K,
               unsigned long long* __restrict__ ov_total) {
  __shared__ unsigned int s_ov[32 * 32];
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) s_ov[i] = 0u;
  __syncthreads();
  const uint32_t* __restrict__ A = AbT + (int64_t)g * P * 32 + lane;
  uint32_t acc[32];
#pragma unroll
  for (int b = 0; b < 32; ++b) acc[b] = 0u;
  uint32_t c[7] = {0, 0, 0, 0, 0, 0, 0};
  int in_planes = 0;
  const int64_t i0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * OV_PAIRS_PER_WARP;
  const int64_t i1 = min(PB, i0 + OV_PAIRS_PER_WARP);
  for (int64_t ib = i0; ib < i1; ib += 32) {
    // 32 pairs per round: lane l fetches the background peak of pair ib + l, every lane walks all 32
    const int64_t mine = ib + lane;
    int qm = -1, pm = 0;
    if (mine < i1) {
      qm = __ldg(bg + mine) - base;
      if (qm < 0 || qm >= P) qm = -1;                     // flagged by k_bg_mult_count_cm
      pm = (int)(mine % P);
    }
    const int n = (int)min((int64_t)32, i1 - ib);
#pragma unroll 8
    for (int u = 0; u < 32; ++u) {
      if (u >= n) break;
      const int q = __shfl_sync(0xffffffffu, qm, u);
      const int p = __shfl_sync(0xffffffffu, pm, u);
      uint32_t x = 0u;
      if (q >= 0) x = __ldg(A + (int64_t)p * 32) & __ldg(A + (int64_t)q * 32);
#pragma unroll
      for (int pl = 0; pl < 7; ++pl) {           // ripple-carry add of the one-bit vector x
        const uint32_t carry = c[pl] & x;
        c[pl] ^= x;
        x = carry;
      }
    }
    in_planes += n;
    if (in_planes + 32 > 127 || ib + 32 >= i1) {
#pragma unroll
      for (int b = 0; b < 32; ++b) {
        uint32_t v = 0u;
#pragma unroll
        for (int pl = 0; pl < 7; ++pl) v |= ((c[pl] >> b) & 1u) << pl;
        acc[b] += v;
      }
#pragma unroll
      for (int pl = 0; pl < 7; ++pl) c[pl] = 0u;
      in_planes = 0;
    }
  }
#pragma unroll
  for (int b = 0; b < 32; ++b)
    if (acc[b]) atomicAdd(s_ov + lane * 32 + b, acc[b]);
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
    const int k = (g * 32 + (i >> 5)) * 32 + (i & 31);
    if (k < K && s_ov[i]) atomicAdd(ov_total + k, (unsigned long long)s_ov[i]);
  }
}

// E2[p][0] = e[p], E2[p][j] = e[bg[p, j]]
__global__ void __launch_bounds__(256)
k_e2_build(const int32_t* __restrict__ bg, int base, int P, int B, const double* __restrict__ e,
           double* __restrict__ E2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int I = B + 1;
  if (i >= (int64_t)P * I) return;
  const int j = (int)(i / P);
  const int p = (int)(i - (int64_t)j * P);
  double v;
  if (j == 0) v = e[p];
  else {
    const int q = bg[(int64_t)(j - 1) * P + p] - base;
    v = (q >= 0 && q < P) ? __ldg(e + q) : 0.0;
  }
  E2[(int64_t)p * I + j] = v;
}

// etab[k][j] = sum_{p in set_k} E2[p][j] in a fixed order (warps stride over the list, lanes own
// iterations, warps are merged in order); matches, overlap from the counts of the gather kernel
__global__ void __launch_bounds__(256)
k_etab_rows(const int64_t* __restrict__ annoptr, const int32_t* __restrict__ annoidx,
            const int32_t* __restrict__ order, const double* __restrict__ E2, int K, int P, int B,
            double* __restrict__ etab, uint32_t* __restrict__ work_ctr) {
  extern __shared__ double s_part[];               // [8][I]
  __shared__ int s_item;
  const int I = B + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (;;) {
    if (threadIdx.x == 0) s_item = (int)atomicAdd(work_ctr, 1u);
    __syncthreads();
    const int item = s_item;
    if (item >= K) break;
    const int k = order[item];
    const int64_t beg = annoptr[k], end = annoptr[k + 1];
    for (int j0 = 0; j0 < I; j0 += 32) {
      const int j = j0 + lane;
      if (j < I) {
        // 8 independent (list entry -> E2 row) chains per lane, summed in list order
        double acc = 0.0;
        for (int64_t idx = beg + warp; idx < end; idx += 64) {
          int pp[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) pp[u] = (idx + 8 * u < end) ? __ldg(annoidx + idx + 8 * u) : -1;
          double v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = pp[u] >= 0 ? __ldg(E2 + (int64_t)pp[u] * I + j) : 0.0;
#pragma unroll
          for (int u = 0; u < 8; ++u) acc += v[u];
        }
        s_part[warp * I + j] = acc;
      }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < I; j += blockDim.x) {
      double s = 0.0;
      for (int wv = 0; wv < 8; ++wv) s += s_part[wv * I + j];
      etab[(int64_t)k * I + j] = s;
    }
    __syncthreads();
  }
}

// fractionMatches (:533) and fractionBackgroundOverlap (:506, :530-534) from the overlap counts
__global__ void k_overlap_finish(const int64_t* __restrict__ annoptr, const unsigned long long* __restrict__ ov_total,
                                 int K, int P, int B, double* __restrict__ overlap, double* __restrict__ matches) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const double cnt = (double)(annoptr[k + 1] - annoptr[k]);
  matches[k] = cnt / (double)P;
  overlap[k] = cnt > 0.0 ? ((double)ov_total[k] / (double)B) / cnt : cv_na_real();
}

// Host: operand M + per-annotation tables + the multiplicity counters (dflags[1]: largest
// #{p : bg[p, j] = q}; dflags[2]: largest multiplicity inside one list, 0xFFFF when a byte wrapped;
// dflags[3]: a list repeats a peak -> the gather formulation does not apply, rebuild with scatter = 1).
// Asynchronous; the caller reads the counters after its stream sync.
int cv_dense_build_rows(cv_handle* h, const DensePlan& pl, int scatter, int tables_only) {
  const int64_t P = h->P, K = h->K;
  const int B = h->B, I = B + 1, G = pl.G;
  const int64_t Ppad = pl.Ppad;
  uint32_t* dflags = h->flags.as<uint32_t>() + CV_DFLAGS;
  CV_TRY(cv_ensure(h, h->tmp1, (size_t)P * B * 4));
  CV_CUDA(h, cudaMemsetAsync(h->tmp1.p, 0, (size_t)P * B * 4, h->stream));
  CV_LAUNCH(h, k_bg_mult_count_cm, (unsigned)((P * B + 255) / 256), 256, 0, h->bg_raw.as<int32_t>(),
            h->index_base, P * B, (int)P, h->tmp1.as<uint32_t>(), dflags);
  CV_LAUNCH(h, k_u32_max_db, 1024, 256, 0, h->tmp1.as<uint32_t>(), P * B, dflags + 1);
  CV_CUDA(h, cudaMemsetAsync(h->work_ctr.p, 0, 64, h->stream));
  // rows of a group that no (annotation, iteration) owns must be zero
  if (!tables_only && !pl.fp4 && (G * I < DB_BLOCK_N || K % G))
    CV_LAUNCH(h, k_zero_pad_rows, pl.n_groups, 256, 0, h->dn_M.as<uint8_t>(), (int)K, G, I, Ppad);
  if (!tables_only && (scatter || h->opt_rows_scatter)) {
    const int budget = h->max_smem_optin - 2048;
    const int chunk2 = (int)std::min<int64_t>(Ppad, ((budget / 2) / 128) * 128);   // two row chunks resident
    CV_CUDA(h, cudaFuncSetAttribute(k_dense_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * chunk2));
    CV_LAUNCH(h, k_dense_rows, (unsigned)std::min<int64_t>(K, h->num_sms), DB_THREADS, (size_t)2 * chunk2,
              h->annoptr.as<int64_t>(), h->annoidx.as<int32_t>(), h->bg_raw.as<int32_t>(), h->index_base,
              h->expect.as<double>(), h->order.as<int32_t>(), (int)K, (int)P, B, G, Ppad, chunk2,
              h->dn_M.as<uint8_t>(), h->etab.as<double>(), h->overlap.as<double>(), h->matches.as<double>(),
              h->work_ctr.as<uint32_t>(), dflags);
    CV_LAUNCH(h, k_row0_max, (unsigned)K, 256, 0, h->dn_M.as<uint8_t>(), h->annoptr.as<int64_t>(), (int)K, G, I, Ppad,
              dflags + 2);
    return CV_OK;
  }
  const int64_t Kw = (K + 31) / 32;
  CV_TRY(cv_ensure(h, h->dn_abits, (size_t)Kw * P * 4));
  const int64_t Kg = (Kw + 31) / 32;                                       // groups of 32 membership words
  CV_TRY(cv_ensure(h, h->dn_abitsT, (size_t)Kg * P * 128));
  if (!tables_only) {
    CV_TRY(cv_ensure(h, h->dn_invptr, (size_t)(P * B + 1) * 4));
    CV_TRY(cv_ensure(h, h->dn_invsrc, (size_t)P * B * 4));
  }
  if (tables_only != 2) CV_TRY(cv_ensure(h, h->dn_E2, (size_t)P * I * 8));
  CV_TRY(cv_ensure(h, h->dn_ov, (size_t)K * 8));
  // Two streams: the operand rows (write-bound) on the compute stream; the expected-fraction table
  // (L2-read-bound) and the overlap counts beside them on the auxiliary stream.  The GEMM waits for
  // the table (ev[6]); the overlap / matches vectors are only joined at the end of the call (ev[7]).
  cudaStream_t compute = h->stream;
  CV_CUDA(h, cudaEventRecord(h->ev_aux[0], compute));                    // inputs are on the device
  CV_CUDA(h, cudaMemsetAsync(h->dn_abits.p, 0, (size_t)Kw * P * 4, compute));
  CV_CUDA(h, cudaMemsetAsync(h->dn_abitsT.p, 0, (size_t)Kg * P * 128, compute));
  CV_LAUNCH(h, k_anno_bits, (unsigned)K, 256, 0, h->annoptr.as<int64_t>(), h->annoidx.as<int32_t>(), (int)K, (int)P,
            h->dn_abits.as<uint32_t>(), h->dn_abitsT.as<uint32_t>(), dflags);
  CV_CUDA(h, cudaEventRecord(h->ev_aux[1], compute));                    // bit columns ready
  if (tables_only) {
    // few-cells path (cv_dense_gather.cu): bit columns and tables only, no stacked operand
  } else if (true) {
  // inverse background map: counting sort of (iteration, destination peak)
  CV_TRY(cv_exclusive_scan_u32_async(h, h->tmp1.as<uint32_t>(), h->dn_invptr.as<uint32_t>(), P * B));
  CV_LAUNCH(h, k_inv_fill, (unsigned)((P * B + 255) / 256), 256, 0, h->bg_raw.as<int32_t>(), h->index_base, P * B,
            (int)P, h->dn_invptr.as<uint32_t>(), h->tmp1.as<uint32_t>(), h->dn_invsrc.as<int32_t>());
  if (pl.fp4) {
    CV_TRY(cv_f4_build_rows(h, pl));               // nibble-packed rows in the super-group layout
  } else {
    dim3 grid((unsigned)((Ppad + GR_TQ - 1) / GR_TQ), (unsigned)I, (unsigned)Kw);
    CV_LAUNCH(h, k_dense_rows_gather, grid, GR_THREADS, 0, h->dn_abits.as<uint32_t>(), h->dn_invptr.as<uint32_t>(),
              h->dn_invsrc.as<int32_t>(), (int)K, (int)P, B, G, Ppad, h->dn_M.as<uint8_t>());
  }
  }
  // ---- auxiliary stream ----
  h->stream = h->s_aux;
  int rc = CV_OK;
  do {
    if (cudaStreamWaitEvent(h->s_aux, h->ev_aux[0], 0) != cudaSuccess) { rc = CV_ERR_CUDA; break; }
    if (cudaMemsetAsync(h->work_ctr.as<uint32_t>() + 4, 0, 4, h->s_aux) != cudaSuccess) { rc = CV_ERR_CUDA; break; }
    if (tables_only != 2) {                                                // (2: the few-cells GEMM computes the table itself)
    k_e2_build<<<(unsigned)((P * I + 255) / 256), 256, 0, h->s_aux>>>(h->bg_raw.as<int32_t>(), h->index_base, (int)P, B,
                                                                      h->expect.as<double>(), h->dn_E2.as<double>());
    k_etab_rows<<<(unsigned)std::min<int64_t>(K, (int64_t)h->num_sms * 4), 256, (size_t)8 * I * 8, h->s_aux>>>(
        h->annoptr.as<int64_t>(), h->annoidx.as<int32_t>(), h->order.as<int32_t>(), h->dn_E2.as<double>(), (int)K,
        (int)P, B, h->etab.as<double>(), h->work_ctr.as<uint32_t>() + 4);
    }
    cudaEventRecord(h->ev_aux[2], h->s_aux);                               // expected-fraction table ready
    cudaStreamWaitEvent(h->s_aux, h->ev_aux[1], 0);
    cudaMemsetAsync(h->dn_ov.p, 0, (size_t)K * 8, h->s_aux);
    dim3 grid2((unsigned)((P * B + (int64_t)OV_PAIRS_PER_WARP * 8 - 1) / ((int64_t)OV_PAIRS_PER_WARP * 8)), (unsigned)Kg);
    k_overlap_bits<<<grid2, 256, 0, h->s_aux>>>(h->dn_abitsT.as<uint32_t>(), h->bg_raw.as<int32_t>(), h->index_base,
                                                P * B, (int)P, (int)K, h->dn_ov.as<unsigned long long>());
    k_overlap_finish<<<(unsigned)((K + 127) / 128), 128, 0, h->s_aux>>>(h->annoptr.as<int64_t>(),
                                                                        h->dn_ov.as<unsigned long long>(), (int)K, (int)P,
                                                                        B, h->overlap.as<double>(), h->matches.as<double>());
    cudaEventRecord(h->ev_aux[3], h->s_aux);                               // overlap / matches ready
    h->launches += 4;
    if (cudaGetLastError() != cudaSuccess) rc = CV_ERR_CUDA;
  } while (0);
  h->stream = compute;
  if (rc != CV_OK) CV_FAIL(h, rc, "launch on the auxiliary stream failed");
  if (!tables_only)                                                        // (few-cells path: only its finishing kernel waits)
    CV_CUDA(h, cudaStreamWaitEvent(compute, h->ev_aux[2], 0));             // the GEMM epilogue reads the table
  h->aux_pending = true;
  return CV_OK;
}

// The u8 operand rows again, from the bit columns / inverse background map a gather build left on
// the device (the e2m1 attempt turned out not to fit): no tables, no auxiliary stream.
int cv_dense_rebuild_rows_u8(cv_handle* h, const DensePlan& pl) {
  const int64_t P = h->P, K = h->K;
  const int B = h->B, I = B + 1, G = pl.G;
  const int64_t Kw = (K + 31) / 32;
  if (G * I < DB_BLOCK_N || K % G)
    CV_LAUNCH(h, k_zero_pad_rows, pl.n_groups, 256, 0, h->dn_M.as<uint8_t>(), (int)K, G, I, pl.Ppad);
  dim3 grid((unsigned)((pl.Ppad + GR_TQ - 1) / GR_TQ), (unsigned)I, (unsigned)Kw);
  CV_LAUNCH(h, k_dense_rows_gather, grid, GR_THREADS, 0, h->dn_abits.as<uint32_t>(), h->dn_invptr.as<uint32_t>(),
            h->dn_invsrc.as<int32_t>(), (int)K, (int)P, B, G, pl.Ppad, h->dn_M.as<uint8_t>());
  return CV_OK;
}

// Cd rows [0, rows) <- cells [c0, c0 + rows), digit plane `shift` / 8.  Asynchronous.
int cv_dense_build_cells(cv_handle* h, const DensePlan& pl, int64_t c0, int64_t rows, int shift) {
  if (h->N > 0 && h->nnz / h->N > 16384) {
    // long columns (bulk data): (cell, peak range) items -- config 4: 3.56 ms -> 0.2 ms per plane
    const int64_t items = rows * ((pl.Ppad + DC_RANGE - 1) / DC_RANGE);
    CV_LAUNCH(h, k_dense_cells_split, (unsigned)std::min<int64_t>(items, (int64_t)h->num_sms * 6), 256, 0,
              h->colptr.as<int64_t>(), h->rowidx.as<int32_t>(), h->val.as<uint32_t>(), (int)h->N, (int)c0, (int)rows, shift,
              pl.Ppad, h->flags.as<uint32_t>(), h->dn_C.as<uint8_t>());
    return CV_OK;
  }
  const int budget = h->max_smem_optin - 2048;
  const int chunk1 = (int)std::min<int64_t>(pl.Ppad, (budget / 128) * 128);
  CV_CUDA(h, cudaMemsetAsync(h->work_ctr.as<uint32_t>() + 1, 0, 4, h->stream));
  CV_CUDA(h, cudaFuncSetAttribute(k_dense_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, chunk1));
  CV_LAUNCH(h, k_dense_cells, (unsigned)std::min<int64_t>(rows, (int64_t)h->num_sms), DB_THREADS, (size_t)chunk1,
            h->colptr.as<int64_t>(), h->rowidx.as<int32_t>(), h->val.as<uint32_t>(), (int)h->N, (int)c0, (int)rows,
            shift, pl.Ppad, chunk1, h->dn_C.as<uint8_t>(), h->work_ctr.as<uint32_t>() + 1);
  return CV_OK;
}
