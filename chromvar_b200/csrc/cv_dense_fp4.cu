// K5a, e2m1 variant: the same contraction as cv_dense.cu on tcgen05.mma.kind::mxf4 -- 4-bit e2m1
// operands with unit ue8m0 block scales, f32 accumulators -- which retires twice the MACs per cycle
// of kind::i8 (tools/fp4_probe.cu: 128 cycles per M128 N256 K64 against K32) and whose f32
// accumulation is EXACT for integer-valued products while every sum stays below 2^24 (same probe).
//
// e2m1 holds {0, 0.5, 1, 1.5, 2, 3, 4, 6}; counts and multiplicities are integers, so the exact
// alphabet is S = {0, 1, 2, 3, 4, 6}.  A value outside S is split into terms from S
// (5 = 4 + 1, 7 = 6 + 1, 13 = 6 + 6 + 1, ...): term 0 stays in the peak's own column, the other
// terms go to OVERFLOW COLUMNS appended behind the P peak columns of both operands.  For a peak q
// whose largest count in this cell chunk needs S_q terms and whose largest operand entry
// M[(k, j), q] needs T_q terms, the columns (s, t) != (0, 0), s < S_q, t < T_q hold term s of the
// count (cells operand) and term t of the multiplicity (annotation operand), so that
//     sum over the columns of q of  Cd * M  =  (sum_s c_s) * (sum_t m_t)  =  C[c, q] * M[(k, j), q].
// Single-cell counts are almost all 1-4 (config 2: 0.43 % of the nonzeros outside S, ~2 800 overflow
// columns per 2 560-cell chunk on 100 000 peaks), so the contraction grows by a few percent in K
// and halves in time.  Everything stays exact integer arithmetic; whatever cannot be represented
// (overflow beyond the reserved columns, sums >= 2^24, repeated peaks in a list, duplicate row
// indices in a column) is detected and the call falls back to the u8 kernel.
//
// Tensor memory: mxf4 needs the block scales in TMEM, so two 256-column accumulator stages no longer
// fit.  The stacked rows are therefore grouped in SUPER-GROUPS of 464 rows = a 256-row sub-tile
// (floor(256 / (B+1)) annotations) + a 208-row sub-tile (floor(208 / (B+1)) annotations): 5 + 4 for
// the default 50 iterations, 459 of 464 columns used.  Sub-tile A accumulates in TMEM columns
// [0, 256), sub-tile B in [256, 464), the scale bytes (all 0x7F = 2^0) live in [464, 512); the
// epilogue of one sub-tile overlaps the MMAs of the other exactly as in the u8 kernel.
#include <cuda.h>

#include <algorithm>

#include "cv_dense_dev.cuh"

#define F4_KPEAKS 256                    // peaks per 128-byte k-block
#define F4_NA 256
#define F4_NB 208
#define F4_SG_ROWS (F4_NA + F4_NB)
#define F4_SF_COL 464                    // first TMEM column of the scale bytes
#define F4_SMEM_BYTES DN_SMEM2_BYTES
#define F4_SYNC_SLOTS 65536              // u32 words of f4_sync: one per (wave, sub-tile) + the last = "lockstep given up"
#define F4_THREADS 384                   // warps 0-3: TMA / MMA / TMEM allocation, warps 4-11: epilogue (two per TMEM lane quarter)
enum { F4_FAIL_CAPACITY = 1u, F4_FAIL_DUPLICATE = 2u, F4_FAIL_NIBBLE = 4u };

// ---- terms over S = {1, 2, 3, 4, 6} ---------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t f4_nterms(uint32_t v) {
  if (v <= 4u || v == 6u) return 1u;
  const uint32_t n6 = v / 6u, r = v - 6u * n6;
  return n6 + (r == 0u ? 0u : (r == 5u ? 2u : 1u));
}
__host__ __device__ __forceinline__ uint32_t f4_term(uint32_t v, uint32_t t) {
  const uint32_t n6 = v / 6u, r = v - 6u * n6;
  if (t < n6) return 6u;
  t -= n6;
  if (r == 5u) return t == 0u ? 4u : (t == 1u ? 1u : 0u);
  return t == 0u ? r : 0u;
}
// e2m1 code of a value of S (0 -> 0, 1 -> 2, 2 -> 4, 3 -> 5, 4 -> 6, 6 -> 7)
__host__ __device__ __forceinline__ uint32_t f4_code(uint32_t v) { return (0x07065420u >> (4u * v)) & 0xFu; }
// e2m1 code of term 0 of any value (5 -> 4, >= 6 -> 6)
__host__ __device__ __forceinline__ uint32_t f4_code0(uint32_t v) { return v >= 7u ? 7u : (0x77665420u >> (4u * v)) & 0xFu; }

// row of (annotation k, iteration j) in the super-group layout
__host__ __device__ __forceinline__ int64_t f4_row(int k, int j, int Ga, int Gb, int I) {
  const int sg = k / (Ga + Gb), r = k - sg * (Ga + Gb);
  return (int64_t)sg * F4_SG_ROWS + (r < Ga ? r * I : F4_NA + (r - Ga) * I) + j;
}

// ---- operand M4: pad rows ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_f4_zero_pad(uint8_t* __restrict__ M4, int K, int Ga, int Gb, int I, int64_t rowb) {
  const int sg = blockIdx.x, half = blockIdx.y;
  const int k0 = sg * (Ga + Gb) + (half ? Ga : 0);
  const int nk = max(0, min(half ? Gb : Ga, K - k0));
  const int64_t first = (int64_t)sg * F4_SG_ROWS + (half ? F4_NA : 0) + (int64_t)nk * I;
  const int64_t end = (int64_t)sg * F4_SG_ROWS + (half ? F4_SG_ROWS : F4_NA);
  uint4* base = reinterpret_cast<uint4*>(M4 + first * rowb);
  const int64_t n16 = ((end - first) * rowb) >> 4;
  for (int64_t i = (int64_t)blockIdx.z * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.z * blockDim.x)
    base[i] = make_uint4(0, 0, 0, 0);
}

// ---- operand M4: the P peak columns (gather formulation of cv_dense_build.cu, nibble-packed) ----------------
// One CTA = 32 annotations (word w) x 1024 peaks of iteration j.  The tile lives in shared memory as
// 4-bit COUNTS: thread t owns peaks 4t .. 4t+3 = one 16-bit word of every row (no atomics, no bank
// conflicts), visits only the set bits of each source's membership word, and the tile leaves as
// 16-byte stores after a SWAR map count -> e2m1 code of term 0 (0,1,2,3 -> 0,2,4,5 without carries;
// larger counts take the slow path and raise T_q).  A count that would pass 15 is flagged
// (F4_FAIL_NIBBLE): such an operand goes to the u8 kernel.  The kernel is instruction-bound, hence
// the packed tile: a quarter of the instructions of the byte tile it replaced.
#define F4_TQ 1024
__device__ __forceinline__ uint32_t f4_codes_of_counts(uint32_t x, int64_t q, uint32_t* __restrict__ Tq) {
  if ((x & 0xCCCCCCCCu) == 0u) return (x << 1) - (x & (x >> 1) & 0x11111111u);       // every count <= 3
  uint32_t out = 0u;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const uint32_t v = (x >> (4 * e)) & 0xFu;
    if (v > 4u && v != 6u) atomicMax(Tq + q + e, f4_nterms(v));
    out |= f4_code0(v) << (4 * e);
  }
  return out;
}
__global__ void __launch_bounds__(256)
k_f4_rows_main(const uint32_t* __restrict__ Abits, const uint32_t* __restrict__ invptr,
               const int32_t* __restrict__ invsrc, int K, int P, int B, int Ga, int Gb, int64_t rowb, int64_t Ppad4,
               uint8_t* __restrict__ M4, uint32_t* __restrict__ Tq, uint32_t* __restrict__ dflags) {
  __shared__ __align__(16) uint16_t tile[32 * (F4_TQ / 4)];
  __shared__ long long s_row[32];
  const int j = blockIdx.y, w = blockIdx.z;
  const int I = B + 1;
  if (threadIdx.x < 32) {
    const int k = w * 32 + threadIdx.x;
    s_row[threadIdx.x] = k < K ? f4_row(k, j, Ga, Gb, I) * rowb : -1ll;
  }
  {
    uint4* t4 = reinterpret_cast<uint4*>(tile);
#pragma unroll
    for (int i = 0; i < (32 * (F4_TQ / 4) * 2 / 16) / 256; ++i) t4[i * 256 + threadIdx.x] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  const uint32_t* __restrict__ Aw = Abits + (int64_t)w * P;
  const int64_t q0 = (int64_t)blockIdx.x * F4_TQ + threadIdx.x * 4;
  uint16_t* col = tile + threadIdx.x;
  if (j == 0) {
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) {
      uint32_t wd = (q0 + qi < P) ? __ldg(Aw + q0 + qi) : 0u;
      while (wd) {
        const int b = __ffs(wd) - 1;
        wd &= wd - 1;
        col[b * (F4_TQ / 4)] |= (uint16_t)(1u << (4 * qi));
      }
    }
  } else if (q0 < P) {
    const uint32_t* __restrict__ ip = invptr + (int64_t)(j - 1) * P;
    uint32_t sp[5];
#pragma unroll
    for (int qi = 0; qi < 5; ++qi) sp[qi] = __ldg(ip + min(q0 + qi, (int64_t)P));
    bool sat = false;
    for (uint32_t s = sp[0]; s < sp[4]; ++s) {
      uint32_t wd = __ldg(Aw + __ldg(invsrc + s));
      const int sh = 4 * ((s >= sp[1]) + (s >= sp[2]) + (s >= sp[3]));
      while (wd) {
        const int b = __ffs(wd) - 1;
        wd &= wd - 1;
        const uint32_t old = col[b * (F4_TQ / 4)];
        sat |= ((old >> sh) & 0xFu) == 0xFu;
        col[b * (F4_TQ / 4)] = (uint16_t)(old + (1u << sh));
      }
    }
    if (sat) atomicOr(dflags + 5, F4_FAIL_NIBBLE);
  }
  __syncthreads();
  // rows out: 16 bytes = 32 peaks per store, 512 bytes per row
  for (int i = threadIdx.x; i < 32 * (F4_TQ / 32); i += 256) {
    const int r = i / (F4_TQ / 32), c = i % (F4_TQ / 32);
    const long long ro = s_row[r];
    const int64_t q = (int64_t)blockIdx.x * F4_TQ + c * 32;
    if (ro < 0 || q >= Ppad4) continue;
    uint4 x = *reinterpret_cast<const uint4*>(tile + r * (F4_TQ / 4) + c * 8);
    x.x = f4_codes_of_counts(x.x, q, Tq);
    x.y = f4_codes_of_counts(x.y, q + 8, Tq);
    x.z = f4_codes_of_counts(x.z, q + 16, Tq);
    x.w = f4_codes_of_counts(x.w, q + 24, Tq);
    *reinterpret_cast<uint4*>(M4 + ro + (q >> 1)) = x;
  }
}

// overflow columns the operand alone needs (sum of T_q - 1): when they already fill the reserve the
// caller switches to the u8 kernel before any cell is touched (dense k-mer annotations: nearly every
// peak meets a multiplicity of 5 in one of the K x B rows)
__global__ void __launch_bounds__(256)
k_f4_tq_total(const uint32_t* __restrict__ Tq, int P, uint32_t* __restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t v = (q < P && Tq[q] > 1u) ? Tq[q] - 1u : 0u;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, v);
}

// ---- per cell chunk: which peaks need overflow columns ---------------------------------------------------------
__global__ void __launch_bounds__(256)
k_f4_terms(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowidx, const uint32_t* __restrict__ val,
           int64_t c0, int64_t c1, uint32_t* __restrict__ Sq) {
  const int64_t e0 = colptr[c0], e1 = colptr[c1];
  for (int64_t e = e0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < e1; e += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t v = val[e];
    if (v > 4u && v != 6u) atomicMax(Sq + rowidx[e], f4_nterms(v));
  }
}
__global__ void __launch_bounds__(256)
k_f4_ncols(const uint32_t* __restrict__ Sq, const uint32_t* __restrict__ Tq, int P, uint32_t* __restrict__ ncol) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= P) return;
  const unsigned long long n = (unsigned long long)max(Sq[q], 1u) * max(Tq[q], 1u) - 1ull;
  ncol[q] = n > 0x3FFFFFFFull ? 0x3FFFFFFFu : (uint32_t)n;
}
// info[0] = 256-peak k-blocks of this chunk (peak columns + overflow columns), info[1] = overflow columns
__global__ void k_f4_info(const uint32_t* __restrict__ off, int P, int64_t Ppad4, uint32_t Hcap, uint32_t* __restrict__ info,
                          uint32_t* __restrict__ dflags) {
  uint32_t h = off[P];
  if (h > Hcap) { atomicOr(dflags + 5, F4_FAIL_CAPACITY); h = 0u; }
  info[0] = (uint32_t)(Ppad4 / F4_KPEAKS) + (h + F4_KPEAKS - 1) / F4_KPEAKS;
  info[1] = h;
  atomicAdd(dflags + 6, h);
}
__global__ void __launch_bounds__(256)
k_f4_collist(const uint32_t* __restrict__ off, const uint32_t* __restrict__ Tq, const uint32_t* __restrict__ info, int P,
             int32_t* __restrict__ colq, uint32_t* __restrict__ colt) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= P || info[1] == 0u) return;
  const uint32_t o = off[q], n = off[q + 1] - o;
  if (!n) return;
  const uint32_t T = max(Tq[q], 1u);
  for (uint32_t i = 0; i < n; ++i) {
    colq[o + i] = q;
    colt[o + i] = (i + 1u) % T;                    // column (s, t) sits at s * T + t - 1
  }
}

// ---- operand M4: the overflow columns of this chunk (term t of the multiplicity of peak colq[x]) ----------------
// Same construction as k_f4_rows_main: 32 annotations x 1024 overflow columns of iteration j per CTA
// iteration, 4-bit counts, thread t owns columns 4t .. 4t+3.  Column x holds term colt[x] of the
// count; terms past the first are zero unless the count is outside S.
#define F4_OV_COLS 1024
__global__ void __launch_bounds__(256)
k_f4_rows_ov(const uint32_t* __restrict__ Abits, const uint32_t* __restrict__ invptr, const int32_t* __restrict__ invsrc,
             const int32_t* __restrict__ colq, const uint32_t* __restrict__ colt, const uint32_t* __restrict__ info,
             int K, int P, int B, int Ga, int Gb, int64_t rowb, int64_t Pb, uint8_t* __restrict__ M4,
             uint32_t* __restrict__ dflags) {
  __shared__ __align__(16) uint16_t tile[32 * (F4_OV_COLS / 4)];
  __shared__ uint8_t s_t[F4_OV_COLS];
  __shared__ uint32_t s_first[F4_OV_COLS / 8];      // nibble mask of the columns holding term 0
  __shared__ long long s_row[32];
  const uint32_t h = info[1];
  const int j = blockIdx.y, w = blockIdx.z;
  const int I = B + 1;
  if (threadIdx.x < 32) {
    const int k = w * 32 + threadIdx.x;
    s_row[threadIdx.x] = k < K ? f4_row(k, j, Ga, Gb, I) * rowb : -1ll;
  }
  const uint32_t* __restrict__ Aw = Abits + (int64_t)w * P;
  const uint32_t* __restrict__ ip = invptr + (int64_t)(j > 0 ? j - 1 : 0) * P;
  const uint32_t hpad = (h + F4_KPEAKS - 1) / F4_KPEAKS * F4_KPEAKS;     // columns past the last used k-block are never read
  for (uint32_t x0 = blockIdx.x * F4_OV_COLS; x0 < hpad; x0 += gridDim.x * F4_OV_COLS) {
    __syncthreads();
    {
      uint4* t4 = reinterpret_cast<uint4*>(tile);
#pragma unroll
      for (int i = 0; i < (32 * (F4_OV_COLS / 4) * 2 / 16) / 256; ++i) t4[i * 256 + threadIdx.x] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    uint16_t* col = tile + threadIdx.x;
    bool sat = false;
    uint32_t tt4 = 0u, first = 0u;
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) {
      const uint32_t x = x0 + 4 * threadIdx.x + qi;
      if (x >= h) { first |= 0xFu << (4 * qi); continue; }          // zero columns: any map gives 0
      const int q = colq[x];
      const uint32_t tt = min(colt[x], 255u);
      tt4 |= tt << (8 * qi);
      if (tt == 0u) first |= 0xFu << (4 * qi);
      const int sh = 4 * qi;
      if (j == 0) {
        uint32_t wd = __ldg(Aw + q);
        while (wd) {
          const int b = __ffs(wd) - 1;
          wd &= wd - 1;
          col[b * (F4_OV_COLS / 4)] |= (uint16_t)(1u << sh);
        }
      } else {
        const uint32_t s1 = __ldg(ip + q + 1);
        for (uint32_t s = __ldg(ip + q); s < s1; ++s) {
          uint32_t wd = __ldg(Aw + __ldg(invsrc + s));
          while (wd) {
            const int b = __ffs(wd) - 1;
            wd &= wd - 1;
            const uint32_t old = col[b * (F4_OV_COLS / 4)];
            sat |= ((old >> sh) & 0xFu) == 0xFu;
            col[b * (F4_OV_COLS / 4)] = (uint16_t)(old + (1u << sh));
          }
        }
      }
    }
    if (sat) atomicOr(dflags + 5, F4_FAIL_NIBBLE);
    reinterpret_cast<uint32_t*>(s_t)[threadIdx.x] = tt4;
    reinterpret_cast<uint16_t*>(s_first)[threadIdx.x] = (uint16_t)first;
    __syncthreads();
    // rows out: 8 columns = one 32-bit word per item, 512 bytes per row
    for (int i = threadIdx.x; i < 32 * (F4_OV_COLS / 8); i += 256) {
      const int r = i / (F4_OV_COLS / 8), cw = i % (F4_OV_COLS / 8);
      const long long ro = s_row[r];
      if (ro < 0 || x0 + 8u * cw >= hpad) continue;
      const uint32_t x = reinterpret_cast<const uint32_t*>(tile + r * (F4_OV_COLS / 4))[cw];
      uint32_t out = 0u;
      if ((x & 0xCCCCCCCCu) == 0u) {
        out = ((x << 1) - (x & (x >> 1) & 0x11111111u)) & s_first[cw];      // counts <= 3: term 0 only
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const uint32_t v = (x >> (4 * e)) & 0xFu;
          if (v) out |= f4_code(f4_term(v, s_t[8 * cw + e])) << (4 * e);
        }
      }
      *reinterpret_cast<uint32_t*>(M4 + ro + Pb + (x0 >> 1) + 4 * cw) = out;
    }
  }
}

// ---- operand Cd4: one cell per CTA iteration, nibbles set with shared-memory atomicOr ---------------------------
// Row r of Cd4 = cell c0 + r (zero row past the last cell).  The row (peak columns + used overflow
// columns) is built in pieces of `piece` bytes.
__global__ void __launch_bounds__(512, 3)
k_f4_cells(const int64_t* __restrict__ colptr, const int32_t* __restrict__ rowidx, const uint32_t* __restrict__ val,
           int N, int c0, int n_rows_padded, const uint32_t* __restrict__ Tq, const uint32_t* __restrict__ Sq,
           const uint32_t* __restrict__ off, const uint32_t* __restrict__ info, int64_t Pb, int64_t rowb, int piece,
           uint8_t* __restrict__ Cd4, uint32_t* __restrict__ work_ctr, uint32_t* __restrict__ dflags) {
  extern __shared__ __align__(16) uint32_t f4_smem[];
  __shared__ int s_item;
  const int64_t used = Pb + (int64_t)((info[1] + F4_KPEAKS - 1) / F4_KPEAKS) * (F4_KPEAKS / 2);   // bytes of a row in use
  for (;;) {
    if (threadIdx.x == 0) s_item = (int)atomicAdd(work_ctr, 1u);
    __syncthreads();
    const int r = s_item;
    if (r >= n_rows_padded) break;
    const int c = c0 + r;
    const int64_t beg = c < N ? colptr[c] : 0, end = c < N ? colptr[c + 1] : 0;
    for (int64_t b0 = 0; b0 < used; b0 += piece) {
      const int nb = (int)min((int64_t)piece, used - b0);          // multiple of 128
      for (int i = threadIdx.x; i < (nb >> 4); i += blockDim.x) reinterpret_cast<uint4*>(f4_smem)[i] = make_uint4(0, 0, 0, 0);
      __syncthreads();
      bool dup = false;
      for (int64_t x = beg + threadIdx.x; x < end; x += blockDim.x) {
        const int q = __ldg(rowidx + x);
        const uint32_t v = __ldg(val + x);
        {
          const int64_t byte = (int64_t)(q >> 1) - b0;
          if (byte >= 0 && byte < nb) {
            const int sh = ((int)(byte & 3) << 3) + ((q & 1) << 2);
            const uint32_t old = atomicOr(f4_smem + (byte >> 2), f4_code0(v) << sh);
            dup |= ((old >> sh) & 0xFu) != 0u;
          }
        }
        const uint32_t o = __ldg(off + q), n = __ldg(off + q + 1) - o;
        if (n) {
          const uint32_t T = max(__ldg(Tq + q), 1u);
          const uint32_t nv = f4_nterms(v);
          for (uint32_t s = 0; s < nv; ++s) {
            const uint32_t code = f4_code(f4_term(v, s));
            for (uint32_t t = (s == 0u ? 1u : 0u); t < T; ++t) {
              const uint32_t colx = o + s * T + t - 1u;              // < o + n by construction (s < S_q)
              const int64_t byte = Pb + (int64_t)(colx >> 1) - b0;
              if (byte >= 0 && byte < nb) {
                const int sh = ((int)(byte & 3) << 3) + ((colx & 1u) << 2);
                atomicOr(f4_smem + (byte >> 2), code << sh);
              }
            }
          }
        }
      }
      if (dup) atomicOr(dflags + 5, F4_FAIL_DUPLICATE);
      __syncthreads();
      for (int i = threadIdx.x; i < (nb >> 4); i += blockDim.x)
        reinterpret_cast<uint4*>(Cd4 + (int64_t)r * rowb + b0)[i] = reinterpret_cast<const uint4*>(f4_smem)[i];
      __syncthreads();
    }
  }
}

// ---- the GEMM ----------------------------------------------------------------------------------------------------
// D[tmem of both CTAs] (+)= A[256 x 64 e2m1: 128 rows from each CTA] * B[64 x N: N / 2 rows from each CTA]
__device__ __forceinline__ void umma_mxf4_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t tmem_sfa, uint32_t tmem_sfb, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
      : "memory");
}
// block-scaled instruction descriptor: A = B = e2m1, ue8m0 scales, both K-major, dense, K = 64
__host__ __device__ constexpr uint32_t make_idesc_mxf4(int M, int N) {
  return (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (1u << 23) | ((uint32_t)(M >> 4) << 24);
}

// Work item = (256-cell unit, super-group): sub-tile A (256 stacked rows) then sub-tile B (208).
// Same roles and barriers as k_dense_contract_pair; the accumulator stage is the sub-tile.
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(F4_THREADS, 1)
k_dense_contract_fp4(const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_ma,
                     const __grid_constant__ CUtensorMap map_mb, const DenseParams prm) {
  extern __shared__ unsigned char dn_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dn_smem_raw) + 1023) &
                                                         ~(uintptr_t)1023);
  unsigned char* smem_a = smem;                                   // [STAGES2][16 KB] own cells
  unsigned char* smem_b = smem + DN_STAGES2 * DN_A_BYTES;         // [STAGES2][16 KB] own half of the sub-tile's rows
  double* s_winv = reinterpret_cast<double*>(smem + DN_STAGES2 * DN_STAGE2_BYTES);
  double* s_wred = s_winv + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DN_STAGES2 * DN_STAGE2_BYTES + 4096);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + DN_STAGES2;
  uint64_t* tfull_bar = bars + 2 * DN_STAGES2;        // [2] per sub-tile
  uint64_t* tempty_bar = bars + 2 * DN_STAGES2 + 2;   // [2] leader's: 512 arrivals (256 epilogue threads of each CTA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * DN_STAGES2 + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_units = (prm.n_cell_blocks + 1) >> 1;
  const int n_kb = (int)*prm.n_kblocks_dev;
  const int Gsg = prm.Ga + prm.Gb;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ma) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_mb) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < DN_STAGES2; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar + s, 1); mbar_init(tempty_bar + s, 2 * (F4_THREADS - 128)); }
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
  if (warp >= 4 && warp < 8) {
    // unit block scales (ue8m0 0x7F = 2^0) in every byte of the scale columns of this CTA's lanes
    const uint32_t la = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (int c = F4_SF_COL; c < 512; ++c)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(la + (uint32_t)c), "r"(0x7F7F7F7Fu) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own cells + own half of the sub-tile's stacked rows =====
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int64_t item = pair_id; item < prm.n_tiles; item += n_pairs) {
        int unit, sg;
        dn_decode_tile(prm.band, prm.n_sg, n_units, item, unit, sg);
        const int yc = (2 * unit + (int)rank) * DN_BLOCK_M;
        for (int half = 0; half < 2; ++half) {
          if (sg * Gsg + (half ? prm.Ga : 0) >= prm.K) {
            // the last super-group may have no second sub-tile: still counted by the others' rendezvous
            if (prm.sync_ctr) atomicAdd(prm.sync_ctr + (((item - pair_id) / n_pairs) * 2 + half), 1u);
            continue;
          }
          const int nh = half ? F4_NB / 2 : F4_NA / 2;                 // stacked rows per CTA
          const int ym = sg * F4_SG_ROWS + (half ? F4_NA : 0) + (int)rank * nh;
          const uint32_t bytes = 2u * (uint32_t)(DN_A_BYTES + nh * DN_BLOCK_K);
          if (prm.sync_ctr) {
            // all producers start a sub-tile together: the CTAs that share an operand stream then ask
            // for the same k-block within a few microseconds and L2 serves all but the first
            // (without it the pairs drift apart by more than L2 holds and re-read from DRAM).
            // sync_lag = 1: a producer only waits for everybody to have STARTED the previous sub-tile, so
            // nobody stalls for the slowest pair at every boundary while the drift stays below one sub-tile.
            const int64_t wave = (item - pair_id) / n_pairs;
            const int64_t idx = wave * 2 + half;
            atomicAdd(prm.sync_ctr + idx, 1u);
            const int64_t widx = idx - prm.sync_lag;
            if (widx >= 0) {
              const int64_t left = prm.n_tiles - (widx >> 1) * n_pairs;
              const uint32_t target = 2u * (uint32_t)(left < n_pairs ? left : n_pairs);
              uint32_t* ctr = prm.sync_ctr + widx;
              uint32_t* gave_up = prm.sync_ctr + F4_SYNC_SLOTS - 1;
              uint32_t seen, off = 0u;
              int spins = 0;
              do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
                if (seen >= target) break;
                // The rendezvous assumes every CTA pair of the wave is resident.  When something else holds
                // SMs (another engine's GEMM on the same GPU, a long auxiliary kernel) some pairs have not
                // started yet and would be waited for forever by pairs that in turn keep their SMs: after
                // ~2 ms of waiting the launch gives the lockstep up for good (a flag word all producers
                // watch) and every pair runs free -- slower through L2, never stuck.
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(off) : "l"(gave_up) : "memory");
                if (off) break;
                __nanosleep(64);
                if (++spins > 16384) { atomicExch(gave_up, 1u); break; }
              } while (true);
            }
          }
          for (int kb = 0; kb < n_kb; ++kb) {
            mbar_wait_relaxed(empty_bar + s, ph ^ 1u, 32);
            if (rank == 0) mbar_expect_tx(full_bar + s, bytes);
            const uint32_t fb = mapa_u32(smem_u32(full_bar + s), 0);
            tma_load_2d_pair(&map_c, fb, smem_a + s * DN_A_BYTES, kb * DN_BLOCK_K, yc);
            tma_load_2d_pair(half ? &map_mb : &map_ma, fb, smem_b + s * DN_A_BYTES, kb * DN_BLOCK_K, ym);
            if (++s == DN_STAGES2) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA; whole warp walks the pipeline, one elected lane issues) =====
    if (rank == 0) {
      int s = 0;
      uint32_t ph = 0;
      uint32_t par = 0u;                                  // bit `half` = parity of that sub-tile's uses
      const uint32_t sfa = tmem_base + (uint32_t)F4_SF_COL, sfb = tmem_base + (uint32_t)F4_SF_COL + 16u;
      for (int64_t item = pair_id; item < prm.n_tiles; item += n_pairs) {
        int unit, sg;
        dn_decode_tile(prm.band, prm.n_sg, n_units, item, unit, sg);
        for (int half = 0; half < 2; ++half) {
          if (sg * Gsg + (half ? prm.Ga : 0) >= prm.K) continue;
          const uint32_t idesc = half ? make_idesc_mxf4(2 * DN_BLOCK_M, F4_NB) : make_idesc_mxf4(2 * DN_BLOCK_M, F4_NA);
          mbar_wait(tempty_bar + half, ((par >> half) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(half * F4_NA);
          for (int kb = 0; kb < n_kb; ++kb) {
            mbar_wait(full_bar + s, ph);
            tc_fence_after();
            const uint64_t da = make_smem_desc(smem_u32(smem_a + s * DN_A_BYTES));
            const uint64_t db = make_smem_desc(smem_u32(smem_b + s * DN_A_BYTES));
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_mxf4_pair(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, sfa, sfb, (kb | k) ? 1u : 0u);
              tc_commit_pair(empty_bar + s, 3);
            }
            __syncwarp();
            if (++s == DN_STAGES2) { s = 0; ph ^= 1u; }
          }
          if (elect_one()) tc_commit_pair(tfull_bar + half, 3);
          __syncwarp();
          par ^= 1u << half;
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs): one thread = one cell (TMEM lane of this CTA) x half of the sub-tile's
    // annotations -- eight warps, two per lane quarter, so the two sweeps of the fixed-point form
    // finish well inside the other sub-tile's MMAs =====
    const int q4 = warp & 3;
    const int set = (warp - 4) >> 2;                     // 0: first half of the annotations, 1: the rest
    const int et = threadIdx.x - 128;
    constexpr int NT = F4_THREADS - 128;
    uint32_t par = 0u;
    for (int64_t item = pair_id; item < prm.n_tiles; item += n_pairs) {
      int unit, sg;
      dn_decode_tile(prm.band, prm.n_sg, n_units, item, unit, sg);
      const int cb = 2 * unit + (int)rank;
      for (int half = 0; half < 2; ++half) {
        const int k0 = sg * Gsg + (half ? prm.Ga : 0);
        if (k0 >= prm.K) continue;
        const int nk = min(half ? prm.Gb : prm.Ga, prm.K - k0);
        dn_epi_stage_tables<MODE, NT>(prm, k0, nk, s_winv, et);
        mbar_wait_relaxed(tfull_bar + half, (par >> half) & 1u, 256);
        tc_fence_after();
        const int a0 = set ? (nk + 1) >> 1 : 0, a1 = set ? nk : (nk + 1) >> 1;
        if (a1 > a0) {
          const int I = prm.B + 1;
          const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(half * F4_NA + a0 * I);
          const double* tab = MODE == DN_EPI_FUSED32
                                  ? reinterpret_cast<const double*>(reinterpret_cast<const uint32_t*>(s_winv) + a0 * I)
                                  : s_winv + a0 * I;
          dn_epi_consume<MODE, true>(prm, cb, k0 + a0, a1 - a0, taddr, tab, s_wred + a0 * 16, q4, lane);
        }
        tc_fence_before();
        mbar_arrive_cluster(mapa_u32(smem_u32(tempty_bar + half), 0));
        if (cb < prm.n_cell_blocks) dn_epi_merge<NT>(prm, cb, k0, nk, s_wred, et);
        else dn_epi_bar<NT>();
        par ^= 1u << half;
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- host side ------------------------------------------------------------------------------------------------------
// The term decomposition the e2m1 operands use, for inspection and tests (no device needed): writes the
// terms of v (each in {1,2,3,4,6}, summing to v) and their 4-bit e2m1 codes; returns the number of terms
// (1 for v = 0, whose single term is 0), or -1 when max_terms is too small.
extern "C" int cv_e2m1_terms(uint32_t v, uint32_t* terms, uint32_t* codes, int max_terms) {
  const int n = (int)f4_nterms(v);
  if (n > max_terms) return -1;
  for (int t = 0; t < n; ++t) {
    const uint32_t x = f4_term(v, (uint32_t)t);
    if (terms) terms[t] = x;
    if (codes) codes[t] = t == 0 ? f4_code0(v) : f4_code(x);
  }
  return n;
}

static inline int64_t f4_round_up(int64_t v, int64_t m) { return ((v + m - 1) / m) * m; }

// Shape of the e2m1 path for the current problem (fills the fp4 members of the plan); 0 = not applicable.
int cv_f4_plan(cv_handle* h, DensePlan* pl) {
  const int I = h->B + 1;
  pl->fp4 = 0;
  if (I > F4_NB || !pl->pair) return 0;
  // at most DN_MAX_G annotations per sub-tile: the epilogue's per-annotation scratch (shared-memory
  // variability partials, the fixed-point form's thread-local state) is sized for that; with few
  // iterations (B < 15) the rest of the sub-tile's rows stay zero
  pl->Ga = std::min(F4_NA / I, DN_MAX_G);
  pl->Gb = std::min(F4_NB / I, DN_MAX_G);
  pl->n_sg = (int)((h->K + pl->Ga + pl->Gb - 1) / (pl->Ga + pl->Gb));
  pl->Ppad4 = f4_round_up(h->P, F4_KPEAKS);
  // reserve of overflow columns: as many as there are peaks (the operand rows are then as long as
  // the u8 rows); unused k-blocks of the reserve are never read, a chunk that needs more goes to u8
  pl->Hcap = std::max<int64_t>(1024, pl->Ppad4);
  pl->rowb = (pl->Ppad4 + pl->Hcap) / 2;
  pl->rowsM4 = (int64_t)pl->n_sg * F4_SG_ROWS;
  pl->fp4 = 1;
  return 1;
}

// Operand M4 (peak columns) + T_q.  Runs after cv_dense_build_rows' gather inputs (bit columns,
// inverse background map) exist on the compute stream.
int cv_f4_build_rows(cv_handle* h, const DensePlan& pl) {
  const int64_t P = h->P, K = h->K;
  const int B = h->B, I = B + 1;
  const int64_t Kw = (K + 31) / 32;
  CV_TRY(cv_ensure(h, h->f4_Tq, (size_t)(P + 1) * 4));
  CV_TRY(cv_ensure(h, h->f4_Sq, (size_t)(P + 1) * 4));
  CV_TRY(cv_ensure(h, h->f4_ncol, (size_t)(P + 1) * 4));
  CV_TRY(cv_ensure(h, h->f4_off, (size_t)(P + 2) * 4));
  CV_TRY(cv_ensure(h, h->f4_colq, (size_t)pl.Hcap * 4));
  CV_TRY(cv_ensure(h, h->f4_colt, (size_t)pl.Hcap * 4));
  CV_TRY(cv_ensure(h, h->f4_info, 64));
  CV_CUDA(h, cudaMemsetAsync(h->f4_Tq.p, 0, (size_t)(P + 1) * 4, h->stream));
  {
    dim3 grid((unsigned)pl.n_sg, 2, 8);
    CV_LAUNCH(h, k_f4_zero_pad, grid, 256, 0, h->dn_M.as<uint8_t>(), (int)K, pl.Ga, pl.Gb, I, pl.rowb);
  }
  dim3 grid((unsigned)((pl.Ppad4 + F4_TQ - 1) / F4_TQ), (unsigned)I, (unsigned)Kw);
  CV_LAUNCH(h, k_f4_rows_main, grid, 256, 0, h->dn_abits.as<uint32_t>(), h->dn_invptr.as<uint32_t>(),
            h->dn_invsrc.as<int32_t>(), (int)K, (int)P, B, pl.Ga, pl.Gb, pl.rowb, pl.Ppad4, h->dn_M.as<uint8_t>(),
            h->f4_Tq.as<uint32_t>(), h->flags.as<uint32_t>() + CV_DFLAGS);
  CV_LAUNCH(h, k_f4_tq_total, (unsigned)((P + 255) / 256), 256, 0, h->f4_Tq.as<uint32_t>(), (int)P,
            h->flags.as<uint32_t>() + CV_DFLAGS + 7);
  return CV_OK;
}

// Cells [c0, c0 + rows) (c1 = first cell past the chunk's real cells): overflow columns of the chunk, both
// operands' overflow parts, Cd4, GEMM + fused epilogue.  prm: the launch parameters cv_dense_chunk filled.
int cv_f4_chunk(cv_handle* h, const DensePlan& pl, const DenseParams& prm_in, int64_t c0, int64_t c1, int64_t rows,
                int fixed_point_epilogue, int check_now) {
  const int64_t P = h->P, K = h->K;
  const int B = h->B, I = B + 1;
  const int64_t Kw = (K + 31) / 32;
  const int64_t Pb = pl.Ppad4 / 2;
  uint32_t* dflags = h->flags.as<uint32_t>() + CV_DFLAGS;
  uint32_t* info = h->f4_info.as<uint32_t>();
  cv_span_begin(h, CV_SPAN_LAYOUT);
  // which peaks of this chunk hold a count outside S, and how many terms the largest one needs
  CV_CUDA(h, cudaMemsetAsync(h->f4_Sq.p, 0, (size_t)(P + 1) * 4, h->stream));
  CV_LAUNCH(h, k_f4_terms, (unsigned)(h->num_sms * 8), 256, 0, h->colptr.as<int64_t>(), h->rowidx.as<int32_t>(),
            h->val.as<uint32_t>(), c0, c1, h->f4_Sq.as<uint32_t>());
  CV_LAUNCH(h, k_f4_ncols, (unsigned)((P + 255) / 256), 256, 0, h->f4_Sq.as<uint32_t>(), h->f4_Tq.as<uint32_t>(), (int)P,
            h->f4_ncol.as<uint32_t>());
  CV_TRY(cv_exclusive_scan_u32_async(h, h->f4_ncol.as<uint32_t>(), h->f4_off.as<uint32_t>(), P));
  CV_LAUNCH(h, k_f4_info, 1, 1, 0, h->f4_off.as<uint32_t>(), (int)P, pl.Ppad4, (uint32_t)pl.Hcap, info, dflags);
  CV_LAUNCH(h, k_f4_collist, (unsigned)((P + 255) / 256), 256, 0, h->f4_off.as<uint32_t>(), h->f4_Tq.as<uint32_t>(), info,
            (int)P, h->f4_colq.as<int32_t>(), h->f4_colt.as<uint32_t>());
  {
    dim3 grid((unsigned)std::min<int64_t>((pl.Hcap + F4_OV_COLS - 1) / F4_OV_COLS, 8), (unsigned)I, (unsigned)Kw);
    CV_LAUNCH(h, k_f4_rows_ov, grid, 256, 0, h->dn_abits.as<uint32_t>(), h->dn_invptr.as<uint32_t>(),
              h->dn_invsrc.as<int32_t>(), h->f4_colq.as<int32_t>(), h->f4_colt.as<uint32_t>(), info, (int)K, (int)P, B,
              pl.Ga, pl.Gb, pl.rowb, Pb, h->dn_M.as<uint8_t>(), dflags);
  }
  {
    // a cell is latency-bound (zero the row, a few dependent loads per thread, write the row), so up
    // to three CTAs share an SM when a row's peak columns plus their overflow columns fit a third of
    // its shared memory (config 2: 50 KB of peak columns -> 3 CTAs of 74.75 KB; ncu capture X: one CTA
    // per SM kept 25 % of the warp slots busy); longer rows are built in pieces by one CTA per SM
    const int budget = h->max_smem_optin - 2048;
    const int64_t sm_total = (int64_t)h->max_smem_optin + 1024;            // per SM: the opt-in maximum plus one CTA's reserve
    int nc = 1;
    // (room for overflow columns worth a quarter of the peak columns: dense k-mer annotations need
    // that much, and a row that does not fit its piece costs a second pass over the cell's column)
    for (int c = 3; c >= 2; --c)
      if (sm_total / c - 1280 >= Pb + Pb / 4 + 8192) { nc = c; break; }
    const int piece = nc == 1 ? (int)std::min<int64_t>(pl.rowb, (budget / 128) * 128)
                              : (int)std::min<int64_t>(pl.rowb, ((sm_total / nc - 1280) / 128) * 128);
    CV_CUDA(h, cudaMemsetAsync(h->work_ctr.as<uint32_t>() + 1, 0, 4, h->stream));
    CV_CUDA(h, cudaFuncSetAttribute(k_f4_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, piece));
    CV_LAUNCH(h, k_f4_cells, (unsigned)std::min<int64_t>(rows, (int64_t)h->num_sms * nc), 512, (size_t)piece,
              h->colptr.as<int64_t>(), h->rowidx.as<int32_t>(), h->val.as<uint32_t>(), (int)h->N, (int)c0, (int)rows,
              h->f4_Tq.as<uint32_t>(), h->f4_Sq.as<uint32_t>(), h->f4_off.as<uint32_t>(), info, Pb, pl.rowb, piece,
              h->dn_C.as<uint8_t>(), h->work_ctr.as<uint32_t>() + 1, dflags);
  }
  cv_span_end(h);
  if (check_now) {
    // first chunk of a call: does it fit the reserve?  (one short sync before the first GEMM; later
    // chunks are only checked at the end of the call)
    uint32_t flag = 0;
    CV_CUDA(h, cudaMemcpyAsync(&flag, dflags + 5, 4, cudaMemcpyDeviceToHost, h->stream));
    CV_CUDA(h, cudaStreamSynchronize(h->stream));
    if (flag) {
      h->f4_veto = true;
      CV_FAIL(h, CV_ERR_UNSUPPORTED, "counts not representable on the e2m1 path (flags %u)", flag);
    }
  }
  cv_span_begin(h, CV_SPAN_CONTRACT);
  CUtensorMap map_c, map_ma, map_mb;
  CV_TRY(cv_make_map_u8(h, &map_c, h->dn_C.p, (uint64_t)rows, (uint64_t)pl.rowb, (uint64_t)pl.rowb, DN_BLOCK_M));
  CV_TRY(cv_make_map_u8(h, &map_ma, h->dn_M.p, (uint64_t)pl.rowsM4, (uint64_t)pl.rowb, (uint64_t)pl.rowb, F4_NA / 2));
  CV_TRY(cv_make_map_u8(h, &map_mb, h->dn_M.p, (uint64_t)pl.rowsM4, (uint64_t)pl.rowb, (uint64_t)pl.rowb, F4_NB / 2));
  DenseParams prm = prm_in;
  prm.Ga = pl.Ga; prm.Gb = pl.Gb; prm.n_sg = pl.n_sg;
  prm.sync_ctr = nullptr;
  prm.sync_lag = h->opt_lockstep == 2 ? 1 : 0;
  if (h->opt_lockstep) {
    CV_TRY(cv_ensure(h, h->f4_sync, 65536 * 4));
    prm.sync_ctr = h->f4_sync.as<uint32_t>();
  }
  prm.n_kblocks_dev = info;
  prm.band = h->opt_band > 0 ? h->opt_band : 6;
  const int n_units = (prm.n_cell_blocks + 1) / 2;
  prm.n_tiles = (int64_t)n_units * pl.n_sg;
  int pairs = h->num_sms / 2;
  if (prm.n_tiles < pairs) pairs = (int)prm.n_tiles;
  if (prm.sync_ctr) {
    const int64_t waves = (prm.n_tiles + pairs - 1) / pairs;
    if (2 * waves >= F4_SYNC_SLOTS - 1) prm.sync_ctr = nullptr;
    else {
      CV_CUDA(h, cudaMemsetAsync(prm.sync_ctr, 0, (size_t)(2 * waves) * 4, h->stream));
      CV_CUDA(h, cudaMemsetAsync(prm.sync_ctr + F4_SYNC_SLOTS - 1, 0, 4, h->stream));     // the "lockstep given up" word
    }
  }
  if (fixed_point_epilogue) {
    CV_CUDA(h, cudaFuncSetAttribute(k_dense_contract_fp4<DN_EPI_FUSED32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    F4_SMEM_BYTES));
    k_dense_contract_fp4<DN_EPI_FUSED32><<<2 * pairs, F4_THREADS, F4_SMEM_BYTES, h->stream>>>(map_c, map_ma, map_mb, prm);
  } else {
    CV_CUDA(h, cudaFuncSetAttribute(k_dense_contract_fp4<DN_EPI_FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    F4_SMEM_BYTES));
    k_dense_contract_fp4<DN_EPI_FUSED><<<2 * pairs, F4_THREADS, F4_SMEM_BYTES, h->stream>>>(map_c, map_ma, map_mb, prm);
  }
  h->launches++;
  CV_CUDA(h, cudaGetLastError());
  if (h->opt_sync_launch) CV_TRY(cv_sync_after_launch(h, "k_dense_contract_fp4"));
  cv_span_end(h);
  return CV_OK;
}
