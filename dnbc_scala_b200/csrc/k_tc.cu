// k_tc.cu -- K6: forward / backward recursions for large state spaces (65 .. 512 states, padded to a
// multiple of 64: NP = 128, 192, ... 512) on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// No reference counterpart (the reference has no forward-backward); parity is
// against the fp64 oracle (oracle/dnbc_oracle.c fb_one), same outputs as k_fb2.cu.
//
// For a batch of 128 sequences the step  alpha_t = b_t * (alpha_{t-1} A)  is a dense
// [128 x N] x [N x N] contraction.  One CTA owns the batch for its whole length:
//   * the accumulator D[128][N] fp32 lives in TMEM (N = 512 uses all 512 columns);
//   * operands are bf16 SPLIT pairs (x = hi + lo, hi = bf16(x), lo = bf16(x - hi)); the
//     product is formed as hi*hi + hi*lo + lo*hi into the same fp32 accumulator (relative
//     error ~2^-16 per term; a single bf16 pass would be ~2^-9, far outside the 1e-5 budget);
//   * the transition matrix (K-major, [to][from] for the forward pass, [from][to] for the
//     backward pass) is kept in global memory as ready-made images of the 128-byte-swizzled
//     K-major shared-memory tiles (k_tc_tables), and so is the staging copy of the step's
//     operand X; a producer thread streams them from L2 in 64-wide K blocks through a 2-stage
//     ring with bulk async copies (cp.async.bulk, completion counted on an mbarrier), a
//     second thread issues the tcgen05.mma's and commits to an mbarrier per stage and per
//     256-column half of the accumulator;
//   * the epilogue (8 warps, thread = (sequence row, half of the columns)) reads TMEM with
//     tcgen05.ld, applies the emission, normalises the row, writes alpha / gamma / u to HBM and
//     the next step's split operand to the L2-resident staging tiles.  The warps of the first
//     column half start while the MMAs of the second half are still running.
// The step-to-step dependency (the epilogue of step t feeds the MMA of step t+1) serialises
// MMA and epilogue within a CTA; 148 CTAs keep all SMs busy on independent batches.
#include <cuda_bf16.h>

#include "dnbc_internal.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TC_M = 128;          // sequences per CTA batch (UMMA M)
constexpr int TC_NH = 256;         // UMMA N per instruction
constexpr int TC_KB = 64;          // K block (one 128-byte swizzled row of bf16)
constexpr int TC_EPI = 256;        // epilogue threads (warps 0-7)
constexpr int TC_THREADS = 320;    // + warp 8 (lane 0: bulk-copy producer) + warp 9 (lane 0: MMA issuer)
constexpr int A_TILE = TC_M * TC_KB * 2;    // 16 KB
constexpr int B_TILE = TC_NH * TC_KB * 2;   // 32 KB
constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;   // hi + lo of both operands: 96 KB
constexpr uint32_t BULK = 16 * 1024;        // bytes per bulk copy
#ifndef TC_FWD_PFE
#define TC_FWD_PFE 1
#endif
#ifndef TC_BWD_PFE
#define TC_BWD_PFE 1
#endif

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// global -> shared bulk async copy (TMA unit, no tensor map: source and destination are contiguous)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI) : "memory"); }

// the initial distribution for columns j0 .. j0+31 (the same 32 values in every lane: broadcast loads)
__device__ __forceinline__ void load_pi32(const float *g, int j0, int N, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = j0 + i < N ? __ldg(g + j0 + i) : 0.f;
}

// ---- coalesced global traffic for the epilogue ----------------------------------------------------
// An epilogue thread owns one sequence row (its TMEM lane), but a warp-wide access in which every lane
// touches a different row costs one LSU cycle per row.  32 x 32 fp32 chunks therefore go through a
// warp-private 4 KB shared-memory tile (16-byte slots XOR-swizzled by the row, conflict-free both ways);
// the tile is read / written by row on the thread's side, and global accesses are made
//   V4 (N a multiple of 4: rows are whole, 16-byte aligned float4's; a slot is all states or all padding):
//       lane = (row it*4 + lane/8, 16-byte slot lane%8), 4 rows x 128 contiguous bytes per instruction;
//   otherwise (any N: the row pitch is only 4-byte aligned):
//       lane = column, one row of 32 consecutive floats per instruction, columns past N masked off.
// The first version of the any-N epilogue let every thread walk its own row (32 rows per warp access): N = 200
// ran at 74 ms per 5e6 points where N = 256 took 12.
constexpr int TB_BYTES = 32 * 128;
constexpr unsigned FULL = 0xffffffffu;
__device__ __forceinline__ uint32_t tb_off(int r, int c4) { return (uint32_t)(r * 128 + ((c4 ^ (r & 7)) << 4)); }
__device__ __forceinline__ uint32_t tb_off1(int r, int c) { return tb_off(r, c >> 2) + (uint32_t)(c & 3) * 4u; }

// issue the loads of columns col .. col+31 of the warp's 32 rows into q (p = this lane's row's point index;
// pp = point index of row it*4 + lane/8, used by the V4 form only).
// LAST: this is the last use of the lines (evict-first), so that the per-step streams (emissions, alpha,
// gamma, u: ~150 MB per step over the chip) do not push the staging tiles and the transition matrix out of L2.
template <bool LAST, bool V4>
__device__ __forceinline__ void chunk_ldg(const float *base, const int (&pp)[8], int p, unsigned mask, int N, int col, int lane,
                                          float fill, float (&q)[32]) {
  if (V4) {
    if (col + (lane & 7) * 4 >= N) mask = 0u;                // this lane's float4 slot is padding
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rr = it * 4 + (lane >> 3);
      const float4 *src = (const float4 *)(base + (int64_t)pp[it] * N + col) + (lane & 7);
      const float4 v = ((mask >> rr) & 1u) ? (LAST ? __ldcs(src) : __ldg(src)) : make_float4(fill, fill, fill, fill);
      q[4 * it] = v.x; q[4 * it + 1] = v.y; q[4 * it + 2] = v.z; q[4 * it + 3] = v.w;
    }
  } else {
    const bool con = col + lane < N;
#pragma unroll
    for (int it = 0; it < 32; ++it) {
      const int pr = __shfl_sync(FULL, p, it);
      const float *src = base + (int64_t)pr * N + col + lane;
      q[it] = (con && ((mask >> it) & 1u)) ? (LAST ? __ldcs(src) : __ldg(src)) : fill;
    }
  }
}
// ... and hand every thread the 32 values of its own row
template <bool V4>
__device__ __forceinline__ void chunk_to_rows(unsigned char *tb, int lane, const float (&q)[32], float (&v)[32]) {
  if (V4) {
#pragma unroll
    for (int it = 0; it < 8; ++it)
      *(float4 *)(tb + tb_off(it * 4 + (lane >> 3), lane & 7)) = make_float4(q[4 * it], q[4 * it + 1], q[4 * it + 2], q[4 * it + 3]);
  } else {
#pragma unroll
    for (int it = 0; it < 32; ++it) *(float *)(tb + tb_off1(it, lane)) = q[it];
  }
  __syncwarp();
#pragma unroll
  for (int c4 = 0; c4 < 8; ++c4) {
    const float4 x = *(const float4 *)(tb + tb_off(lane, c4));
    v[4 * c4] = x.x; v[4 * c4 + 1] = x.y; v[4 * c4 + 2] = x.z; v[4 * c4 + 3] = x.w;
  }
  __syncwarp();
}
__device__ __forceinline__ void rows_to_tb(unsigned char *tb, int lane, const float (&v)[32]) {
#pragma unroll
  for (int c4 = 0; c4 < 8; ++c4)
    *(float4 *)(tb + tb_off(lane, c4)) = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
  __syncwarp();
}
// tile -> rows (point + drow) of a row-major fp32 array, columns col .. col+31, rows selected by mask
template <bool V4>
__device__ __forceinline__ void tb_store_f32(const unsigned char *tb, float *base, const int (&pp)[8], int p, unsigned mask, int N,
                                             int col, int lane, int drow) {
  if (V4) {
    if (col + (lane & 7) * 4 >= N) return;                   // padded columns are never written
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rr = it * 4 + (lane >> 3);
      if ((mask >> rr) & 1u)                                 // streaming store: the consumers are later kernels
        __stcs((float4 *)(base + (int64_t)(pp[it] + drow) * N + col) + (lane & 7), *(const float4 *)(tb + tb_off(rr, lane & 7)));
    }
  } else {
    const bool con = col + lane < N;
#pragma unroll
    for (int it = 0; it < 32; ++it) {
      const int pr = __shfl_sync(FULL, p, it);
      if (con && ((mask >> it) & 1u)) __stcs(base + (int64_t)(pr + drow) * N + col + lane, *(const float *)(tb + tb_off1(it, lane)));
    }
  }
}
// zeros into columns col .. col+31 of the rows selected by mask
template <bool V4>
__device__ __forceinline__ void zero_rows(float *base, const int (&pp)[8], int p, unsigned mask, int N, int col, int lane) {
  if (V4) {
    if (col + (lane & 7) * 4 >= N) return;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rr = it * 4 + (lane >> 3);
      if ((mask >> rr) & 1u) *((float4 *)(base + (int64_t)pp[it] * N + col) + (lane & 7)) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
    const bool con = col + lane < N;
#pragma unroll
    for (int it = 0; it < 32; ++it) {
      const int pr = __shfl_sync(FULL, p, it);
      if (con && ((mask >> it) & 1u)) base[(int64_t)pr * N + col + lane] = 0.f;
    }
  }
}
// pull columns col .. col+31 of the rows selected by mask into L2 (no registers): one request per 128-byte line
template <bool V4>
__device__ __forceinline__ void chunk_prefetch_l2(const float *base, const int (&pp)[8], int p, unsigned mask, int N, int col, int lane) {
  if (V4) {
    if (col + (lane & 7) * 4 >= N || (lane & 7) != 0) mask = 0u;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rr = it * 4 + (lane >> 3);
      if ((mask >> rr) & 1u) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (int64_t)pp[it] * N + col));
    }
  } else {
    if (col < N && ((mask >> lane) & 1u)) {                  // lane = row: the chunk's first and last byte (at most two lines)
      const float *r0 = base + (int64_t)p * N + col;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(r0));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(r0 + (col + 31 < N ? 31 : N - 1 - col)));
    }
  }
}
// tile -> bf16 hi / lo split into the staging tile images (rows row0 .. row0+31, columns c0 .. c0+31 of X):
// lane = (row it*8 + lane/4, 8 columns lane%4), one 16-byte store each for hi and lo
__device__ __forceinline__ void tb_store_split(const unsigned char *tb, unsigned char *xt, int row0, int c0, int lane) {
  unsigned char *tile = xt + (size_t)(c0 >> 6) * (2 * A_TILE);
  const int cb = (c0 & 63) >> 3, i = lane & 3;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int rr = it * 8 + (lane >> 2);
    const float4 u = *(const float4 *)(tb + tb_off(rr, 2 * i)), v = *(const float4 *)(tb + tb_off(rr, 2 * i + 1));
    const float x[8] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * e]), h1 = __float2bfloat16_rn(x[2 * e + 1]);
      const __nv_bfloat16 l0 = __float2bfloat16_rn(x[2 * e] - __bfloat162float(h0));
      const __nv_bfloat16 l1 = __float2bfloat16_rn(x[2 * e + 1] - __bfloat162float(h1));
      h[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      l[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    const uint32_t o = sw128(row0 + rr, cb + i);
    *(uint4 *)(tile + o) = make_uint4(h[0], h[1], h[2], h[3]);
    *(uint4 *)(tile + A_TILE + o) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

struct TcArgs {
  int N, NP;                    // states, states padded to a multiple of 64 (128 .. 512)
  const int64_t *off;
  int64_t s0, s1, p0;
  const float *E;               // [np][N]
  const float *mrow;            // [np] max_j E
  float *AL, *GA, *U, *cs;
  double *stats;
  int64_t st_pi;
  const float *pi;              // [N]
  const unsigned char *Bt;      // operand B (forward: A^T, backward: A) as tile images: [nh][kb][hi | lo][256 x 64]
  unsigned char *Xt;            // [grid] staging of the next step's operand X as tile images: [kb][hi | lo][128 x 64]
};

// Producer (one thread): stream the K-block tiles of X and B for one step into the ring.  The accumulator is
// formed in column blocks of up to 256 (UMMA N); the last block of a padded state count that is not a multiple of
// 256 has 64, 128 or 192 columns, and only those rows of the B tile images are fetched (the first `rows` rows of
// a 128-byte-swizzled K-major tile are its first rows * 128 bytes).
__device__ void tc_produce(const TcArgs &a, uint32_t smem_base, uint32_t bar_full, uint32_t bar_empty, uint32_t (&uses)[2],
                           const unsigned char *xt) {
  const int nkb = a.NP / TC_KB, nnh = (a.NP + TC_NH - 1) / TC_NH;
  int it = 0;
  for (int nh = 0; nh < nnh; ++nh) {
    const int rows = a.NP - nh * TC_NH < TC_NH ? a.NP - nh * TC_NH : TC_NH;
    const uint32_t bbytes = (uint32_t)rows * 128u;           // hi (and lo) bytes of this block's B tile
    for (int kb = 0; kb < nkb; ++kb, ++it) {
      const int s = it & 1;
      const uint32_t st = smem_base + s * STAGE_BYTES, full = bar_full + 8 * s;
      // the stage is free once the MMAs that read its previous contents have completed
      if (uses[s] > 0) mbar_wait(bar_empty + 8 * s, (uses[s] - 1) & 1);
      mbar_expect_tx(full, 2 * A_TILE + 2 * bbytes);
      const unsigned char *xs = xt + (size_t)kb * (2 * A_TILE);
      const unsigned char *bs = a.Bt + ((size_t)nh * nkb + kb) * (2 * B_TILE);
#pragma unroll
      for (uint32_t o = 0; o < 2 * A_TILE; o += BULK) bulk_g2s(st + o, xs + o, BULK, full);
      for (uint32_t o = 0; o < bbytes; o += BULK) {
        const uint32_t n = bbytes - o < BULK ? bbytes - o : BULK;
        bulk_g2s(st + 2 * A_TILE + o, bs + o, n, full);
        bulk_g2s(st + 2 * A_TILE + B_TILE + o, bs + B_TILE + o, n, full);
      }
      uses[s] += 1;
    }
  }
}

// MMA issuer (one thread): D[128][NP] (TMEM) = X[128][NP] * B^T, up to 256 accumulator columns at a time
__device__ void tc_mma(const TcArgs &a, uint32_t smem_base, uint32_t bar_full, uint32_t bar_empty, uint32_t bar_done,
                       uint32_t tmem, uint32_t (&uses)[2]) {
  const int nkb = a.NP / TC_KB, nnh = (a.NP + TC_NH - 1) / TC_NH;
  int it = 0;
  for (int nh = 0; nh < nnh; ++nh) {
    const int rows = a.NP - nh * TC_NH < TC_NH ? a.NP - nh * TC_NH : TC_NH;
    const uint32_t idesc = umma_idesc(TC_M, rows);
    for (int kb = 0; kb < nkb; ++kb, ++it) {
      const int s = it & 1;
      const uint32_t st = smem_base + s * STAGE_BYTES;
      mbar_wait(bar_full + 8 * s, uses[s] & 1);
      tc_fence_after();
      const uint64_t dAh = umma_desc(st), dAl = umma_desc(st + A_TILE);
      const uint64_t dBh = umma_desc(st + 2 * A_TILE), dBl = umma_desc(st + 2 * A_TILE + B_TILE);
      const uint32_t d = tmem + nh * TC_NH;
#pragma unroll
      for (int k16 = 0; k16 < TC_KB / 16; ++k16) {
        const uint64_t ko = (uint64_t)(k16 * 32 >> 4);     // 16 bf16 = 32 B further along K
        umma_bf16(d, dAh + ko, dBh + ko, idesc, (kb | k16) ? 1u : 0u);
        umma_bf16(d, dAh + ko, dBl + ko, idesc, 1u);
        umma_bf16(d, dAl + ko, dBh + ko, idesc, 1u);
      }
      umma_commit(bar_empty + 8 * s);
      if (kb == nkb - 1) umma_commit(bar_done + 8 * nh);   // this block of the accumulator is complete
      uses[s] += 1;
    }
  }
}

// phase timers of one CTA (development aid: -DDNBC_TC_TIMING prints clock64 totals of block 0)
#ifdef DNBC_TC_TIMING
#define TCT_DECL long long tct[6] = {0, 0, 0, 0, 0, 0}, tct0 = 0
#define TCT_START tct0 = clock64()
#define TCT_LAP(i) do { const long long n_ = clock64(); tct[i] += n_ - tct0; tct0 = n_; } while (0)
#else
#define TCT_DECL
#define TCT_START
#define TCT_LAP(i)
#endif

template <int DIR, bool V4>   // DIR: 0 = forward, 1 = backward; V4: N is a multiple of 4 (float4 form of the epilogue traffic)
__global__ void __launch_bounds__(TC_THREADS, 1) k_fb_tc(TcArgs a) {
  TCT_DECL;
  extern __shared__ unsigned char smraw[];
  __shared__ __align__(8) unsigned long long bars[6];        // full[2], empty[2], done[2]
  __shared__ uint32_t tmem_slot;
  __shared__ float rsum[TC_M][2];
  __shared__ int sT[TC_M];
  __shared__ int sTmax;
  const int N = a.N, NP = a.NP;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const bool epi = w < TC_EPI / 32;
  const int row = (w & 3) * 32 + lane, half = (w >> 2) & 1;  // TMEM lane quadrant = warp % 4
  const int ncol = NP / 2, col0 = half * ncol;               // this thread's column range (whole 32-column chunks)
  const uint32_t smem_base = (smem_u32(smraw) + 1023u) & ~1023u;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[2]), bar_done = smem_u32(&bars[4]);
  const uint32_t my_done = bar_done + 8 * ((col0 + ncol - 1) / TC_NH);   // the last accumulator block this thread's columns are in
  unsigned char *tb = smraw + (smem_base - smem_u32(smraw)) + 2 * STAGE_BYTES + (w & 7) * TB_BYTES;   // this warp's transpose tile
  const int row0 = (w & 3) * 32;
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) mbar_init(bar_full + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t trow = tmem + ((uint32_t)((w & 3) * 32) << 16);
  uint32_t uses[2] = {0u, 0u};
  uint32_t gsteps = 0;                                       // products issued so far (parity of the done barriers)
  unsigned char *xt = a.Xt + (size_t)blockIdx.x * TC_M * NP * 4;   // hi + lo, 2 bytes each
  double ll = 0.0;

  for (int64_t sb = a.s0 + (int64_t)blockIdx.x * TC_M; sb < a.s1; sb += (int64_t)gridDim.x * TC_M) {
    const int64_t s = sb + row;
    const int T = (epi && s < a.s1) ? (int)(a.off[s + 1] - a.off[s]) : 0;
    const int64_t pt = (epi && s < a.s1) ? a.off[s] - a.p0 : 0;
    if (tid == 0) sTmax = 0;
    __syncthreads();
    if (epi && half == 0) { sT[row] = T; atomicMax(&sTmax, T); }
    __syncthreads();
    const int Tmax = sTmax;
    for (int k = 0; k < Tmax; ++k) {
      TCT_START;
      if (!epi) {
        // ---- warp 8: producer, warp 9: MMA issuer (one elected lane each)
        if (k > 0 && lane == 0) {
          if (w == TC_EPI / 32) tc_produce(a, smem_base, bar_full, bar_empty, uses, xt);
          else {
            tc_fence_after();
            tc_mma(a, smem_base, bar_full, bar_empty, bar_done, tmem, uses);
          }
        }
        __syncwarp();
      } else {
      // forward walks t = k; backward walks t = T - 1 - k (rows aligned at their last step)
      const int t = DIR == 0 ? k : T - 1 - k;
      const bool live = DIR == 0 ? (k < T) : (t >= 0);
      const int64_t p = pt + (live ? t : 0);
      const int pi32 = (int)p;
      const float mx = live ? a.mrow[p] : 0.f;
      const float mxs = mx == -INFINITY ? 0.f : mx;
      const unsigned lmask = __ballot_sync(FULL, live);
      int pp[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) pp[it] = V4 ? __shfl_sync(FULL, pi32, it * 4 + (lane >> 3)) : 0;
      if (DIR == 0) {
        // ---- v = d * exp(E - m); c = sum v; alpha = v / c ----   (emission chunks are fetched one chunk ahead)
        float sum = 0.f;
        float en[32];
        chunk_ldg<true, V4>(a.E, pp, pi32, lmask, N, col0, lane, -INFINITY, en);
#if TC_FWD_PFE
        for (int ch = 1; ch < ncol / 32; ++ch) chunk_prefetch_l2<V4>(a.E, pp, pi32, lmask, N, col0 + ch * 32, lane);
#endif
        if (k > 0) { mbar_wait(my_done, gsteps & 1); tc_fence_after(); }
        TCT_LAP(0);
        for (int ch = 0; ch < ncol / 32; ++ch) {
          const int c0 = col0 + ch * 32;
          float d[32], e[32];
          chunk_to_rows<V4>(tb, lane, en, e);
          if (ch + 1 < ncol / 32) chunk_ldg<true, V4>(a.E, pp, pi32, lmask, N, c0 + 32, lane, -INFINITY, en);
          if (k > 0) tmem_ld32(trow + c0, d);
          else load_pi32(a.pi, c0, N, d);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            d[i] *= ex2_approx((e[i] - mxs) * (float)DNBC_LOG2E);
            sum += d[i];
          }
          tmem_st32(trow + c0, d);
        }
        rsum[row][half] = sum;
        TCT_LAP(1);
        epi_sync();
        TCT_LAP(2);
        const float c = rsum[row][0] + rsum[row][1];
        const float inv = live ? 1.0f / c : 0.f;
        for (int ch = 0; ch < ncol / 32; ++ch) {
          const int c0 = col0 + ch * 32;
          float d[32];
          tmem_ld32(trow + c0, d);
#pragma unroll
          for (int i = 0; i < 32; ++i) d[i] *= inv;
          rows_to_tb(tb, lane, d);
          tb_store_f32<V4>(tb, a.AL, pp, pi32, lmask, N, c0, lane, 0);
          tb_store_split(tb, xt, row0, c0, lane);
          __syncwarp();
        }
        if (live && half == 0) {
          a.cs[p] = c;
          ll += (double)logf(c) + (double)mxs;
        }
      } else {
        // ---- beta_t in TMEM (1 on the last step); gamma = alpha * beta; next u = exp(E_t - m_t) beta / c_t ----
        const float ic = live ? 1.0f / a.cs[p] : 0.f;
        const unsigned tpos = __ballot_sync(FULL, live && t > 0);
        float qa[32], qe[32];
        chunk_ldg<false, V4>(a.AL, pp, pi32, lmask, N, col0, lane, 0.f, qa);
        // (no L2 prefetch of the step's rows: ~190 MB stream through the 126 MB L2 per step, so lines fetched a
        // product phase ahead were evicted again before their use -- measured 97 -> 94 ms without it)
        if (k > 0) { mbar_wait(my_done, gsteps & 1); tc_fence_after(); }
        TCT_LAP(0);
        // sum_j alpha^_t(j) beta^_t(j) is 1 in exact arithmetic; with split-bf16 products it drifts by
        // ~1e-5 per step, so beta^_t is rescaled to make it 1 at every step (as k_fb2 / k_fb3 do)
        float sdot = 0.f;
        for (int ch = 0; ch < ncol / 32; ++ch) {
          const int c0 = col0 + ch * 32;
          float d[32], al[32];
          chunk_to_rows<V4>(tb, lane, qa, al);
          if (ch + 1 < ncol / 32) chunk_ldg<false, V4>(a.AL, pp, pi32, lmask, N, c0 + 32, lane, 0.f, qa);
#if TC_BWD_PFE
          chunk_prefetch_l2<V4>(a.E, pp, pi32, lmask, N, c0, lane);   // the second pass finds its emission rows in L2
#endif
          if (k > 0) tmem_ld32(trow + c0, d);
          else {
#pragma unroll
            for (int i = 0; i < 32; ++i) d[i] = 1.f;
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) sdot = fmaf(al[i], d[i], sdot);
        }
        rsum[row][half] = sdot;
        TCT_LAP(1);
        epi_sync();
        TCT_LAP(2);
        sdot = rsum[row][0] + rsum[row][1];
        const float fix = (live && sdot > 0.f) ? 1.0f / sdot : 1.f;
        for (int ch = 0; ch < ncol / 32; ++ch) {
          const int c0 = col0 + ch * 32;
          float d[32], e[32], al[32];
          // (both rows were pulled into L2 at the top of the step; the TMEM load overlaps the fetch)
          chunk_ldg<true, V4>(a.AL, pp, pi32, lmask, N, c0, lane, 0.f, qa);
          chunk_ldg<true, V4>(a.E, pp, pi32, lmask, N, c0, lane, -INFINITY, qe);
          if (k > 0) tmem_ld32(trow + c0, d);
          else {
#pragma unroll
            for (int i = 0; i < 32; ++i) d[i] = 1.f;
          }
          chunk_to_rows<V4>(tb, lane, qa, al);
          chunk_to_rows<V4>(tb, lane, qe, e);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            d[i] *= fix;
            al[i] *= d[i];                                                   // gamma
            e[i] = (live && t > 0) ? ex2_approx((e[i] - mxs) * (float)DNBC_LOG2E) * d[i] * ic : 0.f;   // next u
          }
          rows_to_tb(tb, lane, al);
          tb_store_f32<V4>(tb, a.GA, pp, pi32, lmask, N, c0, lane, 0);
          __syncwarp();
          const unsigned t0mask = __ballot_sync(FULL, live && t == 0);
          if (t0mask) {
            // pi-stat += gamma_0: the rows at their first step go through the tile once more and each lane adds up
            // one state column, so a warp issues 32 atomics instead of 32 x 32 (148 batches hitting 512 addresses)
            const bool mine = live && t == 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) al[i] = mine ? al[i] : 0.f;
            rows_to_tb(tb, lane, al);
            float colsum = 0.f;
#pragma unroll 8
            for (int r = 0; r < 32; ++r) colsum += *(const float *)(tb + tb_off1(r, lane));
            if (colsum != 0.f) atomicAdd(a.stats + a.st_pi + c0 + lane, (double)colsum);   // (padded columns sum to zero)
            __syncwarp();
          }
          rows_to_tb(tb, lane, e);
          tb_store_f32<V4>(tb, a.U, pp, pi32, tpos, N, c0, lane, -1);
          if (k == 0) zero_rows<V4>(a.U, pp, pi32, lmask, N, c0, lane);         // no transition out of the last step
          tb_store_split(tb, xt, row0, c0, lane);
          __syncwarp();
        }
      }
      // the staging tiles are read by the bulk copies (async proxy) of the next step
      asm volatile("fence.proxy.async.global;" ::: "memory");
      TCT_LAP(3);
      }
      if (k > 0) gsteps += 1;
      // TMEM reads done and the staging tiles visible before the next step's MMAs / loads
      tc_fence_before();
      __threadfence();
      TCT_LAP(4);
      __syncthreads();
      TCT_LAP(5);
    }
  }
  if (epi && DIR == 0 && half == 0 && ll != 0.0) atomicAdd(a.stats, ll);
#ifdef DNBC_TC_TIMING
  if (blockIdx.x == 0 && (tid == 0 || tid == 128 || tid == 256 || tid == 288))
    printf("k_fb_tc<%d> tid %3d: wait-done %lld  pass1 %lld  sync %lld  pass2 %lld  fence %lld  step-barrier %lld (clk)\n", DIR, tid,
           tct[0], tct[1], tct[2], tct[3], tct[4], tct[5]);
#endif
  tc_fence_before();
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// m[p] = max_j E[p][j]  (one warp per point)
__global__ void k_rowmax(const float *__restrict__ E, int N, int64_t np, float *__restrict__ m) {
  const int lane = threadIdx.x & 31;
  for (int64_t p = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; p < np; p += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    float v = -INFINITY;
    for (int j = lane; j < N; j += 32) v = fmaxf(v, E[p * N + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) m[p] = v;
  }
}

// bf16 split tables of the transition matrix as shared-memory tile images (what a bulk copy drops into a
// ring stage): operand row n, reduction index k -> tile (n / 256, k / 64), [hi | lo], 128-byte swizzled row.
// Tb = [from][to] (backward: n = from, k = to), Tf = [to][from] (forward: n = to, k = from).
__global__ void k_tc_tables(const double *__restrict__ A, int N, int NP, unsigned char *Tb, unsigned char *Tf) {
  const int nkb = NP / TC_KB;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < (int64_t)NP * NP; q += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(q / NP), c = (int)(q % NP);
    const float v = (r < N && c < N) ? (float)A[(int64_t)r * N + c] : 0.f;
    const float vt = (r < N && c < N) ? (float)A[(int64_t)c * N + r] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v), ht = __float2bfloat16_rn(vt);
    const size_t o = ((size_t)(r / TC_NH) * nkb + c / TC_KB) * (2 * B_TILE) + sw128(r % TC_NH, (c % TC_KB) >> 3) + (c & 7) * 2;
    *(__nv_bfloat16 *)(Tb + o) = h;
    *(__nv_bfloat16 *)(Tb + o + B_TILE) = __float2bfloat16_rn(v - __bfloat162float(h));
    *(__nv_bfloat16 *)(Tf + o) = ht;
    *(__nv_bfloat16 *)(Tf + o + B_TILE) = __float2bfloat16_rn(vt - __bfloat162float(ht));
  }
}

}  // namespace

// padded state count of the tensor-core recursion: a multiple of 64 (K blocks of 64 source states, two column halves
// of whole 32-column chunks for the epilogue threads); 0 = not supported
static int tc_np(const dnbc_ctx *c) {
  const int N = c->dm.N;
  return (N > 64 && N <= 512) ? (N + 63) / 64 * 64 : 0;
}
// bytes of one transition-matrix table: [column block of 256][K block][hi | lo][256 x 64]
static size_t tc_table_bytes(int NP) { return (size_t)((NP + TC_NH - 1) / TC_NH) * (NP / TC_KB) * (2 * B_TILE); }

bool fb_tc_supported(const dnbc_ctx *c) { return tc_np(c) != 0; }
// sequences one wave of the recursion covers (one 128-sequence batch per SM): chunks of the E-step are cut at whole
// multiples of it -- a batch costs the same whether all of its rows are live or not, and a chunk of 1.65 waves (what
// the scratch budget gives at N = 512, T = 200) takes the time of two
int64_t fb_tc_wave_sequences(const dnbc_ctx *c) { return tc_np(c) != 0 ? (int64_t)c->sm_count * TC_M : 0; }

void launch_tc_tables(dnbc_ctx *c) {
  const int NP = tc_np(c);
  const size_t tb = tc_table_bytes(NP);
  if (!c->tc_tab) {
    if (cudaMalloc((void **)&c->tc_tab, 2 * tb) != cudaSuccess) return;
  }
  unsigned char *t = (unsigned char *)c->tc_tab;
  c->launches++;
  k_tc_tables<<<c->sm_count * 2, 256, 0, c->stream>>>(c->d_A, c->dm.N, NP, t, t + tb);
}

int launch_fb_tc(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0, int64_t np, bool backward) {
  const int NP = tc_np(c), N = c->dm.N;
  const size_t tbytes = tc_table_bytes(NP);
  const int64_t batches = (s1 - s0 + TC_M - 1) / TC_M;
  int grid = (int)(batches < c->sm_count ? batches : c->sm_count);
  if (grid < 1) grid = 1;
  const size_t xbytes = (size_t)c->sm_count * TC_M * NP * sizeof(__nv_bfloat16);
  if (!c->tc_x) {
    if (cudaMalloc((void **)&c->tc_x, 2 * xbytes) != cudaSuccess) return DNBC_ENOMEM;
  }
  if (c->tc_mrow_cap < np) {
    if (c->tc_mrow) cudaFree(c->tc_mrow);
    c->tc_mrow = nullptr;
    if (cudaMalloc((void **)&c->tc_mrow, (size_t)np * sizeof(float)) != cudaSuccess) return DNBC_ENOMEM;
    c->tc_mrow_cap = np;
  }
  const unsigned char *t = (const unsigned char *)c->tc_tab;
  StatsLayout sl = stats_layout(c->dm);
  TcArgs a{N, NP, c->d_off, s0, s1, p0, c->E, c->tc_mrow, c->AL, c->GA, c->U, c->cs, c->d_stats, sl.pi, c->t_pi,
           backward ? t : t + tbytes, (unsigned char *)c->tc_x};
  const size_t smem = 2 * STAGE_BYTES + 1024 + (TC_EPI / 32) * TB_BYTES;   // ring + alignment slack + transpose tiles
  if (!backward) {
    c->launches++;
    k_rowmax<<<c->sm_count * 8, 256, 0, c->stream>>>(c->E, N, np, c->tc_mrow);
  }
  // (the epilogue addresses points with 32-bit indices relative to the chunk: chunks stay below 2^31 points)
  auto kf = backward ? (N % 4 == 0 ? k_fb_tc<1, true> : k_fb_tc<1, false>) : (N % 4 == 0 ? k_fb_tc<0, true> : k_fb_tc<0, false>);
  cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kf<<<grid, TC_THREADS, smem, c->stream>>>(a);
  return DNBC_OK;
}
