// k_xi_tc.cu -- transition statistics on the 5th-generation tensor cores (state spaces >= 64):
//
//     xi[i][j] = a_ij * sum_p alpha^_p(i) u_p(j)            
//
// the soft-count generalisation of the reference's learnTransitions
// (core/src/main/scala/DynamicNaiveBayesianClassifier.scala:204-216, count[from][to] += 1);
// parity is against the fp64 oracle (oracle/dnbc_oracle.c) and, through the clamped-label bridge
// test, against the reference's counts.
//
// The sum over points is a GEMM  xi = AL^T [N x P] * U [P x N]  with a huge reduction dimension
// and a tiny output, so it is bound by reading AL and U once (2 * 4N bytes per point), not by
// math.  One CTA owns an output tile of 128 x NJ states (NJ = 64 or 128) and a slab of points:
//   * 256 threads read fp32 rows of AL / U (coalesced along the state index), split every value
//     into bf16 hi + lo (x = hi + lo to ~2^-17) and write both TRANSPOSED into K-major,
//     128-byte-swizzled operand tiles (one conflict-free 16-byte store per 8 points);
//   * one thread issues tcgen05.mma (hi*hi + hi*lo + lo*hi into one fp32 TMEM accumulator) and
//     commits to the stage's mbarrier; two stages, so staging block k+1 overlaps the MMAs of k;
//   * three-level accumulation: TMEM over a window of 512 points, fp32 registers (round to
//     nearest) over the CTA's slab, one fp64 atomicAdd per output element per CTA at the end.
#include <cuda_bf16.h>

#include "dnbc_internal.cuh"
#include "tc_common.cuh"

namespace {

constexpr int XT = 256;     // threads per CTA
constexpr int XKB = 64;     // points per K block = one 128-byte swizzled row of bf16
constexpr int XWIN = 8;     // K blocks accumulated in TMEM before the window is folded into registers
constexpr int XM = 128;     // UMMA M (from-states per tile; rows past N are zero)

struct XiArgs {
  int N, Npad;
  const float *AL, *U;      // [np][N]
  int64_t np;
  const float *Amat;        // [N][Npad]
  double *xi;               // [N][N]
  int mtiles, ntiles, slabs;
};

// 8 consecutive points of one state column -> one 16-byte chunk of bf16 hi and of bf16 lo
__device__ __forceinline__ void split8(const float (&x)[8], uint4 &hi, uint4 &lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
    const float2 hf = __bfloat1622float2(hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
    h[i] = *(const uint32_t *)&hh;
    l[i] = *(const uint32_t *)&ll;
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// One staging item = 8 consecutive points of one state column = one 16-byte chunk of a K-major
// tile row.  A thread's items are the same in every K block, so their coordinates are computed
// once: `g` = element offset from the block's first point, `so` = swizzled byte offset in the tile.
struct Item {
  int g, k0;                    // k0 = first point of the chunk inside the K block
  uint32_t so;
  bool on;
};
// Items of states past N inside a partial tile load the tile's last real column again (lines their neighbours fetch
// anyway) and store nothing: their rows of the operand tiles keep the zeros of the initial fill.  They used to walk
// the guarded tail path -- an integer division and eight predicated loads per item and K block -- which is twice the
// instructions of the plain path, so the warps WITHOUT work set the pace of every CTA with a partial tile: N = 320
// took 7.7 ms per 2.5e6 points where N = 384 (nine full tiles) takes 4.2, with 5.97 GB of DRAM reads for 3.07 GB of
// rows (ncu: the slower CTAs of a slab fall out of the L2 window of the faster ones).
__device__ __forceinline__ Item make_item(int q, int rows, int valid, int N, int col0) {
  Item it;
  const int r = q % rows, c = q / rows;
  it.on = r < valid;
  it.g = c * 8 * N + col0 + (r < valid ? r : valid - 1);
  it.k0 = c * 8;
  it.so = sw128(r, c);
  return it;
}
__device__ __forceinline__ void load_item(const Item &it, const float *__restrict__ blk, int N, int64_t left, float (&x)[8]) {
  const float *g = blk + it.g;
  if (left >= XKB) {                                         // whole block inside the slab
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = __ldg(g + (int64_t)k * N);
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = it.k0 + k < left ? __ldg(g + (int64_t)k * N) : 0.f;
  }
}
__device__ __forceinline__ void store_item(const Item &it, uint32_t tile_hi, uint32_t tile_lo, const float (&x)[8]) {
  uint4 hi, lo;
  split8(x, hi, lo);
  if (it.on) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile_hi + it.so), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile_lo + it.so), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
  }
}

// NJ = UMMA N = to-states per tile (64 or 128); AM = rows of the A tile that exist in shared
// memory.  The MMA is always M = 128: with AM = 64 (N <= 64) rows 64..127 of the operand are
// whatever follows the tile in shared memory, which only produces accumulator rows 64..127 that
// nobody reads -- the product is row-independent -- and halves the A tiles, so 3 CTAs fit an SM.
template <int NJ, int AM>
#ifndef DNBC_XI_OCC
#define DNBC_XI_OCC 2
#endif
__global__ void __launch_bounds__(XT, NJ == 64 ? DNBC_XI_OCC : 1) k_xi_tc(XiArgs a) {
  constexpr int A_TILE = AM * XKB * 2, B_TILE = NJ * XKB * 2;
  constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;
  constexpr int NCOL = NJ / 2;                               // accumulator columns per thread
  constexpr int IA = AM * 8 / XT, IB = NJ * 8 / XT;          // staging items per thread
  constexpr bool PREFETCH = (IA + IB) * 8 <= 64;             // next block's rows ride in registers
  extern __shared__ unsigned char smraw[];
  __shared__ __align__(8) unsigned long long bars[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int N = a.N;
  const int tiles = a.mtiles * a.ntiles;
  const int tile = blockIdx.x % tiles, slab = blockIdx.x / tiles;
  const int i0 = (tile % a.mtiles) * XM, j0 = (tile / a.mtiles) * NJ;
  const int vi = N - i0 < AM ? N - i0 : AM, vj = N - j0 < NJ ? N - j0 : NJ;
  const int64_t per = ((a.np + a.slabs - 1) / a.slabs + XKB - 1) / XKB * XKB;
  const int64_t pb = (int64_t)slab * per;
  const int64_t pe = pb + per < a.np ? pb + per : a.np;

  const uint32_t smem_base = (smem_u32(smraw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(&bars[0]);
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(NJ));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // rows that are never staged (states past N inside the tile) stay zero; so does the pad after the
  // last tile that an AM = 64 operand reads as its rows 64..127
  for (int o = tid * 16; o < 2 * STAGE + (XM - AM) * 128; o += XT * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(smem_base + o), "r"(0u) : "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int row = (w & 3) * 32 + lane, col0 = (w >> 2) * NCOL;     // TMEM lane quadrant = warp % 4
  const uint32_t trow = tmem + ((uint32_t)((w & 3) * 32) << 16) + col0;
  const uint32_t idesc = umma_idesc(XM, NJ);
  const bool rows_live = (w & 3) * 32 < vi;                         // this warp's accumulator rows are real states

  Item ia[IA], ib[IB];
#pragma unroll
  for (int n = 0; n < IA; ++n) ia[n] = make_item(tid + n * XT, AM, vi, N, i0);
#pragma unroll
  for (int n = 0; n < IB; ++n) ib[n] = make_item(tid + n * XT, NJ, vj, N, j0);

  float acc[NCOL];
#pragma unroll
  for (int i = 0; i < NCOL; ++i) acc[i] = 0.f;
  float xa[PREFETCH ? IA : 1][8], xb[PREFETCH ? IB : 1][8];
  if (PREFETCH && pb < pe) {
#pragma unroll
    for (int n = 0; n < IA; ++n) load_item(ia[n], a.AL + pb * N, N, pe - pb, xa[n]);
#pragma unroll
    for (int n = 0; n < IB; ++n) load_item(ib[n], a.U + pb * N, N, pe - pb, xb[n]);
  }
  uint32_t uses[2] = {0u, 0u};
  int inwin = 0, it = 0;
  for (int64_t p = pb; p < pe; p += XKB, ++it) {
    const int s = it & 1;
    const uint32_t st = smem_base + s * STAGE;
    // the stage is free once the MMAs that read it two blocks ago have completed
    if (uses[s] > 0) mbar_wait(bar0 + 8 * s, (uses[s] - 1) & 1);
    if (PREFETCH) {
#pragma unroll
      for (int n = 0; n < IA; ++n) store_item(ia[n], st, st + A_TILE, xa[n]);
#pragma unroll
      for (int n = 0; n < IB; ++n) store_item(ib[n], st + 2 * A_TILE, st + 2 * A_TILE + B_TILE, xb[n]);
      if (p + XKB < pe) {                                     // next block's loads fly during this block's MMAs
#pragma unroll
        for (int n = 0; n < IA; ++n) load_item(ia[n], a.AL + (p + XKB) * N, N, pe - p - XKB, xa[n]);
#pragma unroll
        for (int n = 0; n < IB; ++n) load_item(ib[n], a.U + (p + XKB) * N, N, pe - p - XKB, xb[n]);
      }
    } else {
#pragma unroll
      for (int n = 0; n < IA; ++n) {
        float x[8];
        load_item(ia[n], a.AL + p * N, N, pe - p, x);
        store_item(ia[n], st, st + A_TILE, x);
      }
#pragma unroll
      for (int n = 0; n < IB; ++n) {
        float x[8];
        load_item(ib[n], a.U + p * N, N, pe - p, x);
        store_item(ib[n], st + 2 * A_TILE, st + 2 * A_TILE + B_TILE, x);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t dAh = umma_desc(st), dAl = umma_desc(st + A_TILE);
      const uint64_t dBh = umma_desc(st + 2 * A_TILE), dBl = umma_desc(st + 2 * A_TILE + B_TILE);
#pragma unroll
      for (int k16 = 0; k16 < XKB / 16; ++k16) {
        const uint64_t ko = (uint64_t)(k16 * 32 >> 4);       // 16 bf16 = 32 B further along K
        umma_bf16(tmem, dAh + ko, dBh + ko, idesc, (inwin | k16) ? 1u : 0u);
        umma_bf16(tmem, dAh + ko, dBl + ko, idesc, 1u);
        umma_bf16(tmem, dAl + ko, dBh + ko, idesc, 1u);
      }
      umma_commit(bar0 + 8 * s);
    }
    uses[s] += 1;
    if (++inwin == XWIN || p + XKB >= pe) {
      // fold the window into the register accumulators (fp32, round to nearest)
      mbar_wait(bar0 + 8 * s, (uses[s] - 1) & 1);
      tc_fence_after();
      if (rows_live) {
#pragma unroll
        for (int ch = 0; ch < NCOL / 32; ++ch) {
          float d[32];
          tmem_ld32(trow + ch * 32, d);
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[ch * 32 + i] += d[i];
        }
      }
      inwin = 0;
      tc_fence_before();
      __syncthreads();
    }
  }
  const int i = i0 + row;
  if (i < N && rows_live) {
#pragma unroll
    for (int c = 0; c < NCOL; ++c) {
      const int j = j0 + col0 + c;
      if (j < N && acc[c] != 0.f) atomicAdd(a.xi + (int64_t)i * N + j, (double)acc[c] * (double)a.Amat[(int64_t)i * a.Npad + j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(NJ));
}

template <int NJ, int AM>
void launch_t(dnbc_ctx *c, XiArgs a, int ctas_per_sm) {
  constexpr int STAGE = 2 * AM * XKB * 2 + 2 * NJ * XKB * 2;
  const size_t smem = 2 * STAGE + (XM - AM) * 128 + 1024;
  const int tiles = a.mtiles * a.ntiles;
  int64_t slabs = (int64_t)c->sm_count * ctas_per_sm / tiles;
  const int64_t maxslabs = (a.np + XKB * XWIN - 1) / (XKB * XWIN);
  if (slabs > maxslabs) slabs = maxslabs;
  if (slabs < 1) slabs = 1;
  a.slabs = (int)slabs;
  cudaFuncSetAttribute(k_xi_tc<NJ, AM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_xi_tc<NJ, AM><<<(unsigned)(slabs * tiles), XT, smem, c->stream>>>(a);
}

}  // namespace

bool xi_tc_supported(const dnbc_ctx *c) {
  return c->dm.N >= 64;
}

void launch_xi_tc(dnbc_ctx *c, int64_t np) {
  const int N = c->dm.N;
  StatsLayout sl = stats_layout(c->dm);
  XiArgs a{N, c->dm.Npad, c->AL, c->U, np, c->t_A, c->d_stats + sl.xi, (N + XM - 1) / XM, 1, 1};
  if (N <= 64) launch_t<64, 64>(c, a, DNBC_XI_OCC);
  else {
    // 128 x 128 tiles: with 128 x 256 the 12 staging items per thread cannot be prefetched in
    // registers and the load latency is exposed (measured 30 vs 11.7 ms per 4e6 points at N = 512)
    a.ntiles = (N + 127) / 128;
    launch_t<128, 128>(c, a, 1);
  }
}
