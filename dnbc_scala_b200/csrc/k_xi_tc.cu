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
// math.  One CTA owns an output tile of 128 x NJ states and a slab of points:
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
constexpr int XA_TILE = XM * XKB * 2;

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

// rows [0, rows) of a K-major tile: row r = state col0 + r, K index = point p + k
__device__ __forceinline__ void stage_tile(uint32_t tile_hi, uint32_t tile_lo, const float *__restrict__ src, int N, int col0,
                                           int valid, int rows, int64_t p, int64_t pe) {
  for (int q = threadIdx.x; q < rows * 8; q += XT) {
    const int r = q % rows, c = q / rows;
    const int64_t pp = p + c * 8;
    float x[8];
    if (r < valid && pp + 8 <= pe) {
      const float *g = src + pp * N + col0 + r;
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = __ldg(g + (int64_t)k * N);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = (r < valid && pp + k < pe) ? __ldg(src + (pp + k) * N + col0 + r) : 0.f;
    }
    uint4 hi, lo;
    split8(x, hi, lo);
    const uint32_t o = sw128(r, c);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile_hi + o), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile_lo + o), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
  }
}

template <int NJ>   // UMMA N = to-states per tile (64, 128 or 256)
__global__ void __launch_bounds__(XT) k_xi_tc(XiArgs a) {
  constexpr int B_TILE = NJ * XKB * 2;
  constexpr int STAGE = 2 * XA_TILE + 2 * B_TILE;
  constexpr int NCOL = NJ / 2;                               // accumulator columns per thread
  extern __shared__ unsigned char smraw[];
  __shared__ __align__(8) unsigned long long bars[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int N = a.N;
  const int tiles = a.mtiles * a.ntiles;
  const int tile = blockIdx.x % tiles, slab = blockIdx.x / tiles;
  const int i0 = (tile % a.mtiles) * XM, j0 = (tile / a.mtiles) * 256;
  const int vi = N - i0 < XM ? N - i0 : XM, vj = N - j0 < NJ ? N - j0 : NJ;
  const int rows_i = (vi + 31) / 32 * 32, rows_j = (vj + 31) / 32 * 32 < NJ ? (vj + 31) / 32 * 32 : NJ;
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
  // rows that are never staged (states past N inside the 128 x NJ tile) stay zero
  for (int o = tid * 16; o < 2 * STAGE; o += XT * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(smem_base + o), "r"(0u) : "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int row = (w & 3) * 32 + lane, col0 = (w >> 2) * NCOL;     // TMEM lane quadrant = warp % 4
  const uint32_t trow = tmem + ((uint32_t)((w & 3) * 32) << 16) + col0;
  const uint32_t idesc = umma_idesc(XM, NJ);

  float acc[NCOL];
#pragma unroll
  for (int i = 0; i < NCOL; ++i) acc[i] = 0.f;
  uint32_t uses[2] = {0u, 0u};
  int inwin = 0, it = 0;
  for (int64_t p = pb; p < pe; p += XKB, ++it) {
    const int s = it & 1;
    const uint32_t st = smem_base + s * STAGE;
    // the stage is free once the MMAs that read it two blocks ago have completed
    if (uses[s] > 0) mbar_wait(bar0 + 8 * s, (uses[s] - 1) & 1);
    stage_tile(st, st + XA_TILE, a.AL, N, i0, vi, rows_i, p, pe);
    stage_tile(st + 2 * XA_TILE, st + 2 * XA_TILE + B_TILE, a.U, N, j0, vj, rows_j, p, pe);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t dAh = umma_desc(st), dAl = umma_desc(st + XA_TILE);
      const uint64_t dBh = umma_desc(st + 2 * XA_TILE), dBl = umma_desc(st + 2 * XA_TILE + B_TILE);
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
#pragma unroll
      for (int ch = 0; ch < NCOL / 32; ++ch) {
        float d[32];
        tmem_ld32(trow + ch * 32, d);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[ch * 32 + i] += d[i];
      }
      inwin = 0;
      tc_fence_before();
      __syncthreads();
    }
  }
  const int i = i0 + row;
  if (i < N) {
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

template <int NJ>
void launch_t(dnbc_ctx *c, XiArgs a, int ctas_per_sm) {
  constexpr int STAGE = 2 * XA_TILE + 2 * NJ * XKB * 2;
  const size_t smem = 2 * STAGE + 1024;
  const int tiles = a.mtiles * a.ntiles;
  int64_t slabs = (int64_t)c->sm_count * ctas_per_sm / tiles;
  const int64_t maxslabs = (a.np + XKB * XWIN - 1) / (XKB * XWIN);
  if (slabs > maxslabs) slabs = maxslabs;
  if (slabs < 1) slabs = 1;
  a.slabs = (int)slabs;
  cudaFuncSetAttribute(k_xi_tc<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_xi_tc<NJ><<<(unsigned)(slabs * tiles), XT, smem, c->stream>>>(a);
}

}  // namespace

bool xi_tc_supported(const dnbc_ctx *c) {
  static const bool off = getenv("DNBC_NO_TC") != nullptr;
  return !off && c->dm.N >= 64;
}

void launch_xi_tc(dnbc_ctx *c, int64_t np) {
  const int N = c->dm.N;
  StatsLayout sl = stats_layout(c->dm);
  XiArgs a{N, c->dm.Npad, c->AL, c->U, np, c->t_A, c->d_stats + sl.xi, (N + XM - 1) / XM, (N + 255) / 256, 1};
  if (N <= 64) launch_t<64>(c, a, 2);
  else if (N <= 128) launch_t<128>(c, a, 1);
  else launch_t<256>(c, a, 1);
}
