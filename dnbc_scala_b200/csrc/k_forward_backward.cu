// k_forward_backward.cu -- K2: dispatch of the scaled forward / backward recursions and the FP32 xi
// (expected transition count) contraction for 33 .. 63 states.
//
// The reference has NO forward-backward: its learner is supervised
// counting (DynamicNaiveBayesianClassifier.scala:124-161).  The recursions implement
// SURVEY Appendix B and are checked against the fp64 oracle (oracle/dnbc_oracle.c
// fb_one) and, through the clamped-label bridge test, against the reference's
// supervised counts.
//
// The round-1 recursion kernels that lived here (one lane group per sequence, alpha broadcast by
// shuffles, A from shared memory) are gone: every state count now has a second-generation kernel --
// k_fbs.cu (small even N, lane per sequence), k_fb3.cu (33..64, mma.sync), k_fb2.cu (<= 64, register-
// resident A), k_tc.cu (65..512, tcgen05) -- and only state counts above 512 take the CTA-per-sequence
// kernels of k_large_n.cu.
#include "dnbc_internal.cuh"

static void launch_fb(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0, bool backward) {
  if (s1 <= s0) return;
  FamTimer ft(c, backward ? FAM_BACKWARD : FAM_FORWARD);
  if (fbs_supported(c, s1 - s0)) return launch_fbs(c, s0, s1, p0, backward);   // small even N, many sequences: lane per sequence (k_fbs.cu)
  if (fb3_supported(c)) return launch_fb3(c, s0, s1, p0, backward);   // 33..64 states: mma.sync, 16 sequences per warp (k_fb3.cu)
  if (fb2_supported(c)) return launch_fb2(c, s0, s1, p0, backward);   // up to 64 states: register-resident A (k_fb2.cu)
  if (fb_tc_supported(c)) {                                            // 65..512 states: tcgen05 tensor cores (k_tc.cu)
    launch_fb_tc(c, s0, s1, p0, c->h_off[s1] - c->h_off[s0], backward);
    return;
  }
  launch_fb_block(c, s0, s1, p0, backward);                            // above 512 states: CTA per sequence (k_large_n.cu)
}

void launch_forward(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0) { launch_fb(c, s0, s1, p0, false); }
void launch_backward(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0) { launch_fb(c, s0, s1, p0, true); }

// ---------------------------------------------------------------------------------
// xi contraction: Xi[i][j] = sum_p AL[p][i] * U[p][j]  (a [N x np] x [np x N] product over
// the point axis of the chunk), then xi-stat[i][j] += a_ij * Xi[i][j].
// Register-tiled FP32 on the CUDA cores: block tile (16 TM)^2, thread tile TM^2, 32
// points per shared-memory stage; partial sums are folded into fp64 every 4096 points.
// ---------------------------------------------------------------------------------
template <int TM>
__global__ void __launch_bounds__(256) k_xi(int N, int Npad, const float *__restrict__ AL, const float *__restrict__ U,
                                            int64_t np, const float *__restrict__ Amat, double *__restrict__ xi) {
  constexpr int BT = 16 * TM, KS = 32;
  __shared__ float sA[KS][BT + 4];
  __shared__ float sB[KS][BT + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int i0 = blockIdx.y * BT, j0 = blockIdx.x * BT;
  const int64_t slab = (np + gridDim.z - 1) / gridDim.z;
  const int64_t pb = (int64_t)blockIdx.z * slab;
  const int64_t pe = pb + slab < np ? pb + slab : np;
  float acc[TM][TM];
  double acc64[TM][TM];
#pragma unroll
  for (int x = 0; x < TM; ++x)
#pragma unroll
    for (int y = 0; y < TM; ++y) { acc[x][y] = 0.f; acc64[x][y] = 0.0; }
  int stage = 0;
  for (int64_t p = pb; p < pe; p += KS) {
    for (int q = threadIdx.x; q < KS * BT; q += 256) {
      const int kk = q / BT, cc = q % BT;
      const int64_t pp = p + kk;
      sA[kk][cc] = (pp < pe && i0 + cc < N) ? AL[pp * N + i0 + cc] : 0.f;
      sB[kk][cc] = (pp < pe && j0 + cc < N) ? U[pp * N + j0 + cc] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < KS; ++kk) {
      float av[TM], bv[TM];
#pragma unroll
      for (int x = 0; x < TM; ++x) { av[x] = sA[kk][ty * TM + x]; bv[x] = sB[kk][tx * TM + x]; }
#pragma unroll
      for (int x = 0; x < TM; ++x)
#pragma unroll
        for (int y = 0; y < TM; ++y) acc[x][y] = fmaf(av[x], bv[y], acc[x][y]);
    }
    __syncthreads();
    if (++stage == 128) {
      stage = 0;
#pragma unroll
      for (int x = 0; x < TM; ++x)
#pragma unroll
        for (int y = 0; y < TM; ++y) { acc64[x][y] += (double)acc[x][y]; acc[x][y] = 0.f; }
    }
  }
#pragma unroll
  for (int x = 0; x < TM; ++x)
#pragma unroll
    for (int y = 0; y < TM; ++y) {
      const int i = i0 + ty * TM + x, j = j0 + tx * TM + y;
      const double v = acc64[x][y] + (double)acc[x][y];
      if (i < N && j < N && v != 0.0) atomicAdd(xi + (int64_t)i * N + j, v * (double)Amat[i * Npad + j]);
    }
}

void launch_xi(dnbc_ctx *c, int64_t np) {
  if (np <= 0) return;
  if (fb2_fuses_xi(c)) return;                               // accumulated inside k_backward2 (k_fb2.cu)
  FamTimer ft(c, FAM_XI);
  if (xi_tc_supported(c)) {                                  // N >= 64: tcgen05 GEMM (k_xi_tc.cu)
    launch_xi_tc(c, np);
    return;
  }
  StatsLayout sl = stats_layout(c->dm);
  const int N = c->dm.N;
  if (N <= 16) {
    int z = (int)((np + 4095) / 4096);
    if (z > c->sm_count * 8) z = c->sm_count * 8;
    k_xi<1><<<dim3(1, 1, z), 256, 0, c->stream>>>(N, c->dm.Npad, c->AL, c->U, np, c->t_A, c->d_stats + sl.xi);
  } else {
    const int tiles = (N + 63) / 64;
    int z = (int)((np + 4095) / 4096);
    int zcap = c->sm_count * 4 / (tiles * tiles);
    if (zcap < 1) zcap = 1;
    if (z > zcap) z = zcap;
    k_xi<4><<<dim3(tiles, tiles, z), 256, 0, c->stream>>>(N, c->dm.Npad, c->AL, c->U, np, c->t_A, c->d_stats + sl.xi);
  }
}
