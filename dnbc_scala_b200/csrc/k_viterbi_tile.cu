// k_viterbi_tile.cu -- K5 for large state spaces (padded N = 256, 512, ...): fp32 back-traced
// Viterbi as a batched max-plus matrix product.
//
// Reference function it replaces: viterbiInitialize + viterbiCompute,
// core/src/main/scala/DynamicNaiveBayesianClassifier.scala:49-91, in BACKTRACE mode (a true
// Viterbi; REFERENCE mode with the reference's quirks stays on k_large_n.cu).  Same recursion,
// first-maximum tie rule, tie flags and outputs as k_decode_block<float, false>.
//
// A step for a batch of sequences is the tropical product
//     delta_t[b][j] = e_t[b][j] + max_i ( delta_{t-1}[b][i] + ln a_ij ),
// i.e. a [32 x N] (x) [N x N] "GEMM" with (max, +) for (+, *).  It cannot use the tensor cores,
// so it is laid out like a register-tiled SGEMM on the FP32 / ALU pipes:
//   * one CTA (512 threads) owns 32 sequences; delta lives in shared memory (double-buffered);
//   * ln A streams from L2 through a 2-stage cp.async ring of [32 sources] x [256 destinations]
//     tiles -- read once per step per 32 sequences instead of once per sequence;
//   * a thread owns 4 sequences x 4 destinations; per block of 4 sources it forms 64 candidates
//     with packed adds, reduces each destination's 4 candidates with 3-input max, and keeps a
//     running (best, runner-up, block index) per destination: 2.4 issue slots per candidate
//     instead of ~6 for compare-and-select on every candidate;
//   * the arg-max inside the winning block of 4 (and the runner-up inside it, for the tie flag) is
//     recovered afterwards by re-forming just those 4 candidates (bit-identical adds);
//   * exact pruning of source blocks (round 2): ncu shows the kernel bound by the ALU pipe (max / compare / select:
//     76 % busy, FMA pipe 15 %), so the only way down is fewer candidates.  With LB = max_i (delta(i) + min_j ln a_ij),
//     every destination's best candidate is >= LB; a source with delta(i) + max_j ln a_ij < LB - margin can neither
//     win any destination nor come within the tie tolerance of a winner.  Each step starts by turning that test into
//     one bit per block of 4 sources and sequence; the hot loop skips dead blocks (the test is warp-uniform: a warp's
//     threads share their 4 sequences).  Paths, scores and tie flags are unchanged (same CRC over 2.5e7 points); on
//     the reference generator's data at N = 512 about 10 % of the sources are live, i.e. a third of the blocks;
//   * the back-trace is k_backtrace (k_decode.cu): one thread per sequence.
#include "dnbc_internal.cuh"
#include "packed_f32.cuh"

#define DNBC_TIE_TOL_F32 1e-3f

namespace {

constexpr int VT_THREADS = 512;
constexpr int VT_BS = 32;     // sequences per CTA
constexpr int VT_JC = 256;    // destinations per pass
constexpr int VT_IT = 32;     // sources per shared-memory tile

// (lo, hi) + (s, s)
__device__ __forceinline__ void add2s(float lo, float hi, float s, float &olo, float &ohi) {
  unpk(add2(pk(lo, hi), pk(s, s)), olo, ohi);
}
__device__ __forceinline__ void cp16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
}

struct VtArgs {
  int N, NP;
  const int64_t *off;
  int64_t s0, s1, p0;
  const float *E;        // [np][N]
  const float *logpi;    // [N]
  const float *LA;       // [NP][NP] padded ln A, -inf outside the active N x N block
  const float *rowmax;   // [NP] max_j ln a_ij over the active destinations (-inf for an inactive source)
  const float *rowmin;   // [NP] min_j ln a_ij over the active destinations (-inf as soon as one transition is impossible)
  const uint8_t *active;
  uint16_t *path;
  double *score;
  uint8_t *tie;
  uint16_t *bp;
};

// ln A -> [NP][NP], -inf for padded or inactive states
__global__ void k_pad_logA(const float *__restrict__ logA, int lda, const uint8_t *__restrict__ active, int N, int NP,
                           float *__restrict__ out) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < (int64_t)NP * NP; q += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(q / NP), j = (int)(q % NP);
    out[q] = (i < N && j < N && active[i] && active[j]) ? logA[(int64_t)i * lda + j] : -INFINITY;
  }
}

// row extrema of the padded matrix over the active destinations (one warp per source row)
__global__ void k_row_extrema(const float *__restrict__ LA, const uint8_t *__restrict__ active, int N, int NP,
                              float *__restrict__ rowmax, float *__restrict__ rowmin) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= NP) return;
  float mx = -INFINITY, mn = INFINITY;
  if (warp < N && active[warp])
    for (int j = lane; j < N; j += 32)
      if (active[j]) {
        const float v = LA[(int64_t)warp * NP + j];
        mx = fmaxf(mx, v);
        mn = fminf(mn, v);
      }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if (lane == 0) {
    rowmax[warp] = mx;
    rowmin[warp] = mn == INFINITY ? -INFINITY : mn;
  }
}

// candidates more than this below a destination's lower bound are pruned (four times the tie tolerance)
#define DNBC_PRUNE_MARGIN (4.f * DNBC_TIE_TOL_F32)

__global__ void __launch_bounds__(VT_THREADS, 1) k_viterbi_tile(VtArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int N = a.N, NP = a.NP;
  float *dbuf = (float *)smraw;                              // [2][VT_BS][NP]
  float *Lt = dbuf + 2 * VT_BS * NP;                         // [2][VT_IT][VT_JC]
  double *sshift = (double *)(Lt + 2 * VT_IT * VT_JC);       // [VT_BS]
  int64_t *sb = (int64_t *)(sshift + VT_BS);                 // [VT_BS] chunk-relative first point
  int64_t *sgb = sb + VT_BS;                                 // [VT_BS] global first point
  int *sT = (int *)(sgb + VT_BS);                            // [VT_BS]
  int *sflag = sT + VT_BS;                                   // [VT_BS] tie seen in this step
  uint8_t *slive = (uint8_t *)(sflag + VT_BS);               // [VT_BS][NP / 32] one bit per block of 4 sources: may still win
  const int nlw = NP / 32;
  __shared__ int sTmax, sna;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int bq0 = (w >> 1) * 4;                              // this thread's 4 sequences
  const int jl = ((w & 1) * 32 + lane) * 4;                  // and its 4 destinations inside a pass
  const uint32_t Lt_s = (uint32_t)__cvta_generic_to_shared(Lt);
  const int ntile_i = NP / VT_IT, npass = NP / VT_JC, G = ntile_i * npass;

  if (tid == 0) sna = 0;
  __syncthreads();
  for (int j = tid; j < N; j += VT_THREADS)
    if (a.active[j]) atomicAdd(&sna, 1);
  __syncthreads();
  const bool multi = sna >= 2;

  // tile g = (pass, source tile): [VT_IT][VT_JC] floats, 4 x 16 B per thread
  auto load_tile = [&](int g, int stage) {
    const int jc = g / ntile_i, it = g % ntile_i;
    const float *src = a.LA + (int64_t)(it * VT_IT) * NP + jc * VT_JC;
    const uint32_t dst = Lt_s + stage * (VT_IT * VT_JC * 4);
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int q = tid + n * VT_THREADS;                    // 16-byte chunk index: 64 per row
      const int r = q >> 6, c = q & 63;
      cp16(dst + (r * VT_JC + c * 4) * 4, src + (int64_t)r * NP + c * 4);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // live-block bits of this warp's two sequences from the delta rows in `rows` (any common shift of a row is fine)
  auto make_masks = [&](const float *rows) {
    for (int b = w * 2; b < w * 2 + 2; ++b) {
      const float *d = rows + b * NP;
      float lb = -INFINITY;
      for (int j = lane; j < NP; j += 32) lb = fmaxf(lb, d[j] + __ldg(a.rowmin + j));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lb = fmaxf(lb, __shfl_xor_sync(0xffffffffu, lb, o));
      // (lb = -inf -- an impossible transition in every row that carries mass -- keeps every source that has mass)
      const float thr = lb - DNBC_PRUNE_MARGIN;
      for (int j0 = 0; j0 < NP; j0 += 32) {
        const float v = d[j0 + lane] + __ldg(a.rowmax + j0 + lane);
        const unsigned bal = __ballot_sync(0xffffffffu, v >= thr && v > -INFINITY);
        if (lane == 0) {
          unsigned bits = 0;
#pragma unroll
          for (int q = 0; q < 8; ++q) bits |= ((bal >> (4 * q)) & 0xfu) ? (1u << q) : 0u;
          slive[b * nlw + (j0 >> 5)] = (uint8_t)bits;
        }
      }
    }
  };

  for (int64_t s0 = a.s0 + (int64_t)blockIdx.x * VT_BS; s0 < a.s1; s0 += (int64_t)gridDim.x * VT_BS) {
    __syncthreads();
    if (tid == 0) sTmax = 0;
    __syncthreads();
    if (tid < VT_BS) {
      const int64_t s = s0 + tid;
      const bool have = s < a.s1;
      const int T = have ? (int)(a.off[s + 1] - a.off[s]) : 0;
      sT[tid] = T;
      sb[tid] = have ? a.off[s] - a.p0 : 0;
      sgb[tid] = have ? a.off[s] : 0;
      sshift[tid] = 0.0;
      atomicMax(&sTmax, T);
      if (have && T <= 0 && a.score) a.score[s] = 0.0;
    }
    __syncthreads();
    const int Tmax = sTmax;
    for (int q = tid; q < VT_BS * NP; q += VT_THREADS) {
      const int b = q / NP, j = q % NP;
      const bool on = j < N && a.active[j] != 0 && sT[b] > 0;
      dbuf[q] = on ? a.E[sb[b] * N + j] + a.logpi[j] : -INFINITY;
    }
    if (Tmax > 1) load_tile(0, 0);
    __syncthreads();
    make_masks(dbuf);
    __syncthreads();

    for (int t = 1; t < Tmax; ++t) {
      const float *src = dbuf + ((t - 1) & 1) * VT_BS * NP;
      float *dst = dbuf + (t & 1) * VT_BS * NP;
      if (tid < VT_BS) sflag[tid] = 0;
      for (int jc = 0; jc < npass; ++jc) {
        float best[4][4], second[4][4];
        int argb[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int x = 0; x < 4; ++x) { best[q][x] = -INFINITY; second[q][x] = -INFINITY; argb[q][x] = 0; }
        for (int it = 0; it < ntile_i; ++it) {
          const int g = jc * ntile_i + it;
          // tile g has landed; start g + 1 (the tile sequence repeats every step) into the other stage,
          // which every thread finished reading before the barrier below
          asm volatile("cp.async.wait_group 0;" ::: "memory");
          __syncthreads();
          if (g + 1 < G || t + 1 < Tmax) load_tile((g + 1) % G, (g + 1) & 1);
          const float *L = Lt + (g & 1) * (VT_IT * VT_JC) + jl;
          const float *d0 = src + bq0 * NP + it * VT_IT;
          // live bits of this tile's 8 blocks for the thread's four sequences (warp-uniform)
          unsigned lv[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) lv[q] = slive[(bq0 + q) * nlw + it];
          const unsigned lany = lv[0] | lv[1] | lv[2] | lv[3];
#pragma unroll 2
          for (int ib = 0; ib < VT_IT / 4; ++ib) {
            if (!((lany >> ib) & 1u)) continue;                // no sequence of this warp can use these 4 sources
            float4 Lr[4], dl[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) Lr[k] = *(const float4 *)(L + (ib * 4 + k) * VT_JC);
#pragma unroll
            for (int q = 0; q < 4; ++q) dl[q] = *(const float4 *)(d0 + q * NP + ib * 4);
            const int blk = it * (VT_IT / 4) + ib;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (!((lv[q] >> ib) & 1u)) continue;
              float c[4][4];                                 // [source k][destination x]
              add2s(Lr[0].x, Lr[0].y, dl[q].x, c[0][0], c[0][1]);
              add2s(Lr[0].z, Lr[0].w, dl[q].x, c[0][2], c[0][3]);
              add2s(Lr[1].x, Lr[1].y, dl[q].y, c[1][0], c[1][1]);
              add2s(Lr[1].z, Lr[1].w, dl[q].y, c[1][2], c[1][3]);
              add2s(Lr[2].x, Lr[2].y, dl[q].z, c[2][0], c[2][1]);
              add2s(Lr[2].z, Lr[2].w, dl[q].z, c[2][2], c[2][3]);
              add2s(Lr[3].x, Lr[3].y, dl[q].w, c[3][0], c[3][1]);
              add2s(Lr[3].z, Lr[3].w, dl[q].w, c[3][2], c[3][3]);
#pragma unroll
              for (int x = 0; x < 4; ++x) {
                const float m = fmaxf(fmaxf(c[0][x], c[1][x]), fmaxf(c[2][x], c[3][x]));
                second[q][x] = fmaxf(second[q][x], fminf(best[q][x], m));
                argb[q][x] = m > best[q][x] ? blk : argb[q][x];
                best[q][x] = fmaxf(best[q][x], m);
              }
            }
          }
        }
        // ---- winner inside its block of 4 sources, runner-up, new delta, back-pointer
        const int j0 = jc * VT_JC + jl;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int b = bq0 + q;
          const bool live = t < sT[b];
          float dn[4];
          uint16_t ar[4];
          bool onx[4];
          bool tie = false;
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const int j = j0 + x;
            const int i4 = argb[q][x] * 4;
            const float4 dd = *(const float4 *)(src + b * NP + i4);
            const float *la = a.LA + (int64_t)i4 * NP + j;
            const float c0 = dd.x + __ldg(la), c1 = dd.y + __ldg(la + NP), c2 = dd.z + __ldg(la + 2 * NP), c3 = dd.w + __ldg(la + 3 * NP);
            const float bv = best[q][x];
            int k = 3;
            k = c2 == bv ? 2 : k;
            k = c1 == bv ? 1 : k;
            k = c0 == bv ? 0 : k;
            // runner-up: the largest candidate of the block other than the winner, or another block's maximum
            float inb = -INFINITY;
            inb = k != 0 ? fmaxf(inb, c0) : inb;
            inb = k != 1 ? fmaxf(inb, c1) : inb;
            inb = k != 2 ? fmaxf(inb, c2) : inb;
            inb = k != 3 ? fmaxf(inb, c3) : inb;
            const float sec = fmaxf(second[q][x], inb);
            const bool on = j < N && a.active[j] != 0;
            tie = tie || (on && !(bv - sec > DNBC_TIE_TOL_F32));
            dn[x] = on ? bv : -INFINITY;
            onx[x] = on;
            ar[x] = (uint16_t)(i4 + k);
          }
          if (live) {
            const float *e = a.E + (sb[b] + t) * N + j0;
            uint16_t *bp = a.bp + (sb[b] + t) * N + j0;
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              if (onx[x]) {
                dn[x] += e[x];
                bp[x] = ar[x];
              }
            }
            if (tie && multi) sflag[b] = 1;
          }
          *(float4 *)(dst + b * NP + j0) = make_float4(dn[0], dn[1], dn[2], dn[3]);
        }
      }
      __syncthreads();
      // ---- renormalise each live sequence by its maximum (shift carried in fp64); freeze finished ones
      for (int b = w * 2; b < w * 2 + 2; ++b) {
        const bool live = t < sT[b];
        float mx = -INFINITY;
        for (int j = lane; j < NP; j += 32) mx = fmaxf(mx, dst[b * NP + j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const bool shiftit = live && mx > -INFINITY && mx < INFINITY;
        for (int j = lane; j < NP; j += 32) {
          const float v = live ? dst[b * NP + j] - (shiftit ? mx : 0.f) : src[b * NP + j];
          dst[b * NP + j] = v;
        }
        if (lane == 0) {
          if (shiftit) sshift[b] += (double)mx;
          if (live && sflag[b] && a.tie) a.tie[sgb[b] + t] |= 2;
        }
      }
      make_masks(dst);                                       // (each lane re-reads only what it wrote above)
      __syncthreads();
    }
    // ---- last state: first maximum over the active states, tie flag, score
    const float *fin = dbuf + ((Tmax > 0 ? Tmax - 1 : 0) & 1) * VT_BS * NP;
    for (int b = w * 2; b < w * 2 + 2; ++b) {
      const int T = sT[b];
      if (T <= 0) continue;
      float best = -INFINITY, second = -INFINITY;
      int idx = 1 << 30;
      for (int j = lane * (NP / 32); j < (lane + 1) * (NP / 32); ++j) {
        if (j < N && a.active[j]) {
          const float v = fin[b * NP + j];
          if (idx == (1 << 30) || v > best) { second = idx == (1 << 30) ? second : best; best = v; idx = j; }
          else if (v > second) second = v;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, second, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        const bool ov = oi != (1 << 30), mv = idx != (1 << 30);
        if (ov && (!mv || ob > best || (ob == best && oi < idx))) {
          float ns = os;
          if (mv && best > ns) ns = best;
          if (second > ns) ns = second;
          second = ns; best = ob; idx = oi;
        } else if (ov) {
          float ns = second;
          if (ob > ns) ns = ob;
          if (os > ns) ns = os;
          second = ns;
        }
      }
      if (lane == 0) {
        const int last = idx == (1 << 30) ? 0 : idx;
        const int64_t gb = sgb[b];
        if (a.tie && multi && !(best - second > DNBC_TIE_TOL_F32)) a.tie[gb + T - 1] |= 1;
        if (a.score) a.score[s0 + b] = (double)best + sshift[b];
        a.path[gb + T - 1] = (uint16_t)last;                  // k_backtrace starts from here
      }
    }
  }
}

}  // namespace

static size_t vt_smem(int NP) {
  return (size_t)(2 * VT_BS * NP + 2 * VT_IT * VT_JC) * sizeof(float) + VT_BS * (8 + 8 + 8 + 4 + 4) + VT_BS * (NP / 32) + 64;
}
// (the two delta buffers of 32 sequences are 8 NP x 32 bytes: padded state counts above 512 do not fit one SM's shared
// memory and stay on k_decode_block; a third stage of the ln A ring was measured slower, 647 vs 608 ms)
// padded state count of the tiles: a multiple of the 256 destinations of a pass (65 .. 128 states run padded to 256 as
// well: the lane-group kernel of k_decode.cu that used to take them is slower); 0 = not supported
#ifndef VT_MIN_N
#define VT_MIN_N 65
#endif
static int vt_np(const dnbc_ctx *c) {
  const int N = c->dm.N;
  if (N < VT_MIN_N) return 0;
  const int NP = (N + VT_JC - 1) / VT_JC * VT_JC;
  return vt_smem(NP) <= 227 * 1024 ? NP : 0;
}
bool viterbi_tile_supported(const dnbc_ctx *c) { return vt_np(c) != 0; }

int launch_viterbi_tile(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0, const float *E, uint16_t *path, double *score,
                        uint8_t *tie, uint16_t *bp) {
  const int N = c->dm.N, NP = vt_np(c);
  if (!c->vt_logA) {                                         // [NP][NP] padded ln A, then its row maxima and minima
    if (cudaMalloc((void **)&c->vt_logA, ((size_t)NP * NP + 2 * NP) * sizeof(float)) != cudaSuccess) return DNBC_ENOMEM;
  }
  float *rowmax = c->vt_logA + (size_t)NP * NP, *rowmin = rowmax + NP;
  c->launches += 2;
  k_pad_logA<<<c->sm_count * 2, 256, 0, c->stream>>>(c->t_logA, c->dm.Npad, c->d_active, N, NP, c->vt_logA);
  k_row_extrema<<<(NP * 32 + 255) / 256, 256, 0, c->stream>>>(c->vt_logA, c->d_active, N, NP, rowmax, rowmin);
  VtArgs a{N, NP, c->d_off, s0, s1, p0, E, c->t_logpi, c->vt_logA, rowmax, rowmin, c->d_active, path, score, tie, bp};
  const size_t smem = vt_smem(NP);
  const int64_t batches = (s1 - s0 + VT_BS - 1) / VT_BS;
  const int grid = (int)(batches < c->sm_count ? batches : c->sm_count);
  cudaFuncSetAttribute(k_viterbi_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_viterbi_tile<<<grid, VT_THREADS, smem, c->stream>>>(a);
  launch_backtrace(c, s0, s1, p0, bp, path);
  return DNBC_OK;
}
