// k_decode.cu -- K5: batched max-product decoder with on-device back-trace.
//
// Reference function it replaces: viterbiInitialize + viterbiCompute,
// core/src/main/scala/DynamicNaiveBayesianClassifier.scala:49-91 (called once per test
// sequence from Performance.Measure, core/src/main/scala/Performance.scala:45-52).
//
// REFERENCE mode reproduces the reference's recursion including both quirks:
//   path[t-1] = argmax_j delta_{t-1}(j)                      (:75, :89 -- no back-trace)
//   i*        = argmax_i delta_{t-1}(i) + ln a_ij            (:85)
//   delta_t(j) = e_t(j) + delta_{t-1}(i*)                    (:84-85 -- ln a is NOT added)
//   score     = sum_j delta_{T-1}(j), 0.0 when T == 1        (:90)
// with Scala's maxBy tie rule (first maximum in iteration order, strict `>`), states
// iterating in dense-index order over the active set (`transitions.keys`).
// BACKTRACE mode is the textbook Viterbi: delta_t(j) = e_t(j) + max_i(delta_{t-1}(i) +
// ln a_ij), u16 back-pointers in HBM, trace walked on the device.
//
// Mapping as in k_forward_backward.cu: a group of LG lanes per sequence, R states per
// lane, delta_{t-1}(i) broadcast by shuffle.  T = float renormalises delta by its max
// every step (the shift is carried in fp64) so 1000-step sequences keep fp32 resolution;
// T = double mirrors the JVM arithmetic operation by operation.
#include "dnbc_internal.cuh"

#define DNBC_TIE_TOL_F32 1e-3f

template <int LG>
__device__ __forceinline__ unsigned dgroup_mask() {
  if (LG == 32) return 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  return ((1u << (LG & 31)) - 1u) << ((lane / LG) * LG);
}

template <typename T>
struct DecArgs {
  Dims dm;
  const int64_t *off;
  int64_t s0, s1, p0;
  const T *E;         // [np][N] (chunk-relative)
  const T *logpi;     // [N]
  const T *logA;      // [N][lda]
  int lda;
  const uint8_t *active;
  uint16_t *path;     // [P] global point index
  double *score;      // [S] global sequence index
  uint8_t *tie;       // [P] or null
  uint16_t *bp;       // [np][N] (BACKTRACE)
};

template <typename T> __device__ __forceinline__ T neg_inf();
template <> __device__ __forceinline__ float neg_inf<float>() { return -INFINITY; }
template <> __device__ __forceinline__ double neg_inf<double>() { return -(double)INFINITY; }

template <typename T>
__device__ __forceinline__ bool is_tie(T best, T second) {
  if (sizeof(T) == 8) return second == best;
  return !(best - second > (T)DNBC_TIE_TOL_F32);
}

// first maximum over the group's active states: returns index; *tie per is_tie
template <typename T, int LG, int R>
__device__ __forceinline__ int group_first_argmax(const T (&delta)[R], const bool (&act)[R], int lane, unsigned mask,
                                                  bool *tie) {
  T best = neg_inf<T>(), second = neg_inf<T>();
  int idx = 1 << 30;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = lane + LG * r;
    if (act[r]) {
      if (idx == (1 << 30) || delta[r] > best) { second = idx == (1 << 30) ? second : best; best = delta[r]; idx = j; }
      else if (delta[r] > second) second = delta[r];
    }
  }
#pragma unroll
  for (int o = LG / 2; o > 0; o >>= 1) {
    const T ob = __shfl_xor_sync(mask, best, o, LG);
    const T os = __shfl_xor_sync(mask, second, o, LG);
    const int oi = __shfl_xor_sync(mask, idx, o, LG);
    const bool other_valid = oi != (1 << 30), me_valid = idx != (1 << 30);
    if (other_valid && (!me_valid || ob > best || (ob == best && oi < idx))) {
      // other wins: my best becomes a candidate for second
      T ns = os;
      if (me_valid && best > ns) ns = best;
      if (second > ns) ns = second;
      second = ns; best = ob; idx = oi;
    } else if (other_valid) {
      T ns = second;
      if (ob > ns) ns = ob;
      if (os > ns) ns = os;
      second = ns;
    }
  }
  *tie = is_tie<T>(best, second);  // caller masks this when fewer than two states are active
  return idx == (1 << 30) ? 0 : idx;
}

template <typename T, int LG, int R, bool REFMODE>
__global__ void __launch_bounds__(256) k_decode(DecArgs<T> a) {
  const int N = a.dm.N;
  const int lane = threadIdx.x % LG;
  const int gpb = blockDim.x / LG;
  const unsigned mask = dgroup_mask<LG>();
  const int gshift = ((threadIdx.x & 31) / LG) * LG;
  bool act[R];
  unsigned abits[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = lane + LG * r;
    act[r] = j < N && a.active[j] != 0;
    abits[r] = (__ballot_sync(0xffffffffu, act[r]) >> gshift) & (LG == 32 ? 0xffffffffu : ((1u << LG) - 1u));
  }
  int na = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) na += __popc(abits[r]);
  const bool multi = na >= 2;  // a tie needs two candidates
  for (int64_t s = a.s0 + (int64_t)blockIdx.x * gpb + threadIdx.x / LG; s < a.s1; s += (int64_t)gridDim.x * gpb) {
    const int64_t b = a.off[s] - a.p0;          // chunk-relative first point
    const int64_t gb = a.off[s];                // global first point
    const int64_t Tn = a.off[s + 1] - a.off[s];
    if (Tn <= 0) { if (lane == 0 && a.score) a.score[s] = 0.0; continue; }
    T delta[R];
    double shift = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int j = lane + LG * r;
      delta[r] = act[r] ? a.E[b * N + j] + a.logpi[j] : neg_inf<T>();
    }
    for (int64_t t = 1; t < Tn; ++t) {
      uint8_t tflag_prev = 0, tflag_cur = 0;
      if (REFMODE) {
        bool tie0;
        const int pj = group_first_argmax<T, LG, R>(delta, act, lane, mask, &tie0);
        if (lane == 0) a.path[gb + t - 1] = (uint16_t)pj;
        tflag_prev = (tie0 && multi) ? 1 : 0;
      }
      T best[R], second[R], pval[R];
      int arg[R];
#pragma unroll
      for (int r = 0; r < R; ++r) { best[r] = neg_inf<T>(); second[r] = neg_inf<T>(); pval[r] = neg_inf<T>(); arg[r] = -1; }
#pragma unroll
      for (int rp = 0; rp < R; ++rp) {
#pragma unroll 4
        for (int lp = 0; lp < LG; ++lp) {
          const T di = __shfl_sync(mask, delta[rp], lp, LG);
          if ((abits[rp] >> lp) & 1u) {
            const int i = lp + LG * rp;
            const T *row = a.logA + (int64_t)i * a.lda + lane;
#pragma unroll
            for (int r = 0; r < R; ++r) {
              if (act[r]) {
                const T cand = di + row[LG * r];
                if (arg[r] < 0 || cand > best[r]) { if (arg[r] >= 0) second[r] = best[r]; best[r] = cand; arg[r] = i; pval[r] = di; }
                else if (cand > second[r]) second[r] = cand;
              }
            }
          }
        }
      }
      bool tie1 = false;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int j = lane + LG * r;
        if (act[r]) {
          const T e = a.E[(b + t) * N + j];
          if (REFMODE) delta[r] = e + pval[r];
          else { delta[r] = e + best[r]; a.bp[(b + t) * N + j] = (uint16_t)arg[r]; }
          tie1 = tie1 || (multi && is_tie<T>(best[r], second[r]));
        }
      }
      tie1 = __any_sync(mask, tie1);
      tflag_cur = tie1 ? 2 : 0;
      if (sizeof(T) == 4) {  // renormalise
        T mx = neg_inf<T>();
#pragma unroll
        for (int r = 0; r < R; ++r) if (act[r] && delta[r] > mx) mx = delta[r];
#pragma unroll
        for (int o = LG / 2; o > 0; o >>= 1) { T ov = __shfl_xor_sync(mask, mx, o, LG); mx = ov > mx ? ov : mx; }
        if (mx > neg_inf<T>() && mx < -neg_inf<T>()) {
#pragma unroll
          for (int r = 0; r < R; ++r) delta[r] -= mx;
          shift += (double)mx;
        }
      }
      if (a.tie && lane == 0) {
        if (tflag_prev) a.tie[gb + t - 1] |= tflag_prev;
        if (tflag_cur) a.tie[gb + t] |= tflag_cur;
      }
    }
    bool tiel;
    const int last = group_first_argmax<T, LG, R>(delta, act, lane, mask, &tiel);
    if (a.tie && lane == 0 && tiel && multi) a.tie[gb + Tn - 1] |= 1;
    if (REFMODE) {
      double sc = 0.0;
#pragma unroll
      for (int r = 0; r < R; ++r) if (act[r]) sc += (double)delta[r] + shift;
#pragma unroll
      for (int o = LG / 2; o > 0; o >>= 1) sc += __shfl_xor_sync(mask, sc, o, LG);
      if (lane == 0) {
        a.path[gb + Tn - 1] = (uint16_t)last;
        if (a.score) a.score[s] = Tn > 1 ? sc : 0.0;
      }
    } else {
      T bv = neg_inf<T>();
#pragma unroll
      for (int r = 0; r < R; ++r) if (act[r] && lane + LG * r == last) bv = delta[r];
#pragma unroll
      for (int o = LG / 2; o > 0; o >>= 1) { T ov = __shfl_xor_sync(mask, bv, o, LG); bv = ov > bv ? ov : bv; }
      __syncwarp(mask);  // back-pointers written by the other lanes of the group
      if (lane == 0) {
        if (a.score) a.score[s] = (double)bv + shift;
        int q = last;
        for (int64_t t = Tn - 1; t >= 1; --t) {
          a.path[gb + t] = (uint16_t)q;
          q = ((volatile uint16_t *)a.bp)[(b + t) * N + q];
        }
        a.path[gb] = (uint16_t)q;
      }
    }
  }
}

template <typename T, int LG, int R>
static void launch_decode_t(dnbc_ctx *c, int mode, const DecArgs<T> &a) {
  const int gpb = 256 / LG;
  const int64_t nseq = a.s1 - a.s0;
  int64_t blocks = (nseq + gpb - 1) / gpb;
  const int64_t cap = (int64_t)c->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (mode == DNBC_DECODE_REFERENCE) k_decode<T, LG, R, true><<<(int)blocks, 256, 0, c->stream>>>(a);
  else k_decode<T, LG, R, false><<<(int)blocks, 256, 0, c->stream>>>(a);
}

template <typename T>
static void launch_decode_any(dnbc_ctx *c, int mode, const DecArgs<T> &a) {
  switch (c->dm.Npad) {
    case 1:
    case 2: launch_decode_t<T, 2, 1>(c, mode, a); break;
    case 4: launch_decode_t<T, 4, 1>(c, mode, a); break;
    case 8: launch_decode_t<T, 8, 1>(c, mode, a); break;
    case 16: launch_decode_t<T, 16, 1>(c, mode, a); break;
    case 32: launch_decode_t<T, 32, 1>(c, mode, a); break;
    case 64: launch_decode_t<T, 32, 2>(c, mode, a); break;
    case 128: launch_decode_t<T, 32, 4>(c, mode, a); break;
    default: break;  // Npad >= 256 goes to k_large_n.cu
  }
}

void launch_decode_f32(dnbc_ctx *c, int mode, int64_t s0, int64_t s1, int64_t p0, const float *E, uint16_t *path,
                       double *score, uint8_t *tie, void *bp) {
  if (s1 <= s0) return;
  FamTimer ft(c, FAM_DECODE);
  if (c->dm.Npad >= 256) return launch_decode_block_f32(c, mode, s0, s1, p0, E, path, score, tie, bp);
  DecArgs<float> a{c->dm, c->d_off, s0, s1, p0, E, c->t_logpi, c->t_logA, c->dm.Npad, c->d_active, path, score, tie,
                   (uint16_t *)bp};
  launch_decode_any<float>(c, mode, a);
}

void launch_decode_f64(dnbc_ctx *c, int mode, int64_t s0, int64_t s1, int64_t p0, const double *E, uint16_t *path,
                       double *score, uint8_t *tie, void *bp) {
  if (s1 <= s0) return;
  FamTimer ft(c, FAM_DECODE);
  if (c->dm.Npad >= 256) return launch_decode_block_f64(c, mode, s0, s1, p0, E, path, score, tie, bp);
  DecArgs<double> a{c->dm, c->d_off, s0, s1, p0, E, c->t64_logpi, c->t64_logA, c->dm.N, c->d_active, path, score, tie,
                    (uint16_t *)bp};
  launch_decode_any<double>(c, mode, a);
}
