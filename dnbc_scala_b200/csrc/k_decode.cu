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
#include "packed_f32.cuh"

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

// ---------------------------------------------------------------------------------
// fast path: fp32 BACKTRACE mode, 4..64 (padded) states -- the Viterbi throughput metric.
//
// Same recursion, tie rule and outputs as k_decode<float, .., false> above, restructured around
// what the generic kernel spends its time on (one shuffle per source state, a 4-way
// compare/select per candidate, ln a_ij from global memory, a serial back-trace per warp):
//   * column j of ln A lives in registers as packed pairs over the source index, delta_{t-1} is
//     exchanged through a warp-private shared-memory line and read back as broadcast LDS.128;
//   * pass 1 forms the N candidates of a destination with packed adds and reduces them with
//     3-input max instructions to one maximum per block of 4 sources (1.5 issue slots per
//     candidate instead of ~6), then the maximum over blocks;
//   * pass 2 finds the arg-max exactly: the bit mask of blocks within the tie tolerance of the
//     maximum has a single bit except at (flagged) ties, so the winner's block is ffs(mask) and
//     only its 4 candidates are re-formed (ln A from a shared-memory copy, bit-identical adds)
//     and compared for equality; at ties the whole warp takes the exact first-block search;
//   * the back-trace is a separate kernel, one THREAD per sequence: the pointer chase is serial
//     per sequence but 10^5 sequences chase concurrently (~1 ms instead of ~10^2 sequential walks
//     per warp).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void vadd2(u64 a, u64 b, float &lo, float &hi) { unpk(add2(a, b), lo, hi); }

constexpr int VW = 4;   // warps per CTA
void launch_backtrace(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0, const uint16_t *bp, uint16_t *path);

template <int LG, int R>
__global__ void __launch_bounds__(VW * 32) k_viterbi2(DecArgs<float> a) {
  constexpr int NP = LG * R, NB = NP / 4, PG = 32 / LG;
  __shared__ __align__(16) float sv[VW][PG][NP];
  __shared__ float sLA[NP * NP];                             // [from][to] copy of ln A for pass 2
  const int N = a.dm.N;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, lj = lane % LG, gi = lane / LG;
  const unsigned mask = dgroup_mask<LG>();
  bool act[R];
  int jr[R];
  u64 LA[R][NP / 2];
  for (int q = threadIdx.x; q < NP * NP; q += blockDim.x) {
    const int i = q / NP, j = q % NP;
    sLA[q] = (i < N && j < N && a.active[i] && a.active[j]) ? a.logA[(int64_t)i * a.lda + j] : -INFINITY;
  }
  __syncthreads();
  int na = 0;
  for (int j = 0; j < N; ++j) na += a.active[j] != 0;
  const bool multi = na >= 2;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    jr[r] = lj + LG * r;
    act[r] = jr[r] < N && a.active[jr[r]] != 0;
#pragma unroll
    for (int i = 0; i < NP; i += 2) LA[r][i / 2] = pk(sLA[i * NP + jr[r]], sLA[(i + 1) * NP + jr[r]]);
  }
  float *svme = sv[w][gi];
  const float4 *sv4 = (const float4 *)svme;
  const int64_t ngroups = (int64_t)gridDim.x * VW * PG;
  for (int64_t sw = a.s0 + ((int64_t)blockIdx.x * VW + w) * PG; sw < a.s1; sw += ngroups) {
    const int64_t s = sw + gi;
    const bool have = s < a.s1;                              // (a partly filled warp keeps all its lanes in step)
    const int64_t b = have ? a.off[s] - a.p0 : 0, gb = have ? a.off[s] : 0;
    const int Tn = have ? (int)(a.off[s + 1] - a.off[s]) : 0;
    int Tmax = Tn;
#pragma unroll
    for (int o = 16; o >= LG; o >>= 1) { const int x = __shfl_xor_sync(0xffffffffu, Tmax, o); Tmax = x > Tmax ? x : Tmax; }
    if (have && Tn <= 0 && lj == 0 && a.score) a.score[s] = 0.0;
    float delta[R], en[R];
    double shift = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      delta[r] = (act[r] && Tn > 0) ? a.E[b * N + jr[r]] + a.logpi[jr[r]] : -INFINITY;
      en[r] = (act[r] && Tn > 1) ? a.E[(b + 1) * N + jr[r]] : -INFINITY;
    }
    for (int t = 1; t < Tmax; ++t) {
      const bool live = t < Tn;
      __syncwarp();
#pragma unroll
      for (int r = 0; r < R; ++r) svme[jr[r]] = delta[r];
      __syncwarp();
      float e[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        e[r] = en[r];
        en[r] = (act[r] && t + 1 < Tn) ? a.E[(b + t + 1) * N + jr[r]] : -INFINITY;
      }
      // ---- pass 1: block maxima of delta_{t-1}(i) + ln a_ij
      float mb[R][NB];
#pragma unroll
      for (int bk = 0; bk < NB; ++bk) {
        const float4 d = sv4[bk];
        const u64 d01 = pk(d.x, d.y), d23 = pk(d.z, d.w);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float c0, c1, c2, c3;
          vadd2(d01, LA[r][2 * bk], c0, c1);
          vadd2(d23, LA[r][2 * bk + 1], c2, c3);
          mb[r][bk] = fmaxf(fmaxf(c0, c1), fmaxf(c2, c3));
        }
      }
      float best[R], thr[R];
      unsigned Q[R];
      bool odd = false;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float m = mb[r][0];
#pragma unroll
        for (int bk = 1; bk < NB; ++bk) m = fmaxf(m, mb[r][bk]);
        best[r] = m;
        thr[r] = m - DNBC_TIE_TOL_F32;
        unsigned qm = 0;
#pragma unroll
        for (int bk = 0; bk < NB; ++bk) qm |= (mb[r][bk] >= thr[r] ? 1u : 0u) << bk;
        Q[r] = qm;
        odd = odd || __popc(qm) != 1;
      }
      // ---- pass 2: the winner's block, then the winner
      int bs[R];
#pragma unroll
      for (int r = 0; r < R; ++r) bs[r] = __ffs(Q[r]) - 1;
      if (__any_sync(0xffffffffu, odd)) {                     // tie (or all -inf): exact first block holding the maximum
#pragma unroll
        for (int r = 0; r < R; ++r) {
          int f = 0;
#pragma unroll
          for (int bk = NB - 1; bk >= 0; --bk) f = mb[r][bk] == best[r] ? bk : f;
          bs[r] = f;
        }
      }
      bool tie = false;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int bq = bs[r] < 0 ? 0 : bs[r];
        const float4 d = sv4[bq];
        const float *la = sLA + (bq * 4) * NP + jr[r];
        const float c0 = d.x + la[0], c1 = d.y + la[NP], c2 = d.z + la[2 * NP], c3 = d.w + la[3 * NP];
        int k = 3;
        k = c2 == best[r] ? 2 : k;
        k = c1 == best[r] ? 1 : k;
        k = c0 == best[r] ? 0 : k;
        const int near = (c0 >= thr[r]) + (c1 >= thr[r]) + (c2 >= thr[r]) + (c3 >= thr[r]);
        const bool tr = __popc(Q[r]) != 1 || near >= 2 || !(best[r] - thr[r] > 0.f);
        const int arg = bq * 4 + k;
        if (act[r]) {
          delta[r] = e[r] + best[r];
          if (live) a.bp[(b + t) * N + jr[r]] = (uint16_t)arg;
          tie = tie || (multi && tr);
        }
      }
      tie = __any_sync(mask, tie);
      // ---- renormalise (the shift is carried in fp64)
      float mx = -INFINITY;
#pragma unroll
      for (int r = 0; r < R; ++r) mx = (act[r] && delta[r] > mx) ? delta[r] : mx;
#pragma unroll
      for (int o = LG / 2; o > 0; o >>= 1) { const float ov = __shfl_xor_sync(mask, mx, o, LG); mx = ov > mx ? ov : mx; }
      if (live && mx > -INFINITY && mx < INFINITY) {
#pragma unroll
        for (int r = 0; r < R; ++r) delta[r] -= mx;
        shift += (double)mx;
      }
      if (a.tie && lj == 0 && live && tie) a.tie[gb + t] |= 2;
      if (!live) {                                           // shorter sequence of a shared warp: freeze
#pragma unroll
        for (int r = 0; r < R; ++r) delta[r] = svme[jr[r]];
      }
    }
    bool tiel;
    const int last = group_first_argmax<float, LG, R>(delta, act, lj, mask, &tiel);
    float bv = -INFINITY;
#pragma unroll
    for (int r = 0; r < R; ++r) if (act[r] && jr[r] == last) bv = delta[r];
#pragma unroll
    for (int o = LG / 2; o > 0; o >>= 1) { const float ov = __shfl_xor_sync(mask, bv, o, LG); bv = ov > bv ? ov : bv; }
    if (have && Tn > 0 && lj == 0) {
      if (a.tie && tiel && multi) a.tie[gb + Tn - 1] |= 1;
      if (a.score) a.score[s] = (double)bv + shift;
      a.path[gb + Tn - 1] = (uint16_t)last;                   // k_backtrace starts from here
    }
  }
}

// one thread per sequence: walks the back-pointers from the last state k_viterbi2 left in path[]
__global__ void k_backtrace(const int64_t *__restrict__ off, int64_t s0, int64_t s1, int64_t p0, int N,
                            const uint16_t *__restrict__ bp, uint16_t *__restrict__ path) {
  for (int64_t s = s0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < s1; s += (int64_t)gridDim.x * blockDim.x) {
    const int64_t gb = off[s], b = gb - p0;
    const int64_t Tn = off[s + 1] - gb;
    if (Tn <= 0) continue;
    int q = path[gb + Tn - 1];
    for (int64_t t = Tn - 1; t >= 1; --t) {
      q = __ldg(bp + (b + t) * N + q);
      q = q < N ? q : N - 1;      // (only reachable on a sequence of probability zero, whose pointers are unset)
      path[gb + t - 1] = (uint16_t)q;
    }
  }
}

template <int LG, int R>
static void launch_viterbi2_t(dnbc_ctx *c, const DecArgs<float> &a) {
  const int per_block = VW * (32 / LG);
  const int64_t nseq = a.s1 - a.s0;
  int64_t blocks = (nseq + per_block - 1) / per_block;
  const int64_t cap = (int64_t)c->sm_count * (LG * R > 32 ? 2 : 4);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_viterbi2<LG, R><<<(int)blocks, VW * 32, 0, c->stream>>>(a);
  launch_backtrace(c, a.s0, a.s1, a.p0, a.bp, a.path);
}

void launch_backtrace(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0, const uint16_t *bp, uint16_t *path) {
  int64_t bb = (s1 - s0 + 127) / 128;
  if (bb > c->sm_count * 16) bb = c->sm_count * 16;
  if (bb < 1) bb = 1;
  c->launches++;
  k_backtrace<<<(int)bb, 128, 0, c->stream>>>(c->d_off, s0, s1, p0, c->dm.N, bp, path);
}

static bool viterbi2(dnbc_ctx *c, int mode, const DecArgs<float> &a) {
  if (mode != DNBC_DECODE_BACKTRACE) return false;
  switch (c->dm.Npad) {
    case 4: launch_viterbi2_t<4, 1>(c, a); return true;
    case 8: launch_viterbi2_t<8, 1>(c, a); return true;
    case 16: launch_viterbi2_t<16, 1>(c, a); return true;
    case 32: launch_viterbi2_t<32, 1>(c, a); return true;
    case 64: launch_viterbi2_t<32, 2>(c, a); return true;
    default: return false;
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
  if (mode == DNBC_DECODE_BACKTRACE && viterbi_tile_supported(c) &&
      launch_viterbi_tile(c, s0, s1, p0, E, path, score, tie, (uint16_t *)bp) == DNBC_OK)
    return;                                                  // max-plus tiles (k_viterbi_tile.cu)
  if (c->dm.Npad >= 256) return launch_decode_block_f32(c, mode, s0, s1, p0, E, path, score, tie, bp);
  DecArgs<float> a{c->dm, c->d_off, s0, s1, p0, E, c->t_logpi, c->t_logA, c->dm.Npad, c->d_active, path, score, tie,
                   (uint16_t *)bp};
  if (viterbi2(c, mode, a)) return;
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
