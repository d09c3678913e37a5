// k_large_n.cu -- forward / backward / decode recursions for large state spaces
// (Npad >= 256): one CTA per sequence, one thread per hidden state, the previous step's
// vector in shared memory.  The transition matrix (1 MB fp32 at N = 512) is streamed from
// L2 with unit stride across the CTA.  Same arithmetic and outputs as the lane-group
// kernels in k_forward_backward.cu / k_decode.cu (see those files for the reference
// citations: DynamicNaiveBayesianClassifier.scala:49-91 for decode; SURVEY Appendix B for
// forward-backward).
#include "dnbc_internal.cuh"

#define DNBC_TIE_TOL_F32 1e-3f

__device__ __forceinline__ float block_sum(float v, float *red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int q = 0; q < nw; ++q) t += red[q];
  return t;
}
__device__ __forceinline__ float block_max(float v, float *red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float t = -INFINITY;
  for (int q = 0; q < nw; ++q) t = fmaxf(t, red[q]);
  return t;
}

struct FbBlockArgs {
  Dims dm;
  DeviceTables tb;
  const int64_t *off;
  int64_t s0, s1, p0;
  const float *E;
  float *AL, *GA, *U, *cs;
  double *stats;
  int64_t st_pi;
};

__global__ void __launch_bounds__(1024) k_forward_block(FbBlockArgs a) {
  extern __shared__ float sm[];
  const int N = a.dm.N, Npad = a.dm.Npad;
  float *vec = sm, *red = sm + Npad;
  const int j = threadIdx.x;
  const bool on = j < N;
  double ll = 0.0;
  for (int64_t s = a.s0 + blockIdx.x; s < a.s1; s += gridDim.x) {
    const int64_t b = a.off[s] - a.p0, T = a.off[s + 1] - a.off[s];
    for (int64_t t = 0; t < T; ++t) {
      const float e = on ? a.E[(b + t) * N + j] : -INFINITY;
      float m = block_max(e, red);
      if (m == -INFINITY) m = 0.f;
      float acc = 0.f;
      if (t == 0) {
        acc = on ? a.tb.pi[j] : 0.f;
      } else if (on) {
        const float *col = a.tb.A + j;
#pragma unroll 8
        for (int i = 0; i < N; ++i) acc = fmaf(vec[i], col[(int64_t)i * Npad], acc);
      }
      acc *= ex2_approx((e - m) * (float)DNBC_LOG2E);
      const float csum = block_sum(acc, red);
      const float al = acc * (1.0f / csum);
      __syncthreads();
      vec[j] = al;
      if (on) a.AL[(b + t) * N + j] = al;
      if (j == 0) {
        a.cs[b + t] = csum;
        ll += (double)logf(csum) + (double)m;
      }
      __syncthreads();
    }
  }
  if (j == 0 && ll != 0.0) atomicAdd(a.stats, ll);
}

__global__ void __launch_bounds__(1024) k_backward_block(FbBlockArgs a) {
  extern __shared__ float sm[];
  const int N = a.dm.N, Npad = a.dm.Npad;
  float *vec = sm, *red = sm + Npad;
  const int j = threadIdx.x;
  const bool on = j < N;
  double piacc = 0.0;
  for (int64_t s = a.s0 + blockIdx.x; s < a.s1; s += gridDim.x) {
    const int64_t b = a.off[s] - a.p0, T = a.off[s + 1] - a.off[s];
    float beta = 1.f;
    for (int64_t t = T - 1; t >= 0; --t) {
      if (t < T - 1) {
        const float e = on ? a.E[(b + t + 1) * N + j] : -INFINITY;
        float m = block_max(e, red);
        if (m == -INFINITY) m = 0.f;
        const float u = ex2_approx((e - m) * (float)DNBC_LOG2E) * beta * (1.0f / a.cs[b + t + 1]);
        __syncthreads();
        vec[j] = on ? u : 0.f;
        if (on) a.U[(b + t) * N + j] = u;
        __syncthreads();
        float acc = 0.f;
        if (on) {
          const float *col = a.tb.At + j;  // At[jj][i]: row = to-state
#pragma unroll 8
          for (int jj = 0; jj < N; ++jj) acc = fmaf(vec[jj], col[(int64_t)jj * Npad], acc);
        }
        beta = acc;
      } else if (on) {
        a.U[(b + t) * N + j] = 0.f;
      }
      if (on) {
        const float g = a.AL[(b + t) * N + j] * beta;
        a.GA[(b + t) * N + j] = g;
        if (t == 0) piacc += (double)g;
      }
    }
  }
  if (on && piacc != 0.0) atomicAdd(a.stats + a.st_pi + j, piacc);
}

void launch_fb_block(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0, bool backward) {
  StatsLayout sl = stats_layout(c->dm);
  FbBlockArgs a{c->dm, c->tables(), c->d_off, s0, s1, p0, c->E, c->AL, c->GA, c->U, c->cs, c->d_stats, sl.pi};
  const int threads = c->dm.Npad;
  const size_t smem = (size_t)(c->dm.Npad + 32) * sizeof(float);
  const int per_sm = 2048 / threads;
  int grid = (int)std::min<int64_t>(s1 - s0, (int64_t)c->sm_count * per_sm);
  if (backward) k_backward_block<<<grid, threads, smem, c->stream>>>(a);
  else k_forward_block<<<grid, threads, smem, c->stream>>>(a);
}

// ---------------------------------------------------------------------------------
// decode
// ---------------------------------------------------------------------------------
template <typename T>
struct DecBlockArgs {
  Dims dm;
  const int64_t *off;
  int64_t s0, s1, p0;
  const T *E, *logpi, *logA;
  int lda;
  const uint8_t *active;
  uint16_t *path;
  double *score;
  uint8_t *tie;
  uint16_t *bp;
};

template <typename T> __device__ __forceinline__ T ninf();
template <> __device__ __forceinline__ float ninf<float>() { return -INFINITY; }
template <> __device__ __forceinline__ double ninf<double>() { return -(double)INFINITY; }
template <typename T>
__device__ __forceinline__ bool tie_of(T best, T second) {
  if (sizeof(T) == 8) return second == best;
  return !(best - second > (T)DNBC_TIE_TOL_F32);
}

// first maximum over the CTA's active states (lowest index wins ties, like Scala's maxBy
// over states in dense-index order); *second = runner-up value
template <typename T>
__device__ __forceinline__ int block_first_argmax(T v, bool valid, T *rv, T *rs, int *ri, T *best_out, T *second_out) {
  const int none = 1 << 30;
  T best = valid ? v : ninf<T>(), second = ninf<T>();
  int idx = valid ? (int)threadIdx.x : none;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T ob = __shfl_xor_sync(0xffffffffu, best, o);
    const T os = __shfl_xor_sync(0xffffffffu, second, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (oi != none && (idx == none || ob > best || (ob == best && oi < idx))) {
      T ns = os;
      if (idx != none && best > ns) ns = best;
      if (second > ns) ns = second;
      second = ns; best = ob; idx = oi;
    } else if (oi != none) {
      T ns = second;
      if (ob > ns) ns = ob;
      if (os > ns) ns = os;
      second = ns;
    }
  }
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { rv[w] = best; rs[w] = second; ri[w] = idx; }
  __syncthreads();
  best = ninf<T>(); second = ninf<T>(); idx = none;
  for (int q = 0; q < nw; ++q) {
    const T ob = rv[q], os = rs[q];
    const int oi = ri[q];
    if (oi == none) continue;
    if (idx == none || ob > best) {
      T ns = os;
      if (idx != none && best > ns) ns = best;
      if (second > ns) ns = second;
      second = ns; best = ob; idx = oi;
    } else {
      if (ob > second) second = ob;
      if (os > second) second = os;
    }
  }
  *best_out = best;
  *second_out = second;
  return idx == none ? 0 : idx;
}

template <typename T, bool REFMODE>
__global__ void __launch_bounds__(1024) k_decode_block(DecBlockArgs<T> a) {
  extern __shared__ unsigned char smraw[];
  const int N = a.dm.N, Npad = a.dm.Npad;
  T *vec = (T *)smraw;
  T *rv = vec + Npad, *rs = rv + 32;
  int *ri = (int *)(rs + 32);
  __shared__ int s_na, s_any;
  const int j = threadIdx.x;
  const bool act = j < N && a.active[j] != 0;
  if (j == 0) s_na = 0;
  __syncthreads();
  if (act) atomicAdd(&s_na, 1);
  __syncthreads();
  const bool multi = s_na >= 2;
  for (int64_t s = a.s0 + blockIdx.x; s < a.s1; s += gridDim.x) {
    const int64_t b = a.off[s] - a.p0, gb = a.off[s], Tn = a.off[s + 1] - a.off[s];
    if (Tn <= 0) { if (j == 0 && a.score) a.score[s] = 0.0; continue; }
    T delta = act ? a.E[b * N + j] + a.logpi[j] : ninf<T>();
    double shift = 0.0;
    T bst, snd;
    for (int64_t t = 1; t < Tn; ++t) {
      if (REFMODE) {
        const int pj = block_first_argmax<T>(delta, act, rv, rs, ri, &bst, &snd);
        if (j == 0) {
          a.path[gb + t - 1] = (uint16_t)pj;
          if (a.tie && multi && tie_of<T>(bst, snd)) a.tie[gb + t - 1] |= 1;
        }
      }
      __syncthreads();
      vec[j] = delta;
      if (j == 0) s_any = 0;
      __syncthreads();
      T best = ninf<T>(), second = ninf<T>(), pval = ninf<T>();
      int arg = -1;
      if (act) {
        const T *col = a.logA + j;
        for (int i = 0; i < N; ++i) {
          if (!a.active[i]) continue;
          const T di = vec[i];
          const T cand = di + col[(int64_t)i * a.lda];
          if (arg < 0 || cand > best) { if (arg >= 0) second = best; best = cand; arg = i; pval = di; }
          else if (cand > second) second = cand;
        }
        const T e = a.E[(b + t) * N + j];
        if (REFMODE) delta = e + pval;
        else { delta = e + best; a.bp[(b + t) * N + j] = (uint16_t)arg; }
        if (multi && tie_of<T>(best, second)) s_any = 1;
      }
      __syncthreads();
      if (sizeof(T) == 4) {  // renormalise; the shift is carried in fp64
        T mx, dummy;
        block_first_argmax<T>(delta, act, rv, rs, ri, &mx, &dummy);
        if (mx > ninf<T>() && mx < -ninf<T>()) { delta -= mx; shift += (double)mx; }
      }
      if (j == 0 && a.tie && s_any) a.tie[gb + t] |= 2;
    }
    const int last = block_first_argmax<T>(delta, act, rv, rs, ri, &bst, &snd);
    if (j == 0 && a.tie && multi && tie_of<T>(bst, snd)) a.tie[gb + Tn - 1] |= 1;
    if (REFMODE) {
      // score = sum_j delta_{T-1}(j) over the active states, in index order
      __syncthreads();
      vec[j] = delta;
      __syncthreads();
      if (j == 0) {
        double sc = 0.0;
        for (int q = 0; q < N; ++q)
          if (a.active[q]) sc += (double)vec[q] + shift;
        a.path[gb + Tn - 1] = (uint16_t)last;
        if (a.score) a.score[s] = Tn > 1 ? sc : 0.0;
      }
    } else {
      __syncthreads();  // back-pointers of this sequence are visible to thread 0
      if (j == 0) {
        if (a.score) a.score[s] = (double)bst + shift;
        int q = last;
        for (int64_t t = Tn - 1; t >= 1; --t) {
          a.path[gb + t] = (uint16_t)q;
          q = ((volatile uint16_t *)a.bp)[(b + t) * N + q];
        }
        a.path[gb] = (uint16_t)q;
      }
    }
    __syncthreads();
  }
}

template <typename T>
static void launch_decode_block_t(dnbc_ctx *c, int mode, const DecBlockArgs<T> &a) {
  const int threads = c->dm.Npad;
  const size_t smem = (size_t)(c->dm.Npad + 64) * sizeof(T) + 32 * sizeof(int);
  const int per_sm = 2048 / threads;
  const int grid = (int)std::min<int64_t>(a.s1 - a.s0, (int64_t)c->sm_count * per_sm);
  if (mode == DNBC_DECODE_REFERENCE) k_decode_block<T, true><<<grid, threads, smem, c->stream>>>(a);
  else k_decode_block<T, false><<<grid, threads, smem, c->stream>>>(a);
}

void launch_decode_block_f32(dnbc_ctx *c, int mode, int64_t s0, int64_t s1, int64_t p0, const float *E, uint16_t *path,
                             double *score, uint8_t *tie, void *bp) {
  DecBlockArgs<float> a{c->dm, c->d_off, s0, s1, p0, E, c->t_logpi, c->t_logA, c->dm.Npad, c->d_active, path, score, tie,
                        (uint16_t *)bp};
  launch_decode_block_t<float>(c, mode, a);
}
void launch_decode_block_f64(dnbc_ctx *c, int mode, int64_t s0, int64_t s1, int64_t p0, const double *E, uint16_t *path,
                             double *score, uint8_t *tie, void *bp) {
  DecBlockArgs<double> a{c->dm, c->d_off, s0, s1, p0, E, c->t64_logpi, c->t64_logA, c->dm.N, c->d_active, path, score,
                         tie, (uint16_t *)bp};
  launch_decode_block_t<double>(c, mode, a);
}
