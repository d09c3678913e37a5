// k_supervised.cu -- the reference's supervised learner on the device, fp64.
//
// Reference code this stands in for (paths relative to the reference root):
//   mle accumulate loop              core/src/main/scala/DynamicNaiveBayesianClassifier.scala:145-159
//   learnDiscreteEmissions           ...scala:163-178  + DiscreteEdge.learn  core/src/main/scala/Edge.scala:46-51
//   learnTransitions (with its quirk) ...scala:204-216
//   learnInitialEdge                 ...scala:218-220
//   DiscreteEdge.learnFinalize       Edge.scala:53-59
//   ContinuousEdge.learnFinalize     Edge.scala:87-104
//   KMeans.run / repartitionStep / centroidStep   core/src/main/java/Tools/KMeans.java:20-37, 83-106, 116-123
//   EM1D.initialize / run / logLikelihood         core/src/main/java/Tools/ExpectationMaximization1D.java:21-59, 68-145, 154-159
//   UnivariateGaussian.density       core/src/main/java/jMEF/UnivariateGaussian.java:197
//
// The reference walks sequences one at a time on the Spark driver and then runs one
// k-means + EM per (continuous variable, hidden state) edge inside a Spark task.  Here
// every pass is a single sweep over the resident points: a point's label selects the edge
// it belongs to, all C*N edges advance together, and per-edge "done" flags reproduce the
// reference's independent per-edge stopping rules.
#include "dnbc_internal.cuh"

// ---------------------------------------------------------------------------------
// counts
// ---------------------------------------------------------------------------------
// first_seq[f] = first sequence in which state f occurs as a transition source.  With the
// learnTransitions quirk the membership test looks at the map as it was when
// the sequence started, so inside that first sequence every occurrence of f replaces its
// edge with a fresh one: only the LAST transition out of f in that sequence survives.
__global__ void k_first_seq(int N, const int64_t *__restrict__ off, int64_t S, const int32_t *__restrict__ labels,
                            int32_t *first_seq) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = off[s], e = off[s + 1];
    const int si = s > 0x7fffffff ? 0x7fffffff : (int)s;
    for (int64_t n = b; n + 1 < e; ++n) {
      const int f = labels[n];
      if (f >= 0 && f < N && first_seq[f] > si) atomicMin(first_seq + f, si);
    }
  }
}

// thread per sequence, walking backwards so "is this the last transition out of f in this
// sequence" is a bit test
__global__ void k_counts(Dims dm, const int64_t *__restrict__ off, int64_t S, int64_t P,
                         const int32_t *__restrict__ labels, const uint8_t *__restrict__ disc,
                         const int32_t *__restrict__ first_seq, int quirk, int dbl, unsigned long long *cnt_pi,
                         unsigned long long *cnt_A, unsigned long long *cnt_B) {
  const int N = dm.N, D = dm.D, Vmax = dm.Vmax;
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = off[s], e = off[s + 1];
    if (e <= b) continue;
    const unsigned long long wt = (dbl && s == 0) ? 2ull : 1ull;
    uint32_t seen[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) seen[q] = 0u;
    for (int64_t n = e - 1; n >= b; --n) {
      const int j = labels[n];
      if (j < 0 || j >= N) continue;
      for (int d = 0; d < D; ++d) {
        const int v = disc[(int64_t)d * P + n];
        if (v < Vmax) atomicAdd(cnt_B + ((int64_t)d * N + j) * Vmax + v, wt);
      }
      if (n + 1 < e) {
        const int to = labels[n + 1];
        if (to >= 0 && to < N) {
          unsigned long long add;
          if (quirk && first_seq[j] == (int)s) {
            const bool last = !((seen[j >> 5] >> (j & 31)) & 1u);
            add = (last ? 1ull : 0ull) + (wt - 1ull);  // second pass over sequence 0 counts normally
          } else {
            add = wt;
          }
          if (add) atomicAdd(cnt_A + (int64_t)j * N + to, add);
        }
        seen[j >> 5] |= 1u << (j & 31);
      }
    }
    const int j0 = labels[b];
    if (j0 >= 0 && j0 < N) atomicAdd(cnt_pi + j0, wt);
  }
}

void launch_supervised_counts(dnbc_ctx *c, int flags, unsigned long long *cnt_pi, unsigned long long *cnt_A,
                              unsigned long long *cnt_B) {
  const Dims &d = c->dm;
  const size_t total = (size_t)d.N + (size_t)d.N * d.N + (size_t)d.D * d.N * d.Vmax;
  cudaMemsetAsync(cnt_pi, 0, total * sizeof(unsigned long long), c->stream);
  cudaMemsetAsync(c->d_first_seq, 0x7f, d.N * sizeof(int32_t), c->stream);
  if (c->S == 0) return;
  const int quirk = (flags & DNBC_QUIRK_TRANSITIONS) != 0, dbl = (flags & DNBC_DOUBLE_COUNT_FIRST) != 0;
  int grid = (int)std::min<int64_t>((c->S + 127) / 128, (int64_t)c->sm_count * 16);
  c->launches += 2;
  k_first_seq<<<grid, 128, 0, c->stream>>>(d.N, c->d_off, c->S, c->d_labels, c->d_first_seq);
  k_counts<<<grid, 128, 0, c->stream>>>(d, c->d_off, c->S, c->P, c->d_labels, c->d_disc, c->d_first_seq, quirk, dbl,
                                        cnt_pi, cnt_A, cnt_B);
}

// DiscreteEdge.learnFinalize (Edge.scala:53-59): P = count / sum.  Rows with no counts stay
// zero; the decode state set is `transitions.keys` = states with a non-empty A row.
__global__ void k_normalize_counts(Dims dm, const unsigned long long *cnt_pi, const unsigned long long *cnt_A,
                                   const unsigned long long *cnt_B, double *pi, double *A, double *B, uint8_t *active) {
  const int N = dm.N, D = dm.D, Vmax = dm.Vmax;
  const int rows = 1 + N + D * N;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
    const unsigned long long *src;
    double *dst;
    int cols;
    if (r == 0) { src = cnt_pi; dst = pi; cols = N; }
    else if (r <= N) { src = cnt_A + (int64_t)(r - 1) * N; dst = A + (int64_t)(r - 1) * N; cols = N; }
    else { src = cnt_B + (int64_t)(r - 1 - N) * Vmax; dst = B + (int64_t)(r - 1 - N) * Vmax; cols = Vmax; }
    unsigned long long sum = 0;
    for (int q = 0; q < cols; ++q) sum += src[q];
    for (int q = 0; q < cols; ++q) dst[q] = sum ? (double)src[q] / (double)sum : 0.0;
    if (r >= 1 && r <= N) active[r - 1] = sum != 0;
  }
}

void launch_normalize_counts(dnbc_ctx *c, const unsigned long long *cnt_pi, const unsigned long long *cnt_A,
                             const unsigned long long *cnt_B) {
  const int rows = 1 + c->dm.N + c->dm.D * c->dm.N;
  c->launches++;
  k_normalize_counts<<<(rows + 127) / 128, 128, 0, c->stream>>>(c->dm, cnt_pi, cnt_A, cnt_B, c->d_pi, c->d_A, c->d_B,
                                                                c->d_active);
}

// ---------------------------------------------------------------------------------
// per-edge GMM learner
// ---------------------------------------------------------------------------------
// fp64 workspace, E = C*N edges, EK = E*Kmax:
struct EdgeWs {
  double *cen, *ksum, *kcnt;   // [EK] k-means centroids, running sum / weight
  double *fsum, *fcnt;         // [EK] sum / weight of the final clusters
  double *vsum;                // [EK] sum (x - mean)^2 of the final clusters
  double *s0, *s1, *s2;        // [EK] EM moments centred on the old mean
  double *npts;                // [E]  points on the edge (weighted)
  double *ll_new, *ll_old, *thr;  // [E]
  double *xmin, *xmax;         // [E]
  int32_t *changed, *kdone, *kit, *done, *it;  // [E]
  int32_t *pending;            // [1] edges still running
};
static size_t edge_ws_doubles(int64_t E, int64_t EK) { return 9 * EK + 6 * E; }
static EdgeWs edge_ws(dnbc_ctx *c) {
  const int64_t E = (int64_t)c->dm.C * c->dm.N, EK = E * c->dm.Kmax;
  EdgeWs w;
  double *p = c->d_edge;
  w.cen = p; p += EK; w.ksum = p; p += EK; w.kcnt = p; p += EK; w.fsum = p; p += EK; w.fcnt = p; p += EK;
  w.vsum = p; p += EK; w.s0 = p; p += EK; w.s1 = p; p += EK; w.s2 = p; p += EK;
  w.npts = p; p += E; w.ll_new = p; p += E; w.ll_old = p; p += E; w.thr = p; p += E; w.xmin = p; p += E; w.xmax = p;
  int32_t *q = c->d_edge_i;
  w.changed = q; q += E; w.kdone = q; q += E; w.kit = q; q += E; w.done = q; q += E; w.it = q; q += E; w.pending = q;
  return w;
}

struct PointArgs {
  Dims dm;
  int64_t P, first_end;   // first_end = off[1] when the first sequence counts twice, else 0
  const int32_t *labels;
  const float *cont;
  const double *cont64;
  const int32_t *K;
};
__device__ __forceinline__ double point_x(const PointArgs &a, int c, int64_t n) {
  return a.cont64 ? a.cont64[(int64_t)c * a.P + n] : (double)a.cont[(int64_t)c * a.P + n];
}

__device__ __forceinline__ double atomic_min_f64(double *addr, double v) {
  unsigned long long *p = (unsigned long long *)addr, old = *p, assumed;
  do {
    assumed = old;
    if (!(v < __longlong_as_double(assumed))) break;
    old = atomicCAS(p, assumed, __double_as_longlong(v));
  } while (assumed != old);
  return __longlong_as_double(old);
}
__device__ __forceinline__ double atomic_max_f64(double *addr, double v) {
  unsigned long long *p = (unsigned long long *)addr, old = *p, assumed;
  do {
    assumed = old;
    if (!(v > __longlong_as_double(assumed))) break;
    old = atomicCAS(p, assumed, __double_as_longlong(v));
  } while (assumed != old);
  return __longlong_as_double(old);
}

#define POINT_LOOP(a)                                                                                   \
  for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < (a).P; n += (int64_t)gridDim.x * blockDim.x)

// weighted point count and range per edge
__global__ void k_edge_range(PointArgs a, EdgeWs w) {
  const int N = a.dm.N, C = a.dm.C;
  POINT_LOOP(a) {
    const int j = a.labels[n];
    if (j < 0 || j >= N) continue;
    const double wt = n < a.first_end ? 2.0 : 1.0;
    for (int c = 0; c < C; ++c) {
      const int e = c * N + j;
      const double x = point_x(a, c, n);
      atomicAdd(w.npts + e, wt);
      atomic_min_f64(w.xmin + e, x);
      atomic_max_f64(w.xmax + e, x);
    }
  }
}

// default seeding when the caller injects no centroids: K evenly spaced values across the
// edge's range (the reference draws K distinct data points from an UNSEEDED java.util.Random,
// KMeans.java:46-72, so there is no reference sequence to reproduce)
__global__ void k_seed_centroids(Dims dm, const int32_t *K, EdgeWs w) {
  const int E = dm.C * dm.N;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    const int k = K[e / dm.N];
    for (int q = 0; q < k; ++q)
      w.cen[(int64_t)e * dm.Kmax + q] = w.xmin[e] + (w.xmax[e] - w.xmin[e]) * ((q + 0.5) / k);
  }
}

// KMeans.repartitionStep :83-106 -- nearest centroid by sqrt((x-c)^2), strict `<` (lowest
// index wins, NaN centroids never win), plus the running sums centroidStep needs.
__global__ void k_km_assign(PointArgs a, EdgeWs w, uint8_t *assign) {
  const int N = a.dm.N, C = a.dm.C, Kmax = a.dm.Kmax;
  POINT_LOOP(a) {
    const int j = a.labels[n];
    if (j < 0 || j >= N) continue;
    const double wt = n < a.first_end ? 2.0 : 1.0;
    for (int c = 0; c < C; ++c) {
      const int e = c * N + j;
      if (w.kdone[e]) continue;
      const double x = point_x(a, c, n);
      const int k = a.K[c];
      int index = 0;
      double dist = 1.7976931348623157e308;
      for (int q = 0; q < k; ++q) {
        const double diff = x - w.cen[(int64_t)e * Kmax + q];
        const double dt = sqrt(diff * diff);
        if (dt < dist) { dist = dt; index = q; }
      }
      uint8_t *slot = assign + (int64_t)c * a.P + n;
      if (*slot != (uint8_t)index) { w.changed[e] = 1; *slot = (uint8_t)index; }
      atomicAdd(w.ksum + (int64_t)e * Kmax + index, wt * x);
      atomicAdd(w.kcnt + (int64_t)e * Kmax + index, wt);
    }
  }
}

// KMeans.centroidStep :116-123 and the loop test of KMeans.run :28-34
__global__ void k_km_update(Dims dm, const int32_t *K, EdgeWs w, int max_it) {
  const int E = dm.C * dm.N, Kmax = dm.Kmax;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    if (w.kdone[e]) continue;
    const int k = K[e / dm.N];
    for (int q = 0; q < k; ++q) {
      const int64_t o = (int64_t)e * Kmax + q;
      w.cen[o] = w.ksum[o] * (1.0 / w.kcnt[o]);  // empty cluster: 0 * inf = NaN, like the reference
      w.fsum[o] = w.ksum[o];
      w.fcnt[o] = w.kcnt[o];
      w.ksum[o] = 0.0;
      w.kcnt[o] = 0.0;
    }
    w.kit[e] += 1;
    if (!w.changed[e] || w.kit[e] >= max_it || w.npts[e] == 0.0) w.kdone[e] = 1;
    else atomicAdd(w.pending, 1);
    w.changed[e] = 0;
  }
}

// EM1D.initialize :21-59 -- variance pass around the cluster mean
__global__ void k_em_init_var(PointArgs a, EdgeWs w, const uint8_t *assign) {
  const int N = a.dm.N, C = a.dm.C, Kmax = a.dm.Kmax;
  POINT_LOOP(a) {
    const int j = a.labels[n];
    if (j < 0 || j >= N) continue;
    const double wt = n < a.first_end ? 2.0 : 1.0;
    for (int c = 0; c < C; ++c) {
      const int64_t o = (int64_t)(c * N + j) * Kmax + assign[(int64_t)c * a.P + n];
      const double mean = w.fsum[o] / w.fcnt[o];
      const double diff = point_x(a, c, n) - mean;
      atomicAdd(w.vsum + o, wt * diff * diff);
    }
  }
}

__global__ void k_em_init_params(Dims dm, const int32_t *K, EdgeWs w, double *pw, double *pmu, double *pvar) {
  const int E = dm.C * dm.N, Kmax = dm.Kmax;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    const int k = K[e / dm.N];
    for (int q = 0; q < Kmax; ++q) {
      const int64_t o = (int64_t)e * Kmax + q;
      if (q < k && w.npts[e] > 0.0) {
        pw[o] = w.fcnt[o] / w.npts[e];
        pmu[o] = w.fsum[o] / w.fcnt[o];
        pvar[o] = w.vsum[o] / w.fcnt[o];
      } else {
        pw[o] = 0.0; pmu[o] = 0.0; pvar[o] = 1.0;
      }
    }
    w.done[e] = w.npts[e] > 0.0 ? 0 : 1;
    w.it[e] = 0;
  }
}

// jMEF/UnivariateGaussian.java:197
__device__ __forceinline__ double jmef_density(double x, double mu, double var) {
  return exp(-(x - mu) * (x - mu) / (2.0 * var)) / sqrt(2.0 * DNBC_PI * var);
}

// EM1D.logLikelihood :154-159
__global__ void k_em_loglik(PointArgs a, EdgeWs w, const double *pw, const double *pmu, const double *pvar) {
  const int N = a.dm.N, C = a.dm.C, Kmax = a.dm.Kmax;
  POINT_LOOP(a) {
    const int j = a.labels[n];
    if (j < 0 || j >= N) continue;
    const double wt = n < a.first_end ? 2.0 : 1.0;
    for (int c = 0; c < C; ++c) {
      const int e = c * N + j;
      if (w.done[e]) continue;
      const double x = point_x(a, c, n);
      double dens = 0.0;
      for (int q = 0; q < a.K[c]; ++q) {
        const int64_t o = (int64_t)e * Kmax + q;
        dens += pw[o] * jmef_density(x, pmu[o], pvar[o]);
      }
      atomicAdd(w.ll_new + e, wt * log(dens));
    }
  }
}

// EM1D.run E-step :92-101 and the sums of the M-step :104-132, moments centred on the old mean
__global__ void k_em_moments(PointArgs a, EdgeWs w, const double *pw, const double *pmu, const double *pvar) {
  const int N = a.dm.N, C = a.dm.C, Kmax = a.dm.Kmax;
  POINT_LOOP(a) {
    const int j = a.labels[n];
    if (j < 0 || j >= N) continue;
    const double wt = n < a.first_end ? 2.0 : 1.0;
    for (int c = 0; c < C; ++c) {
      const int e = c * N + j;
      if (w.done[e]) continue;
      const double x = point_x(a, c, n);
      const int k = a.K[c];
      double p[32], sum = 0.0;
      for (int q = 0; q < k; ++q) {
        const int64_t o = (int64_t)e * Kmax + q;
        p[q] = pw[o] * jmef_density(x, pmu[o], pvar[o]);
        sum += p[q];
      }
      for (int q = 0; q < k; ++q) {
        const int64_t o = (int64_t)e * Kmax + q;
        const double r = wt * (p[q] / sum), dx = x - pmu[o];
        atomicAdd(w.s0 + o, r);
        atomicAdd(w.s1 + o, r * dx);
        atomicAdd(w.s2 + o, r * dx * dx);
      }
    }
  }
}

// M-step update :126-131 (mu = sum x p / sum p; sigma^2 around the NEW mean; w = sum p / n)
__global__ void k_em_update(Dims dm, const int32_t *K, EdgeWs w, double *pw, double *pmu, double *pvar) {
  const int E = dm.C * dm.N, Kmax = dm.Kmax;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    if (w.done[e]) continue;
    const int k = K[e / dm.N];
    for (int q = 0; q < k; ++q) {
      const int64_t o = (int64_t)e * Kmax + q;
      const double sh = w.s1[o] / w.s0[o];
      pmu[o] = pmu[o] + sh;
      pvar[o] = w.s2[o] / w.s0[o] - sh * sh;
      pw[o] = w.s0[o] / w.npts[e];
      w.s0[o] = w.s1[o] = w.s2[o] = 0.0;
    }
    w.it[e] += 1;
    w.ll_old[e] = w.ll_new[e];
    w.ll_new[e] = 0.0;
  }
}

// after a log-likelihood pass: record the series and apply the stop rule :81, :141
__global__ void k_em_check(Dims dm, EdgeWs w, int max_it, double rel_tol, double *ll_series, int first) {
  const int E = dm.C * dm.N;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    if (w.done[e]) continue;
    if (ll_series) ll_series[(int64_t)e * (max_it + 1) + w.it[e]] = w.ll_new[e];
    if (first) {
      w.thr[e] = fabs(w.ll_new[e]) * rel_tol;
      atomicAdd(w.pending, 1);
    } else if (fabs(w.ll_new[e] - w.ll_old[e]) > w.thr[e] && w.it[e] < max_it) {
      atomicAdd(w.pending, 1);
    } else {
      w.done[e] = 1;
    }
  }
}

static int read_pending(dnbc_ctx *c, EdgeWs &w, int *out) {
  CU_TRY(c, cudaMemcpyAsync(out, w.pending, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  CU_TRY(c, cudaMemsetAsync(w.pending, 0, sizeof(int), c->stream));
  return DNBC_OK;
}

int run_gmm_mle(dnbc_ctx *c, int flags, const double *h_init_centroids, int max_iters, double rel_tol,
                double *h_ll_series, int32_t *h_n_iters) {
  const Dims &d = c->dm;
  const int64_t E = (int64_t)d.C * d.N, EK = E * d.Kmax;
  cudaStream_t st = c->stream;
  if (!c->d_edge) {
    CU_TRY(c, cudaMalloc((void **)&c->d_edge, edge_ws_doubles(E, EK) * sizeof(double)));
    CU_TRY(c, cudaMalloc((void **)&c->d_edge_i, (5 * E + 1) * sizeof(int32_t)));
  }
  if (c->assign_cap < (int64_t)d.C * c->P || !c->d_assign) {
    if (c->d_assign) cudaFree(c->d_assign);
    c->d_assign = nullptr;
    CU_TRY(c, cudaMalloc((void **)&c->d_assign, std::max<size_t>((size_t)d.C * c->P, 1)));
    c->assign_cap = (int64_t)d.C * c->P;
  }
  EdgeWs w = edge_ws(c);
  CU_TRY(c, cudaMemsetAsync(c->d_edge, 0, edge_ws_doubles(E, EK) * sizeof(double), st));
  CU_TRY(c, cudaMemsetAsync(c->d_edge_i, 0, (5 * E + 1) * sizeof(int32_t), st));
  CU_TRY(c, cudaMemsetAsync(c->d_assign, 0, std::max<size_t>((size_t)d.C * c->P, 1), st));
  {  // xmin = +inf, xmax = -inf
    std::vector<double> init(2 * E);
    for (int64_t e = 0; e < E; ++e) { init[e] = INFINITY; init[E + e] = -INFINITY; }
    CU_TRY(c, cudaMemcpyAsync(w.xmin, init.data(), 2 * E * sizeof(double), cudaMemcpyHostToDevice, st));
    CU_TRY(c, cudaStreamSynchronize(st));
  }
  double *d_series = nullptr;
  if (h_ll_series) {
    CU_TRY(c, cudaMalloc((void **)&d_series, (size_t)E * (max_iters + 1) * sizeof(double)));
    CU_TRY(c, cudaMemsetAsync(d_series, 0, (size_t)E * (max_iters + 1) * sizeof(double), st));
  }
  PointArgs a{d, c->P, ((flags & DNBC_DOUBLE_COUNT_FIRST) && c->S > 0) ? c->h_off[1] : 0, c->d_labels, c->d_cont,
              c->have_cont64 ? c->d_cont64 : nullptr, c->d_K};
  const int pgrid = (int)std::min<int64_t>((c->P + 255) / 256 + 1, (int64_t)c->sm_count * 8);
  const int egrid = (int)((E + 127) / 128);
  int rc = DNBC_OK, pending = 0;
#define STEP(expr) do { expr; c->launches++; } while (0)
  STEP((k_edge_range<<<pgrid, 256, 0, st>>>(a, w)));
  if (h_init_centroids) CU_TRY(c, cudaMemcpyAsync(w.cen, h_init_centroids, EK * sizeof(double), cudaMemcpyHostToDevice, st));
  else STEP((k_seed_centroids<<<egrid, 128, 0, st>>>(d, c->d_K, w)));
  // k-means: at most 30 rounds (KMeans.java:11)
  for (int it = 0; it < 30; ++it) {
    STEP((k_km_assign<<<pgrid, 256, 0, st>>>(a, w, c->d_assign)));
    STEP((k_km_update<<<egrid, 128, 0, st>>>(d, c->d_K, w, 30)));
    if ((rc = read_pending(c, w, &pending))) goto out;
    if (!pending) break;
  }
  STEP((k_em_init_var<<<pgrid, 256, 0, st>>>(a, w, c->d_assign)));
  STEP((k_em_init_params<<<egrid, 128, 0, st>>>(d, c->d_K, w, c->d_w, c->d_mu, c->d_var)));
  STEP((k_em_loglik<<<pgrid, 256, 0, st>>>(a, w, c->d_w, c->d_mu, c->d_var)));
  STEP((k_em_check<<<egrid, 128, 0, st>>>(d, w, max_iters, rel_tol, d_series, 1)));
  if ((rc = read_pending(c, w, &pending))) goto out;
  while (pending) {
    STEP((k_em_moments<<<pgrid, 256, 0, st>>>(a, w, c->d_w, c->d_mu, c->d_var)));
    STEP((k_em_update<<<egrid, 128, 0, st>>>(d, c->d_K, w, c->d_w, c->d_mu, c->d_var)));
    STEP((k_em_loglik<<<pgrid, 256, 0, st>>>(a, w, c->d_w, c->d_mu, c->d_var)));
    STEP((k_em_check<<<egrid, 128, 0, st>>>(d, w, max_iters, rel_tol, d_series, 0)));
    if ((rc = read_pending(c, w, &pending))) goto out;
  }
#undef STEP
  if (h_ll_series)
    cudaMemcpyAsync(h_ll_series, d_series, (size_t)E * (max_iters + 1) * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (h_n_iters) cudaMemcpyAsync(h_n_iters, w.it, E * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
out:
  if (d_series) cudaFree(d_series);
  if (rc == DNBC_OK && cudaGetLastError() != cudaSuccess) rc = dnbc_fail(c, DNBC_ECUDA, "gmm mle: kernel failure");
  return rc;
}

// ---------------------------------------------------------------------------------
// precision conversions for the observation planes
// ---------------------------------------------------------------------------------
__global__ void k_f64_to_f32(const double *__restrict__ src, float *__restrict__ dst, int64_t n) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
    dst[q] = (float)src[q];
}
__global__ void k_f32_to_f64(const float *__restrict__ src, double *__restrict__ dst, int64_t n) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
    dst[q] = (double)src[q];
}
void launch_f64_to_f32(dnbc_ctx *c, const double *src, float *dst, int64_t n) {
  if (n <= 0) return;
  c->launches++;
  k_f64_to_f32<<<(int)std::min<int64_t>((n + 255) / 256, (int64_t)c->sm_count * 16), 256, 0, c->stream>>>(src, dst, n);
}
void launch_f32_to_f64(dnbc_ctx *c, const float *src, double *dst, int64_t n) {
  if (n <= 0) return;
  c->launches++;
  k_f32_to_f64<<<(int)std::min<int64_t>((n + 255) / 256, (int64_t)c->sm_count * 16), 256, 0, c->stream>>>(src, dst, n);
}
