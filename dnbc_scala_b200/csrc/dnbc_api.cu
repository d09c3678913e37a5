// dnbc_api.cu -- the C ABI of include/dnbc_b200.h: context, parameter / observation
// residency, and the orchestration of the kernels (chunking, streams, timing).
//
// The reference has no native boundary; each entry point names the reference code it
// stands in for in include/dnbc_b200.h.  Nothing here computes on the CPU except the
// natural logs of the (tiny) fp64 decode tables, which are taken with the host libm so
// the F64 decode path rounds like a JVM StrictMath/glibc oracle.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "dnbc_internal.cuh"

// ---------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------
static thread_local std::string g_create_error;

int dnbc_fail(dnbc_ctx *c, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  else g_create_error = buf;
  return code;
}

extern "C" const char *dnbc_last_error(const dnbc_ctx *ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" int dnbc_abi_version(void) { return DNBC_B200_ABI_VERSION; }

template <typename T>
static cudaError_t dev_alloc(T **p, size_t n) {
  return cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T));
}

// grow-only device buffer
template <typename T>
static cudaError_t ensure(T **p, int64_t *cap, int64_t need) {
  if (need <= *cap && *p) return cudaSuccess;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  cudaError_t e = dev_alloc(p, (size_t)need);
  if (e == cudaSuccess) *cap = need;
  return e;
}

// ---------------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------------
extern "C" int dnbc_create(const dnbc_config *cfg, dnbc_ctx **out) {
  if (!cfg || !out) return dnbc_fail(nullptr, DNBC_EINVAL, "dnbc_create: null argument");
  *out = nullptr;
  const int N = cfg->n_states, D = cfg->n_disc, C = cfg->n_cont;
  if (N < 1 || N > 1024) return dnbc_fail(nullptr, DNBC_EINVAL, "number of hidden states must be in [1, 1024]");
  if (D < 0 || C < 0 || D > 255 || C > 255) return dnbc_fail(nullptr, DNBC_EINVAL, "bad number of observed variables");
  if ((D > 0 && !cfg->alphabet) || (C > 0 && !cfg->mixture))
    // the reference's message for a missing / short hint list (DNBC.scala:274, typo included)
    return dnbc_fail(nullptr, DNBC_EINVAL,
                     "Number of continuous variable hints does not correspond to the number of continous variables");
  int Vmax = 1, Kmax = 1;
  for (int d = 0; d < D; ++d) {
    if (cfg->alphabet[d] < 1 || cfg->alphabet[d] > 254)
      return dnbc_fail(nullptr, DNBC_EINVAL, "alphabet size of a discrete variable must be in [1, 254]");
    Vmax = std::max(Vmax, cfg->alphabet[d]);
  }
  for (int c = 0; c < C; ++c) {
    if (cfg->mixture[c] < 1)  // DNBC.scala:276
      return dnbc_fail(nullptr, DNBC_EINVAL,
                       "Number of suggested normal distributions in a continuous variable must be at least one");
    if (cfg->mixture[c] > 32) return dnbc_fail(nullptr, DNBC_EINVAL, "at most 32 normal distributions per mixture");
    Kmax = std::max(Kmax, cfg->mixture[c]);
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev) {
    cudaGetLastError();
    return dnbc_fail(nullptr, DNBC_ENODEVICE,
                     "no usable CUDA device (%s); dnbc_b200 has no CPU path",
                     e != cudaSuccess ? cudaGetErrorString(e) : "device ordinal out of range");
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major < 10)
    return dnbc_fail(nullptr, DNBC_ENODEVICE, "device %d is not an sm_100-class GPU", cfg->device);

  dnbc_ctx *c = new dnbc_ctx();
  c->device = cfg->device;
  c->sm_count = prop.multiProcessorCount;
  c->dm = Dims{N, D, C, Vmax, Kmax, N <= 32 ? next_pow2(N < 2 ? 2 : N) : ((N + 31) / 32) * 32};
  // lane-group kernels are instantiated for Npad in {2..32, 64, 128, 256, 512, 1024}
  if (c->dm.Npad > 32) c->dm.Npad = next_pow2(c->dm.Npad);
  c->V.assign(cfg->alphabet, cfg->alphabet + D);
  c->K.assign(cfg->mixture, cfg->mixture + C);
  const int Npad = c->dm.Npad;
#define CREATE_TRY(expr)                                                           \
  do {                                                                             \
    cudaError_t e__ = (expr);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      int rc__ = dnbc_fail(nullptr, DNBC_ECUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
      dnbc_destroy(c);                                                             \
      return rc__;                                                                 \
    }                                                                              \
  } while (0)
  CREATE_TRY(cudaSetDevice(c->device));
  CREATE_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CREATE_TRY(cudaEventCreate(&c->ev0));
  CREATE_TRY(cudaEventCreate(&c->ev1));
  CREATE_TRY(dev_alloc(&c->d_pi, N));
  CREATE_TRY(dev_alloc(&c->d_A, (size_t)N * N));
  CREATE_TRY(dev_alloc(&c->d_B, (size_t)D * N * Vmax));
  CREATE_TRY(dev_alloc(&c->d_w, (size_t)C * N * Kmax));
  CREATE_TRY(dev_alloc(&c->d_mu, (size_t)C * N * Kmax));
  CREATE_TRY(dev_alloc(&c->d_var, (size_t)C * N * Kmax));
  CREATE_TRY(dev_alloc(&c->d_active, N));
  CREATE_TRY(dev_alloc(&c->d_V, D));
  CREATE_TRY(dev_alloc(&c->d_K, C));
  CREATE_TRY(dev_alloc(&c->t_pi, N));
  CREATE_TRY(dev_alloc(&c->t_logpi, N));
  CREATE_TRY(dev_alloc(&c->t_A, (size_t)N * Npad));
  CREATE_TRY(dev_alloc(&c->t_At, (size_t)N * Npad));
  CREATE_TRY(dev_alloc(&c->t_logA, (size_t)N * Npad));
  CREATE_TRY(dev_alloc(&c->t_cat, (size_t)D * (Vmax + 1) * Npad));
  CREATE_TRY(dev_alloc(&c->t_gs, (size_t)C * Kmax * Npad));
  CREATE_TRY(dev_alloc(&c->t_gm, (size_t)C * Kmax * Npad));
  CREATE_TRY(dev_alloc(&c->t_gc, (size_t)C * Kmax * Npad));
  CREATE_TRY(dev_alloc(&c->t64_logpi, N));
  CREATE_TRY(dev_alloc(&c->t64_logA, (size_t)N * N));
  CREATE_TRY(dev_alloc(&c->t64_logB, (size_t)D * (Vmax + 1) * N));
  CREATE_TRY(dev_alloc(&c->t64_u, (size_t)C * N * Kmax));
  CREATE_TRY(dev_alloc(&c->t64_rs, (size_t)C * N * Kmax));
  c->stats_size = stats_layout(c->dm).total;
  CREATE_TRY(dev_alloc(&c->d_stats, (size_t)c->stats_size));
  CREATE_TRY(dev_alloc(&c->d_cnt, (size_t)N + (size_t)N * N + (size_t)D * N * Vmax));
  CREATE_TRY(dev_alloc(&c->d_first_seq, N));
  CREATE_TRY(dev_alloc(&c->d_pooled, C));
  CREATE_TRY(dev_alloc(&c->d_floored, 1));
  CREATE_TRY(cudaMemset(c->d_floored, 0, sizeof(unsigned long long)));
  c->h_pinned_bytes = std::max<size_t>(1 << 20, (size_t)c->stats_size * sizeof(double));
  CREATE_TRY(cudaMallocHost((void **)&c->h_pinned, c->h_pinned_bytes));
  if (D) CREATE_TRY(cudaMemcpy(c->d_V, c->V.data(), D * sizeof(int32_t), cudaMemcpyHostToDevice));
  if (C) CREATE_TRY(cudaMemcpy(c->d_K, c->K.data(), C * sizeof(int32_t), cudaMemcpyHostToDevice));
  CREATE_TRY(cudaMemset(c->d_active, 1, N));
  CREATE_TRY(cudaMemset(c->d_stats, 0, c->stats_size * sizeof(double)));
#undef CREATE_TRY
  *out = c;
  return DNBC_OK;
}

extern "C" int dnbc_destroy(dnbc_ctx *c) {
  if (!c) return DNBC_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  dnbc_comm_destroy(c);
  void *ptrs[] = {c->d_pi,    c->d_A,     c->d_B,    c->d_w,      c->d_mu,     c->d_var,   c->d_active, c->d_V,
                  c->d_K,     c->t_pi,    c->t_A,    c->t_At,     c->t_logpi,  c->t_logA,  c->t_cat,    c->t_gs,
                  c->t_gm,    c->t_gc,    c->t64_logpi, c->t64_logA, c->t64_logB, c->t64_u, c->t64_rs,  c->d_off,
                  c->d_disc,  c->d_cont,  c->d_cont64, c->d_labels, c->E,      c->AL,      c->GA,       c->U,
                  c->cs,      c->d_stats, c->E64,    c->d_bp,     c->d_path,   c->d_score, c->d_tie,    c->d_cnt,
                  c->d_first_seq, c->d_assign, c->d_edge, c->d_edge_i, c->tc_tab, c->tc_x, c->tc_mrow, c->vt_logA, c->d_pooled, c->d_floored};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  for (auto &p : c->ev_pool) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (auto e : c->chunk_events) cudaEventDestroy(e);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return DNBC_OK;
}

// ---------------------------------------------------------------------------------
// parameters
// ---------------------------------------------------------------------------------
extern "C" int dnbc_set_params(dnbc_ctx *c, const double *pi, const double *A, const double *B, const double *w,
                               const double *mu, const double *var, const uint8_t *active) {
  if (!c) return DNBC_EINVAL;
  const Dims &d = c->dm;
  if (!pi || !A || (d.D && !B) || (d.C && (!w || !mu || !var))) return dnbc_fail(c, DNBC_EINVAL, "set_params: null array");
  CU_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  CU_TRY(c, cudaMemcpyAsync(c->d_pi, pi, d.N * sizeof(double), cudaMemcpyHostToDevice, st));
  CU_TRY(c, cudaMemcpyAsync(c->d_A, A, (size_t)d.N * d.N * sizeof(double), cudaMemcpyHostToDevice, st));
  if (d.D) CU_TRY(c, cudaMemcpyAsync(c->d_B, B, (size_t)d.D * d.N * d.Vmax * sizeof(double), cudaMemcpyHostToDevice, st));
  if (d.C) {
    const size_t n = (size_t)d.C * d.N * d.Kmax * sizeof(double);
    CU_TRY(c, cudaMemcpyAsync(c->d_w, w, n, cudaMemcpyHostToDevice, st));
    CU_TRY(c, cudaMemcpyAsync(c->d_mu, mu, n, cudaMemcpyHostToDevice, st));
    CU_TRY(c, cudaMemcpyAsync(c->d_var, var, n, cudaMemcpyHostToDevice, st));
  }
  if (active) CU_TRY(c, cudaMemcpyAsync(c->d_active, active, d.N, cudaMemcpyHostToDevice, st));
  else CU_TRY(c, cudaMemsetAsync(c->d_active, 1, d.N, st));
  c->tables64_valid = false;
  launch_prepare_tables(c);
  CU_TRY(c, cudaStreamSynchronize(st));  // the caller may free its arrays on return
  CU_TRY(c, cudaGetLastError());
  return DNBC_OK;
}

extern "C" int dnbc_get_params(dnbc_ctx *c, double *pi, double *A, double *B, double *w, double *mu, double *var,
                               uint8_t *active) {
  if (!c) return DNBC_EINVAL;
  const Dims &d = c->dm;
  CU_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  if (pi) CU_TRY(c, cudaMemcpyAsync(pi, c->d_pi, d.N * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (A) CU_TRY(c, cudaMemcpyAsync(A, c->d_A, (size_t)d.N * d.N * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (B && d.D) CU_TRY(c, cudaMemcpyAsync(B, c->d_B, (size_t)d.D * d.N * d.Vmax * sizeof(double), cudaMemcpyDeviceToHost, st));
  const size_t n = (size_t)d.C * d.N * d.Kmax * sizeof(double);
  if (w && d.C) CU_TRY(c, cudaMemcpyAsync(w, c->d_w, n, cudaMemcpyDeviceToHost, st));
  if (mu && d.C) CU_TRY(c, cudaMemcpyAsync(mu, c->d_mu, n, cudaMemcpyDeviceToHost, st));
  if (var && d.C) CU_TRY(c, cudaMemcpyAsync(var, c->d_var, n, cudaMemcpyDeviceToHost, st));
  if (active) CU_TRY(c, cudaMemcpyAsync(active, c->d_active, d.N, cudaMemcpyDeviceToHost, st));
  CU_TRY(c, cudaStreamSynchronize(st));
  return DNBC_OK;
}

// fp64 decode tables: natural logs with the host libm (see file header)
static int build_tables64(dnbc_ctx *c) {
  if (c->tables64_valid) return DNBC_OK;
  const Dims &d = c->dm;
  const int N = d.N, D = d.D, C = d.C, Vmax = d.Vmax, Kmax = d.Kmax;
  std::vector<double> pi(N), A((size_t)N * N), B((size_t)D * N * Vmax), var((size_t)C * N * Kmax);
  int rc = dnbc_get_params(c, pi.data(), A.data(), B.data(), nullptr, nullptr, var.data(), nullptr);
  if (rc) return rc;
  std::vector<double> lpi(N), lA((size_t)N * N), lB((size_t)D * (Vmax + 1) * N), u(var.size()), rs(var.size());
  for (int j = 0; j < N; ++j) lpi[j] = log(pi[j]);
  for (size_t q = 0; q < lA.size(); ++q) lA[q] = log(A[q]);
  for (int dd = 0; dd < D; ++dd)
    for (int v = 0; v <= Vmax; ++v)
      for (int j = 0; j < N; ++j) {
        double p = v < c->V[dd] ? B[((size_t)dd * N + j) * Vmax + v] : 0.0;
        lB[((size_t)dd * (Vmax + 1) + v) * N + j] = log(p);
      }
  // spark-mllib 2.1.0 MultivariateGaussian (1x1): u = -0.5 (ln 2pi + ln var), rootSigmaInv = sqrt(1/var);
  // a non-positive variance falls to the pseudo-inverse branch: u = -0.5 ln 2pi, rootSigmaInv = 0
  for (size_t q = 0; q < var.size(); ++q) {
    if (var[q] > 0.0) {
      u[q] = -0.5 * (log(2.0 * DNBC_PI) + log(var[q]));
      rs[q] = sqrt(1.0 / var[q]);
    } else {
      u[q] = -0.5 * (log(2.0 * DNBC_PI) + 0.0);
      rs[q] = 0.0;
    }
  }
  cudaStream_t st = c->stream;
  CU_TRY(c, cudaMemcpyAsync(c->t64_logpi, lpi.data(), lpi.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  CU_TRY(c, cudaMemcpyAsync(c->t64_logA, lA.data(), lA.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  if (!lB.empty()) CU_TRY(c, cudaMemcpyAsync(c->t64_logB, lB.data(), lB.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  if (!u.empty()) {
    CU_TRY(c, cudaMemcpyAsync(c->t64_u, u.data(), u.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CU_TRY(c, cudaMemcpyAsync(c->t64_rs, rs.data(), rs.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  CU_TRY(c, cudaStreamSynchronize(st));
  c->tables64_valid = true;
  return DNBC_OK;
}

// ---------------------------------------------------------------------------------
// observations
// ---------------------------------------------------------------------------------
static int reserve_points(dnbc_ctx *c, int64_t S, int64_t P, bool want64, bool want_labels) {
  const Dims &d = c->dm;
  if (P > c->cap_points || !c->d_disc) {
    void *old[] = {c->d_disc, c->d_cont, c->d_cont64, c->d_labels};
    for (void *p : old)
      if (p) cudaFree(p);
    c->d_disc = nullptr; c->d_cont = nullptr; c->d_cont64 = nullptr; c->d_labels = nullptr;
    c->cap_points = 0;
    CU_TRY(c, dev_alloc(&c->d_disc, (size_t)d.D * P));
    CU_TRY(c, dev_alloc(&c->d_cont, (size_t)d.C * P));
    c->cap_points = P;
  }
  if (want64 && !c->d_cont64) CU_TRY(c, dev_alloc(&c->d_cont64, (size_t)d.C * c->cap_points));
  if (want_labels && !c->d_labels) CU_TRY(c, dev_alloc(&c->d_labels, (size_t)c->cap_points));
  if (S + 1 > c->cap_seqs || !c->d_off) {
    if (c->d_off) cudaFree(c->d_off);
    c->d_off = nullptr;
    CU_TRY(c, dev_alloc(&c->d_off, (size_t)S + 1));
    c->cap_seqs = S + 1;
  }
  return DNBC_OK;
}

extern "C" int dnbc_upload(dnbc_ctx *c, int64_t n_seq, const int64_t *offsets, const uint8_t *disc,
                           const double *cont_f64, const float *cont_f32, const int32_t *labels) {
  if (!c) return DNBC_EINVAL;
  const Dims &d = c->dm;
  if (n_seq < 0 || !offsets) return dnbc_fail(c, DNBC_EINVAL, "upload: null offsets");
  if (offsets[0] != 0) return dnbc_fail(c, DNBC_EINVAL, "upload: offsets[0] must be 0");
  int64_t Tmax = 0;
  for (int64_t s = 0; s < n_seq; ++s) {
    if (offsets[s + 1] < offsets[s]) return dnbc_fail(c, DNBC_EINVAL, "upload: offsets must be non-decreasing");
    Tmax = std::max(Tmax, offsets[s + 1] - offsets[s]);
  }
  const int64_t P = offsets[n_seq];
  // shape errors carry the reference's message (DNBC.scala:46, :148)
  if (d.D > 0 && P > 0 && !disc) return dnbc_fail(c, DNBC_EINVAL, "Number of observed variables is inconsistent");
  if (d.C > 0 && P > 0 && !cont_f64 && !cont_f32)
    return dnbc_fail(c, DNBC_EINVAL, "Number of observed variables is inconsistent");
  if (cont_f64 && cont_f32) return dnbc_fail(c, DNBC_EINVAL, "upload: pass exactly one of cont_f64 / cont_f32");
  CU_TRY(c, cudaSetDevice(c->device));
  int rc = reserve_points(c, n_seq, P, cont_f64 != nullptr && d.C > 0, labels != nullptr);
  if (rc) return rc;
  cudaStream_t st = c->stream;
  CU_TRY(c, cudaMemcpyAsync(c->d_off, offsets, (size_t)(n_seq + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  // planes are [var][P] on both sides, but the device capacity may exceed P: copy plane by plane
  for (int v = 0; v < d.D && P > 0; ++v)
    CU_TRY(c, cudaMemcpyAsync(c->d_disc + (size_t)v * P, disc + (size_t)v * P, (size_t)P, cudaMemcpyHostToDevice, st));
  if (d.C > 0 && P > 0) {
    if (cont_f32) {
      CU_TRY(c, cudaMemcpyAsync(c->d_cont, cont_f32, (size_t)d.C * P * sizeof(float), cudaMemcpyHostToDevice, st));
    } else {
      CU_TRY(c, cudaMemcpyAsync(c->d_cont64, cont_f64, (size_t)d.C * P * sizeof(double), cudaMemcpyHostToDevice, st));
      launch_f64_to_f32(c, c->d_cont64, c->d_cont, (int64_t)d.C * P);
    }
  }
  if (labels && P > 0) CU_TRY(c, cudaMemcpyAsync(c->d_labels, labels, (size_t)P * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  c->h_off.assign(offsets, offsets + n_seq + 1);
  c->S = n_seq;
  c->P = P;
  c->Tmax = Tmax;
  c->have_labels = labels != nullptr;
  c->have_cont64 = cont_f64 != nullptr && d.C > 0;
  CU_TRY(c, cudaStreamSynchronize(st));
  return DNBC_OK;
}

extern "C" int dnbc_generate(dnbc_ctx *c, int64_t n_seq, int32_t seq_len, uint64_t seed) {
  if (!c) return DNBC_EINVAL;
  if (n_seq < 0 || seq_len < 1) return dnbc_fail(c, DNBC_EINVAL, "generate: bad shape");
  CU_TRY(c, cudaSetDevice(c->device));
  const int64_t P = n_seq * (int64_t)seq_len;
  int rc = reserve_points(c, n_seq, P, false, true);
  if (rc) return rc;
  c->h_off.resize(n_seq + 1);
  for (int64_t s = 0; s <= n_seq; ++s) c->h_off[s] = s * (int64_t)seq_len;
  CU_TRY(c, cudaMemcpyAsync(c->d_off, c->h_off.data(), (size_t)(n_seq + 1) * sizeof(int64_t), cudaMemcpyHostToDevice,
                            c->stream));
  c->S = n_seq;
  c->P = P;
  c->Tmax = seq_len;
  c->have_labels = true;
  c->have_cont64 = false;
  launch_generate(c, n_seq, seq_len, seed);
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  CU_TRY(c, cudaGetLastError());
  return DNBC_OK;
}

extern "C" int dnbc_download(dnbc_ctx *c, int64_t first, int64_t count, int64_t *offsets, uint8_t *disc,
                             double *cont_f64, int32_t *labels) {
  if (!c) return DNBC_EINVAL;
  if (first < 0 || count < 0 || first + count > c->S) return dnbc_fail(c, DNBC_EINVAL, "download: range outside the data");
  const Dims &d = c->dm;
  CU_TRY(c, cudaSetDevice(c->device));
  const int64_t p0 = c->h_off[first], p1 = c->h_off[first + count], np = p1 - p0;
  if (offsets)
    for (int64_t s = 0; s <= count; ++s) offsets[s] = c->h_off[first + s] - p0;
  cudaStream_t st = c->stream;
  if (disc)
    for (int v = 0; v < d.D && np > 0; ++v)
      CU_TRY(c, cudaMemcpyAsync(disc + (size_t)v * np, c->d_disc + (size_t)v * c->P + p0, (size_t)np, cudaMemcpyDeviceToHost, st));
  if (cont_f64 && d.C > 0 && np > 0) {
    if (c->have_cont64) {
      for (int v = 0; v < d.C; ++v)
        CU_TRY(c, cudaMemcpyAsync(cont_f64 + (size_t)v * np, c->d_cont64 + (size_t)v * c->P + p0, (size_t)np * sizeof(double),
                                  cudaMemcpyDeviceToHost, st));
    } else {
      // widen on the device into the E64 scratch, plane by plane
      CU_TRY(c, ensure(&c->E64, &c->e64_points, np));
      for (int v = 0; v < d.C; ++v) {
        launch_f32_to_f64(c, c->d_cont + (size_t)v * c->P + p0, c->E64, np);
        CU_TRY(c, cudaMemcpyAsync(cont_f64 + (size_t)v * np, c->E64, (size_t)np * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU_TRY(c, cudaStreamSynchronize(st));
      }
    }
  }
  if (labels && np > 0) {
    if (!c->have_labels) return dnbc_fail(c, DNBC_ESTATE, "download: no labels resident");
    CU_TRY(c, cudaMemcpyAsync(labels, c->d_labels + p0, (size_t)np * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  }
  CU_TRY(c, cudaStreamSynchronize(st));
  return DNBC_OK;
}

extern "C" int dnbc_download_f32(dnbc_ctx *c, int64_t first, int64_t count, uint8_t *disc, float *cont_f32) {
  if (!c) return DNBC_EINVAL;
  if (first < 0 || count < 0 || first + count > c->S) return dnbc_fail(c, DNBC_EINVAL, "download: range outside the data");
  const Dims &d = c->dm;
  CU_TRY(c, cudaSetDevice(c->device));
  const int64_t p0 = c->h_off[first], np = c->h_off[first + count] - p0;
  for (int v = 0; v < d.D && np > 0 && disc; ++v)
    CU_TRY(c, cudaMemcpyAsync(disc + (size_t)v * np, c->d_disc + (size_t)v * c->P + p0, (size_t)np, cudaMemcpyDeviceToHost, c->stream));
  for (int v = 0; v < d.C && np > 0 && cont_f32; ++v)
    CU_TRY(c, cudaMemcpyAsync(cont_f32 + (size_t)v * np, c->d_cont + (size_t)v * c->P + p0, (size_t)np * sizeof(float),
                              cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return DNBC_OK;
}

extern "C" int dnbc_data_shape(const dnbc_ctx *c, int64_t *n_seq, int64_t *n_points) {
  if (!c) return DNBC_EINVAL;
  if (n_seq) *n_seq = c->S;
  if (n_points) *n_points = c->P;
  return DNBC_OK;
}

// ---------------------------------------------------------------------------------
// chunking: the EM / decode pipelines materialise [points][N] scratch arrays, so the
// resident sequences are walked in chunks of whole sequences.
// ---------------------------------------------------------------------------------
struct Chunk { int64_t s0, s1, p0, np; };

// tensor-core recursions (65 .. 512 states): a chunk that is not the last one ends on a whole number of waves of
// 128-sequence batches (k_tc.cu: fb_tc_wave_sequences)
static int64_t whole_waves(const dnbc_ctx *c, int64_t s, int64_t e) {
  const int64_t wave = fb_tc_wave_sequences(c);
  if (wave <= 0 || e >= c->S || e - s < wave) return e;
  return s + (e - s) / wave * wave;
}

static std::vector<Chunk> make_chunks(const dnbc_ctx *c, int64_t max_points) {
  std::vector<Chunk> out;
  int64_t s = 0;
  while (s < c->S) {
    int64_t e = s;
    const int64_t p0 = c->h_off[s];
    // largest e with off[e] - p0 <= max_points (at least one sequence)
    e = std::upper_bound(c->h_off.begin() + s + 1, c->h_off.end(), p0 + max_points) - c->h_off.begin() - 1;
    if (e <= s) e = s + 1;
    if (e > c->S) e = c->S;
    e = whole_waves(c, s, e);
    out.push_back(Chunk{s, e, p0, c->h_off[e] - p0});
    s = e;
  }
  return out;
}

// streamed upload: the first chunk's copy cannot overlap anything, so chunks start small and
// double until they reach the scratch size (a kernel launch over few sequences still walks all
// their timesteps, so many small chunks would leave the recursion kernels mostly idle)
static std::vector<Chunk> make_chunks_progressive(const dnbc_ctx *c, int64_t first_points, int64_t max_points) {
  std::vector<Chunk> out;
  int64_t s = 0, want = std::min(first_points, max_points);
  while (s < c->S) {
    const int64_t p0 = c->h_off[s];
    int64_t e = std::upper_bound(c->h_off.begin() + s + 1, c->h_off.end(), p0 + want) - c->h_off.begin() - 1;
    if (e <= s) e = s + 1;
    if (e > c->S) e = c->S;
    e = whole_waves(c, s, e);
    out.push_back(Chunk{s, e, p0, c->h_off[e] - p0});
    s = e;
    want = std::min(want * 2, max_points);
  }
  return out;
}

// points the E-step scratch should hold: everything resident if it already fits, otherwise a
// budget from the free memory (cudaMemGetInfo costs milliseconds, so it is only asked when the
// scratch has to grow)
static int64_t chunk_budget_points(const dnbc_ctx *c, size_t bytes_per_point);
static int64_t em_scratch_target(const dnbc_ctx *c) {
  if (c->E && c->scratch_points >= c->P) return c->scratch_points;
  if (c->E && c->scratch_points >= ((int64_t)4 << 20) && c->scratch_points >= c->Tmax) return c->scratch_points;
  return chunk_budget_points(c, (size_t)c->dm.N * 4 * sizeof(float) + 4);
}

static int64_t chunk_budget_points(const dnbc_ctx *c, size_t bytes_per_point) {
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  size_t held = 0;  // scratch we already own can be reused
  // up to a third of the free HBM: large chunks keep >= 148 batches of 128 sequences per launch at
  // N = 512 (8 KB of scratch per point) and cut launch counts at N = 64
  size_t budget = std::min<size_t>((free_b + held) / 3, (size_t)48 << 30);
  int64_t pts = (int64_t)(budget / std::max<size_t>(bytes_per_point, 1));
  pts = std::min<int64_t>(pts, (int64_t)1 << 30);   // kernels index chunk-relative points with 32 bits
  if (c->opt_max_chunk_points > 0) pts = std::min<int64_t>(pts, c->opt_max_chunk_points);   // DNBC_OPT_MAX_CHUNK_POINTS
  pts = std::max<int64_t>(pts, c->Tmax);
  return std::min<int64_t>(pts, std::max<int64_t>(c->P, 1));
}

static int reserve_em_scratch(dnbc_ctx *c, int64_t points) {
  if (points <= c->scratch_points && c->E) return DNBC_OK;
  void *old[] = {c->E, c->AL, c->GA, c->U, c->cs};
  for (void *p : old)
    if (p) cudaFree(p);
  c->E = c->AL = c->GA = c->U = c->cs = nullptr;
  c->scratch_points = 0;
  const size_t n = (size_t)points * c->dm.N;
  CU_TRY(c, dev_alloc(&c->E, n));
  CU_TRY(c, dev_alloc(&c->AL, n));
  CU_TRY(c, dev_alloc(&c->GA, n));
  CU_TRY(c, dev_alloc(&c->U, n));
  CU_TRY(c, dev_alloc(&c->cs, (size_t)points));
  c->scratch_points = points;
  return DNBC_OK;
}

// ---------------------------------------------------------------------------------
// K1 through the ABI
// ---------------------------------------------------------------------------------
extern "C" int dnbc_emission_logprob(dnbc_ctx *c, int precision, void *out) {
  if (!c || !out) return DNBC_EINVAL;
  if (c->P == 0) return DNBC_OK;
  CU_TRY(c, cudaSetDevice(c->device));
  const int N = c->dm.N;
  if (precision == DNBC_PREC_F64) {
    int rc = build_tables64(c);
    if (rc) return rc;
    const int64_t maxp = std::min<int64_t>(c->P, std::max<int64_t>(1, ((int64_t)1 << 28) / N));
    CU_TRY(c, ensure(&c->E64, &c->e64_points, maxp * N));
    for (int64_t p0 = 0; p0 < c->P; p0 += maxp) {
      const int64_t np = std::min(maxp, c->P - p0);
      launch_emission_f64(c, p0, np, c->E64);
      CU_TRY(c, cudaMemcpyAsync((double *)out + p0 * N, c->E64, (size_t)np * N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      CU_TRY(c, cudaStreamSynchronize(c->stream));
    }
  } else {
    const int64_t maxp = chunk_budget_points(c, (size_t)N * 4 * sizeof(float) + 4);
    int rc = reserve_em_scratch(c, std::max(maxp, c->scratch_points));
    if (rc) return rc;
    for (int64_t p0 = 0; p0 < c->P; p0 += c->scratch_points) {
      const int64_t np = std::min(c->scratch_points, c->P - p0);
      launch_emission_f32(c, p0, np, c->E);
      CU_TRY(c, cudaMemcpyAsync((float *)out + p0 * N, c->E, (size_t)np * N * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
      CU_TRY(c, cudaStreamSynchronize(c->stream));
    }
  }
  CU_TRY(c, cudaGetLastError());
  return DNBC_OK;
}

// ---------------------------------------------------------------------------------
// supervised learner
// ---------------------------------------------------------------------------------
extern "C" int dnbc_supervised_counts(dnbc_ctx *c, int flags, int64_t *cnt_pi, int64_t *cnt_A, int64_t *cnt_B) {
  if (!c) return DNBC_EINVAL;
  if (!c->have_labels) return dnbc_fail(c, DNBC_ESTATE, "supervised learning needs labels");
  CU_TRY(c, cudaSetDevice(c->device));
  const Dims &d = c->dm;
  unsigned long long *pi = c->d_cnt, *A = pi + d.N, *B = A + (size_t)d.N * d.N;
  launch_supervised_counts(c, flags, pi, A, B);
  cudaStream_t st = c->stream;
  if (cnt_pi) CU_TRY(c, cudaMemcpyAsync(cnt_pi, pi, d.N * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  if (cnt_A) CU_TRY(c, cudaMemcpyAsync(cnt_A, A, (size_t)d.N * d.N * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  if (cnt_B && d.D) CU_TRY(c, cudaMemcpyAsync(cnt_B, B, (size_t)d.D * d.N * d.Vmax * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CU_TRY(c, cudaStreamSynchronize(st));
  CU_TRY(c, cudaGetLastError());
  return DNBC_OK;
}

extern "C" int dnbc_mle(dnbc_ctx *c, int flags, const double *init_centroids, int max_iters, double rel_tol,
                        double *ll_series, int32_t *n_iters) {
  if (!c) return DNBC_EINVAL;
  if (!c->have_labels) return dnbc_fail(c, DNBC_ESTATE, "supervised learning needs labels");
  CU_TRY(c, cudaSetDevice(c->device));
  const Dims &d = c->dm;
  unsigned long long *pi = c->d_cnt, *A = pi + d.N, *B = A + (size_t)d.N * d.N;
  launch_supervised_counts(c, flags, pi, A, B);
  launch_normalize_counts(c, pi, A, B);
  if (d.C > 0) {
    int rc = run_gmm_mle(c, flags, init_centroids, max_iters <= 0 ? 30 : max_iters, rel_tol <= 0.0 ? 0.01 : rel_tol,
                         ll_series, n_iters);
    if (rc) return rc;
  }
  c->tables64_valid = false;
  launch_prepare_tables(c);
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  CU_TRY(c, cudaGetLastError());
  return DNBC_OK;
}

// ---------------------------------------------------------------------------------
// Baum-Welch
// ---------------------------------------------------------------------------------
extern "C" int64_t dnbc_stats_size(const dnbc_ctx *c) { return c ? c->stats_size : 0; }

// E-step over the resident sequences, chunk by chunk.  When `ready` is given, chunk k first
// waits for event ready[k] (its observations are still in flight on the copy stream).
static int estep_chunks(dnbc_ctx *c, int clamp_labels, const std::vector<Chunk> &chunks, const std::vector<cudaEvent_t> *ready) {
  for (size_t k = 0; k < chunks.size(); ++k) {
    const Chunk &ch = chunks[k];
    if (ch.np > c->scratch_points) {
      int rc = reserve_em_scratch(c, ch.np);
      if (rc) return rc;
    }
    if (ready) CU_TRY(c, cudaStreamWaitEvent(c->stream, (*ready)[k], 0));
    if (!clamp_labels && fused_estep_wanted(c, ch.s1 - ch.s0)) {   // N <= 16, few sequences: the whole E-step in one launch (k_fused.cu)
      launch_estep_fused(c, ch.s0, ch.s1, ch.p0);
      continue;
    }
    launch_emission_f32(c, ch.p0, ch.np, c->E);
    if (clamp_labels) launch_clamp_labels(c, ch.p0, ch.np, c->E);
    launch_forward(c, ch.s0, ch.s1, ch.p0);
    launch_backward(c, ch.s0, ch.s1, ch.p0);
    launch_xi(c, ch.np);
    if (mixture_stats_fast(c)) {
      launch_mixture_stats(c, ch.p0, ch.np);
    } else {
      launch_stats_gmm(c, ch.p0, ch.np);
      launch_stats_cat(c, ch.p0, ch.np);
    }
  }
  CU_TRY(c, cudaGetLastError());
  return DNBC_OK;
}

extern "C" int dnbc_em_estep(dnbc_ctx *c, int clamp_labels) {
  if (!c) return DNBC_EINVAL;
  if (clamp_labels && !c->have_labels) return dnbc_fail(c, DNBC_ESTATE, "clamped E-step needs labels");
  CU_TRY(c, cudaSetDevice(c->device));
  CU_TRY(c, cudaMemsetAsync(c->d_stats, 0, (size_t)c->stats_size * sizeof(double), c->stream));
  if (c->S == 0) return DNBC_OK;
  const int64_t maxp = em_scratch_target(c);
  int rc = reserve_em_scratch(c, std::max(maxp, c->scratch_points));
  if (rc) return rc;
  return estep_chunks(c, clamp_labels, make_chunks(c, c->scratch_points), nullptr);
}

// Upload + E-step in one call, the host->device copy of chunk k+1 overlapping the kernels of
// chunk k (copy stream + per-chunk events).  The observations stay resident afterwards exactly
// as after dnbc_upload.  Host buffers should be pinned for the copies to be asynchronous.
extern "C" int dnbc_em_estep_streamed(dnbc_ctx *c, int64_t n_seq, const int64_t *offsets, const uint8_t *disc,
                                      const float *cont_f32, const int32_t *labels, int clamp_labels) {
  if (!c) return DNBC_EINVAL;
  const Dims &d = c->dm;
  if (n_seq < 0 || !offsets || offsets[0] != 0) return dnbc_fail(c, DNBC_EINVAL, "estep_streamed: bad offsets");
  int64_t Tmax = 0;
  for (int64_t s = 0; s < n_seq; ++s) {
    if (offsets[s + 1] < offsets[s]) return dnbc_fail(c, DNBC_EINVAL, "estep_streamed: offsets must be non-decreasing");
    Tmax = std::max(Tmax, offsets[s + 1] - offsets[s]);
  }
  const int64_t P = offsets[n_seq];
  if ((d.D > 0 && P > 0 && !disc) || (d.C > 0 && P > 0 && !cont_f32))
    return dnbc_fail(c, DNBC_EINVAL, "Number of observed variables is inconsistent");   // DNBC.scala:46, :148
  if (clamp_labels && !labels) return dnbc_fail(c, DNBC_ESTATE, "clamped E-step needs labels");
  CU_TRY(c, cudaSetDevice(c->device));
  int rc = reserve_points(c, n_seq, P, false, labels != nullptr);
  if (rc) return rc;
  c->h_off.assign(offsets, offsets + n_seq + 1);
  c->S = n_seq; c->P = P; c->Tmax = Tmax;
  c->have_labels = labels != nullptr;
  c->have_cont64 = false;
  CU_TRY(c, cudaMemcpyAsync(c->d_off, offsets, (size_t)(n_seq + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(c, cudaMemsetAsync(c->d_stats, 0, (size_t)c->stats_size * sizeof(double), c->stream));
  if (n_seq == 0) return DNBC_OK;
  const int64_t maxp = em_scratch_target(c);
  rc = reserve_em_scratch(c, std::max(maxp, c->scratch_points));
  if (rc) return rc;
  const std::vector<Chunk> chunks = make_chunks_progressive(c, std::max<int64_t>(Tmax, (int64_t)2 << 20), c->scratch_points);
  if (!c->copy_stream) CU_TRY(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  while (c->chunk_events.size() < chunks.size() + 1) {
    cudaEvent_t e;
    CU_TRY(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->chunk_events.push_back(e);
  }
  // the copy stream must not overwrite planes that earlier kernels are still reading
  cudaEvent_t fence = c->chunk_events[chunks.size()];
  CU_TRY(c, cudaEventRecord(fence, c->stream));
  CU_TRY(c, cudaStreamWaitEvent(c->copy_stream, fence, 0));
  for (size_t k = 0; k < chunks.size(); ++k) {
    const Chunk &ch = chunks[k];
    for (int v = 0; v < d.D; ++v)
      CU_TRY(c, cudaMemcpyAsync(c->d_disc + (size_t)v * P + ch.p0, disc + (size_t)v * P + ch.p0, (size_t)ch.np,
                                cudaMemcpyHostToDevice, c->copy_stream));
    for (int v = 0; v < d.C; ++v)
      CU_TRY(c, cudaMemcpyAsync(c->d_cont + (size_t)v * P + ch.p0, cont_f32 + (size_t)v * P + ch.p0,
                                (size_t)ch.np * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
    if (labels)
      CU_TRY(c, cudaMemcpyAsync(c->d_labels + ch.p0, labels + ch.p0, (size_t)ch.np * sizeof(int32_t), cudaMemcpyHostToDevice,
                                c->copy_stream));
    CU_TRY(c, cudaEventRecord(c->chunk_events[k], c->copy_stream));
  }
  std::vector<cudaEvent_t> ready(c->chunk_events.begin(), c->chunk_events.begin() + chunks.size());
  return estep_chunks(c, clamp_labels, chunks, &ready);
}

extern "C" int dnbc_em_stats_device_ptr(dnbc_ctx *c, void **ptr) {
  if (!c || !ptr) return DNBC_EINVAL;
  *ptr = c->d_stats;
  return DNBC_OK;
}

extern "C" int dnbc_em_get_stats(dnbc_ctx *c, double *stats) {
  if (!c || !stats) return DNBC_EINVAL;
  CU_TRY(c, cudaSetDevice(c->device));
  CU_TRY(c, cudaMemcpyAsync(stats, c->d_stats, (size_t)c->stats_size * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return DNBC_OK;
}

extern "C" int dnbc_em_set_stats(dnbc_ctx *c, const double *stats) {
  if (!c || !stats) return DNBC_EINVAL;
  CU_TRY(c, cudaSetDevice(c->device));
  CU_TRY(c, cudaMemcpyAsync(c->d_stats, stats, (size_t)c->stats_size * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  return DNBC_OK;
}

extern "C" int dnbc_em_mstep(dnbc_ctx *c, int64_t total_sequences) {
  if (!c) return DNBC_EINVAL;
  if (total_sequences <= 0) return dnbc_fail(c, DNBC_EINVAL, "mstep: total_sequences must be positive");
  CU_TRY(c, cudaSetDevice(c->device));
  launch_mstep(c, total_sequences);
  CU_TRY(c, cudaGetLastError());
  return DNBC_OK;
}

extern "C" int dnbc_set_option(dnbc_ctx *c, int option, int64_t value) {
  if (!c) return DNBC_EINVAL;
  switch (option) {
    case DNBC_OPT_MAX_CHUNK_POINTS:
      if (value < 0) return dnbc_fail(c, DNBC_EINVAL, "set_option: chunk size must be >= 0");
      c->opt_max_chunk_points = value;
      // the scratch arrays shrink with the next E-step / decode
      CU_TRY(c, cudaSetDevice(c->device));
      CU_TRY(c, cudaStreamSynchronize(c->stream));
      {
        void *old[] = {c->E, c->AL, c->GA, c->U, c->cs};
        for (void *p : old)
          if (p) cudaFree(p);
        c->E = c->AL = c->GA = c->U = c->cs = nullptr;
        c->scratch_points = 0;
      }
      return DNBC_OK;
    case DNBC_OPT_FB2_INTERLEAVE:
      c->opt_fb2_interleave = value != 0;
      return DNBC_OK;
    case DNBC_OPT_FUSED_ESTEP:
      if (value < 0 || value > 2) return dnbc_fail(c, DNBC_EINVAL, "set_option: fused E-step mode must be 0, 1 or 2");
      c->opt_fused_estep = (int)value;
      return DNBC_OK;
    case DNBC_OPT_FB_SEQ_LANES:
      if (value < 0 || value > 2) return dnbc_fail(c, DNBC_EINVAL, "set_option: lane-per-sequence mode must be 0, 1 or 2");
      c->opt_fb_seq_lanes = (int)value;
      return DNBC_OK;
    default:
      return dnbc_fail(c, DNBC_EINVAL, "set_option: unknown option %d", option);
  }
}

extern "C" int dnbc_em_set_variance_floor(dnbc_ctx *c, double rel_floor) {
  if (!c) return DNBC_EINVAL;
  if (!(rel_floor >= 0.0)) return dnbc_fail(c, DNBC_EINVAL, "variance floor must be >= 0");
  c->var_floor_rel = rel_floor;
  return DNBC_OK;
}

extern "C" int dnbc_em_floored_components(dnbc_ctx *c, int64_t *count) {
  if (!c || !count) return DNBC_EINVAL;
  CU_TRY(c, cudaSetDevice(c->device));
  unsigned long long v = 0;
  CU_TRY(c, cudaMemcpyAsync(&v, c->d_floored, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  *count = (int64_t)v;
  return DNBC_OK;
}

extern "C" int dnbc_em_iterate(dnbc_ctx *c, int n_iters, double *loglik_out) {
  if (!c) return DNBC_EINVAL;
  if (c->S == 0) return dnbc_fail(c, DNBC_ESTATE, "em_iterate: no observations resident");
  for (int i = 0; i < n_iters; ++i) {
    int rc = dnbc_em_estep(c, 0);
    if (rc) return rc;
    if (loglik_out)
      CU_TRY(c, cudaMemcpyAsync(loglik_out + i, c->d_stats, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    rc = dnbc_em_mstep(c, c->S);
    if (rc) return rc;
  }
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  CU_TRY(c, cudaGetLastError());
  return DNBC_OK;
}

extern "C" int dnbc_em_get_posteriors(dnbc_ctx *c, float *gamma) {
  if (!c || !gamma) return DNBC_EINVAL;
  CU_TRY(c, cudaSetDevice(c->device));
  if (c->S == 0) return DNBC_OK;
  const int N = c->dm.N;
  const int64_t maxp = chunk_budget_points(c, (size_t)N * 4 * sizeof(float) + 4);
  int rc = reserve_em_scratch(c, std::max(maxp, c->scratch_points));
  if (rc) return rc;
  // recomputed chunk by chunk (the E-step keeps no [P][N] array); statistics are untouched
  double *saved = nullptr;
  CU_TRY(c, dev_alloc(&saved, (size_t)c->stats_size));
  CU_TRY(c, cudaMemcpyAsync(saved, c->d_stats, (size_t)c->stats_size * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  for (const Chunk &ch : make_chunks(c, c->scratch_points)) {
    if (ch.np > c->scratch_points) {
      rc = reserve_em_scratch(c, ch.np);
      if (rc) { cudaFree(saved); return rc; }
    }
    launch_emission_f32(c, ch.p0, ch.np, c->E);
    launch_forward(c, ch.s0, ch.s1, ch.p0);
    launch_backward(c, ch.s0, ch.s1, ch.p0);
    cudaMemcpyAsync(gamma + ch.p0 * N, c->GA, (size_t)ch.np * N * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
  }
  cudaMemcpyAsync(c->d_stats, saved, (size_t)c->stats_size * sizeof(double), cudaMemcpyDeviceToDevice, c->stream);
  cudaStreamSynchronize(c->stream);
  cudaFree(saved);
  CU_TRY(c, cudaGetLastError());
  return DNBC_OK;
}

// ---------------------------------------------------------------------------------
// decode
// ---------------------------------------------------------------------------------
extern "C" int dnbc_decode(dnbc_ctx *c, int mode, int precision, uint16_t *path_out, double *score_out,
                           uint8_t *tie_out) {
  if (!c) return DNBC_EINVAL;
  if (mode != DNBC_DECODE_REFERENCE && mode != DNBC_DECODE_BACKTRACE) return dnbc_fail(c, DNBC_EINVAL, "decode: bad mode");
  if (precision != DNBC_PREC_F32 && precision != DNBC_PREC_F64) return dnbc_fail(c, DNBC_EINVAL, "decode: bad precision");
  CU_TRY(c, cudaSetDevice(c->device));
  if (c->S == 0) return DNBC_OK;
  const int N = c->dm.N;
  const bool f64 = precision == DNBC_PREC_F64;
  if (f64) {
    int rc = build_tables64(c);
    if (rc) return rc;
  }
  CU_TRY(c, ensure(&c->d_path, &c->path_cap, c->P));
  CU_TRY(c, ensure(&c->d_tie, &c->tie_cap, c->P));
  CU_TRY(c, ensure(&c->d_score, &c->score_cap, c->S));
  CU_TRY(c, cudaMemsetAsync(c->d_tie, 0, (size_t)c->P, c->stream));
  const size_t per_point = (size_t)N * ((f64 ? sizeof(double) : 4 * sizeof(float)) + (mode == DNBC_DECODE_BACKTRACE ? 2 : 0)) + 8;
  int64_t maxp = chunk_budget_points(c, per_point);
  if (f64) {
    CU_TRY(c, ensure(&c->E64, &c->e64_points, maxp * N));
    maxp = c->e64_points / N;
  } else {
    int rc = reserve_em_scratch(c, std::max(maxp, c->scratch_points));
    if (rc) return rc;
    maxp = c->scratch_points;
  }
  for (const Chunk &ch : make_chunks(c, maxp)) {
    if (f64 && ch.np * N > c->e64_points) CU_TRY(c, ensure(&c->E64, &c->e64_points, ch.np * N));
    if (!f64 && ch.np > c->scratch_points) {
      int rc = reserve_em_scratch(c, ch.np);
      if (rc) return rc;
    }
    if (mode == DNBC_DECODE_BACKTRACE) CU_TRY(c, ensure(&c->d_bp, &c->bp_elems, ch.np * N));
    if (f64) {
      launch_emission_f64(c, ch.p0, ch.np, c->E64);
      launch_decode_f64(c, mode, ch.s0, ch.s1, ch.p0, c->E64, c->d_path, c->d_score, c->d_tie, c->d_bp);
    } else {
      launch_emission_f32(c, ch.p0, ch.np, c->E);
      launch_decode_f32(c, mode, ch.s0, ch.s1, ch.p0, c->E, c->d_path, c->d_score, c->d_tie, c->d_bp);
    }
  }
  cudaStream_t st = c->stream;
  if (path_out) CU_TRY(c, cudaMemcpyAsync(path_out, c->d_path, (size_t)c->P * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
  if (score_out) CU_TRY(c, cudaMemcpyAsync(score_out, c->d_score, (size_t)c->S * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (tie_out) CU_TRY(c, cudaMemcpyAsync(tie_out, c->d_tie, (size_t)c->P, cudaMemcpyDeviceToHost, st));
  CU_TRY(c, cudaStreamSynchronize(st));
  CU_TRY(c, cudaGetLastError());
  return DNBC_OK;
}

// ---------------------------------------------------------------------------------
// streams / timing hooks
// ---------------------------------------------------------------------------------
extern "C" int dnbc_synchronize(dnbc_ctx *c) {
  if (!c) return DNBC_EINVAL;
  CU_TRY(c, cudaSetDevice(c->device));
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  CU_TRY(c, cudaGetLastError());
  return DNBC_OK;
}

extern "C" int dnbc_stream(dnbc_ctx *c, void **stream) {
  if (!c || !stream) return DNBC_EINVAL;
  *stream = (void *)c->stream;
  return DNBC_OK;
}

extern "C" int64_t dnbc_launch_count(const dnbc_ctx *c) { return c ? c->launches : 0; }

extern "C" int dnbc_profile_enable(dnbc_ctx *c, int on) {
  if (!c) return DNBC_EINVAL;
  c->profile = on != 0;
  return DNBC_OK;
}

extern "C" int dnbc_profile_read(dnbc_ctx *c, double *ms_out, int64_t *launches_out, int reset) {
  if (!c) return DNBC_EINVAL;
  if (c->ev_used) {
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    for (size_t q = 0; q < c->ev_used; ++q) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, c->ev_pool[q].a, c->ev_pool[q].b) == cudaSuccess) c->fam_ms[c->ev_pool[q].fam] += ms;
    }
    c->ev_used = 0;
  }
  for (int f = 0; f < FAM_COUNT; ++f) {
    if (ms_out) ms_out[f] = c->fam_ms[f];
    if (launches_out) launches_out[f] = c->fam_launches[f];
    if (reset) { c->fam_ms[f] = 0.0; c->fam_launches[f] = 0; }
  }
  return DNBC_OK;
}
