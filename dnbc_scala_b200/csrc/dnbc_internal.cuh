// dnbc_internal.cuh -- context, device tables and launch helpers shared by the kernels.
// sm_100a only; no CPU fallback anywhere in this directory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/dnbc_b200.h"

#define DNBC_LOG2E 1.4426950408889634074
#define DNBC_LN2 0.69314718055994530942
#define DNBC_PI 3.14159265358979323846

// kernel families for the profile counters (order fixed by dnbc_profile_read)
enum { FAM_EMISSION = 0, FAM_FORWARD, FAM_BACKWARD, FAM_XI, FAM_STATS_GMM, FAM_STATS_CAT, FAM_MSTEP, FAM_DECODE, FAM_COUNT };

// Read-only fp32 tables derived from the fp64 master parameters (rebuilt by
// k_prepare_tables after every set_params / M-step).
struct DeviceTables {
  const float *pi;      // [N]        probabilities
  const float *A;       // [N][Npad]  row = from-state, zero padded
  const float *At;      // [N][Npad]  transposed (row = to-state)
  const float *logpi;   // [N]        natural log
  const float *logA;    // [N][Npad]  natural log, row = from-state
  const float *cat;     // [D][Vmax+1][Npad] log2 P(v | j); row Vmax = unseen symbol = -inf
  const float *gs;      // [C][Kmax][Npad]   z = x*gs + gm
  const float *gm;      // [C][Kmax][Npad]
  const float *gc;      // [C][Kmax][Npad]   log2 w - 0.5 log2(2 pi var);  arg = gc - z*z
  const int32_t *K;     // [C]
  const uint8_t *active;// [N]
};

struct Dims {
  int N, D, C, Vmax, Kmax, Npad;
};

struct DataView {
  int64_t P;            // total points resident
  const int64_t *off;   // [S+1] device
  const uint8_t *disc;  // [D][P]
  const float *cont;    // [C][P]
  const double *cont64; // [C][P] or nullptr
  const int32_t *labels;// [P] or nullptr
};

struct dnbc_ctx {
  Dims dm{};
  std::vector<int32_t> V, K;
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;        // host->device copies overlapped with the E-step
  std::vector<cudaEvent_t> chunk_events;
  std::string err;
  int64_t launches = 0;

  // fp64 master parameters (device)
  double *d_pi = nullptr, *d_A = nullptr, *d_B = nullptr, *d_w = nullptr, *d_mu = nullptr, *d_var = nullptr;
  uint8_t *d_active = nullptr;
  int32_t *d_V = nullptr, *d_K = nullptr;
  // fp32 derived tables
  float *t_pi = nullptr, *t_A = nullptr, *t_At = nullptr, *t_logpi = nullptr, *t_logA = nullptr, *t_cat = nullptr;
  float *t_gs = nullptr, *t_gm = nullptr, *t_gc = nullptr;
  // fp64 decode tables, logs taken on the host with glibc (bit-compatible with a JVM-style libm oracle)
  double *t64_logpi = nullptr, *t64_logA = nullptr, *t64_logB = nullptr;  // logB [D][Vmax+1][N]
  double *t64_u = nullptr, *t64_rs = nullptr;                              // [C][N][Kmax]
  bool tables64_valid = false;

  // observations
  int64_t S = 0, P = 0, Tmax = 0;
  std::vector<int64_t> h_off;
  int64_t *d_off = nullptr;
  uint8_t *d_disc = nullptr;
  float *d_cont = nullptr;
  double *d_cont64 = nullptr;
  int32_t *d_labels = nullptr;
  int64_t cap_points = 0, cap_seqs = 0;
  bool have_labels = false, have_cont64 = false;

  // scratch for the EM pipeline, sized for one chunk of sequences
  float *E = nullptr, *AL = nullptr, *GA = nullptr, *U = nullptr, *cs = nullptr;
  int64_t scratch_points = 0;
  double *d_stats = nullptr;
  int64_t stats_size = 0;
  // M-step variance floor (relative to the pooled variance of the observed variable) and the
  // number of components it raised since the context was created
  double var_floor_rel = 1e-10;
  double *d_pooled = nullptr;             // [C]
  unsigned long long *d_floored = nullptr;
  double *h_pinned = nullptr;   // small pinned bounce buffer
  size_t h_pinned_bytes = 0;

  // tensor-core recursion (Npad >= 256): bf16 split tables, operand staging, row maxima
  void *tc_tab = nullptr, *tc_x = nullptr;
  float *tc_mrow = nullptr;
  float *vt_logA = nullptr;          // [Npad][Npad] padded ln A for k_viterbi_tile.cu
  int64_t tc_mrow_cap = 0;

  // decode scratch
  double *E64 = nullptr;
  int64_t e64_points = 0;
  uint16_t *d_bp = nullptr;
  int64_t bp_elems = 0;
  uint16_t *d_path = nullptr;
  double *d_score = nullptr;
  uint8_t *d_tie = nullptr;
  int64_t path_cap = 0, score_cap = 0, tie_cap = 0;

  // supervised learner scratch
  unsigned long long *d_cnt = nullptr;   // [N | N*N | D*N*Vmax]
  int32_t *d_first_seq = nullptr;        // [N]
  uint8_t *d_assign = nullptr;           // [C][P] k-means assignments
  int64_t assign_cap = 0;
  double *d_edge = nullptr;              // per-edge fp64 workspace (layout in k_supervised.cu)
  int32_t *d_edge_i = nullptr;           // per-edge int workspace

  // dnbc_set_option
  int64_t opt_max_chunk_points = 0;   // 0 = sized from the free memory
  bool opt_fb2_interleave = false;    // keep the four-sequences-per-lane-group recursions on small inputs
  int opt_fused_estep = 0;            // 0 auto, 1 always (when it fits), 2 never
  int opt_fb_seq_lanes = 0;           // lane-per-sequence recursions (k_fbs.cu): 0 auto, 1 always (when N allows), 2 never

  // library-owned NCCL communicator (dnbc_comm.cu); void* keeps nccl.h out of the kernel files
  void *comm = nullptr;
  int comm_ranks = 1, comm_rank = 0;
  int64_t collectives = 0;

  // profiling
  // (event pairs are only recorded, never waited on, while the stream runs; they are
  // resolved in dnbc_profile_read so the timed region is not perturbed)
  bool profile = false;
  double fam_ms[FAM_COUNT] = {0};
  int64_t fam_launches[FAM_COUNT] = {0};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  struct EvPair { cudaEvent_t a, b; int fam; };
  std::vector<EvPair> ev_pool;
  size_t ev_used = 0;

  DeviceTables tables() const {
    return DeviceTables{t_pi, t_A, t_At, t_logpi, t_logA, t_cat, t_gs, t_gm, t_gc, d_K, d_active};
  }
  // the fp64 planes are only valid for the upload that filled them (a later f32 upload / generate leaves them stale)
  DataView data() const {
    return DataView{P, d_off, d_disc, d_cont, have_cont64 ? d_cont64 : nullptr, have_labels ? d_labels : nullptr};
  }
};

// statistics buffer offsets (doubles)
struct StatsLayout {
  int64_t ll, pi, xi, gamma, cat, gmm, total;
};
static inline StatsLayout stats_layout(const Dims &d) {
  StatsLayout s;
  s.ll = 0;
  s.pi = 1;
  s.xi = s.pi + d.N;
  s.gamma = s.xi + (int64_t)d.N * d.N;
  s.cat = s.gamma + d.N;
  s.gmm = s.cat + (int64_t)d.D * d.N * d.Vmax;
  s.total = s.gmm + 3 * (int64_t)d.C * d.N * d.Kmax;
  return s;
}

// ---- kernel launchers (each defined next to its kernels) -----------------------
// [s0, s1) = sequence range of the chunk, p0 = first point of the chunk; scratch
// arrays are indexed by (point - p0).
void launch_prepare_tables(dnbc_ctx *c);
void launch_emission_f32(dnbc_ctx *c, int64_t p0, int64_t np, float *out);
void launch_emission_f64(dnbc_ctx *c, int64_t p0, int64_t np, double *out);
void launch_clamp_labels(dnbc_ctx *c, int64_t p0, int64_t np, float *E);
void launch_forward(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0);
void launch_backward(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0);
void launch_xi(dnbc_ctx *c, int64_t np);
void launch_stats_gmm(dnbc_ctx *c, int64_t p0, int64_t np);
void launch_stats_cat(dnbc_ctx *c, int64_t p0, int64_t np);
void launch_mstep(dnbc_ctx *c, int64_t total_sequences);
void launch_decode_f32(dnbc_ctx *c, int mode, int64_t s0, int64_t s1, int64_t p0, const float *E,
                       uint16_t *path, double *score, uint8_t *tie, void *bp);
void launch_decode_f64(dnbc_ctx *c, int mode, int64_t s0, int64_t s1, int64_t p0, const double *E,
                       uint16_t *path, double *score, uint8_t *tie, void *bp);
bool mixture_emission_fast(const dnbc_ctx *c);
bool mixture_stats_fast(const dnbc_ctx *c);
void launch_mixture_emission(dnbc_ctx *c, int64_t p0, int64_t np, float *out);
void launch_mixture_stats(dnbc_ctx *c, int64_t p0, int64_t np);
bool stats2_supported(const dnbc_ctx *c);
bool fused_estep_supported(const dnbc_ctx *c);
bool fused_estep_wanted(const dnbc_ctx *c, int64_t n_seq);
void launch_estep_fused(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0);
int64_t launch_stats2(dnbc_ctx *c, int64_t p0, int64_t np);
bool fb3_supported(const dnbc_ctx *c);
void launch_fb3(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0, bool backward);
bool fb2_supported(const dnbc_ctx *c);
bool fb2_fuses_xi(const dnbc_ctx *c);
void launch_fb2(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0, bool backward);
bool fbs_supported(const dnbc_ctx *c, int64_t nseq);
void launch_fbs(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0, bool backward);
bool viterbi_tile_supported(const dnbc_ctx *c);
int launch_viterbi_tile(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0, const float *E, uint16_t *path, double *score,
                        uint8_t *tie, uint16_t *bp);
void launch_backtrace(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0, const uint16_t *bp, uint16_t *path);
bool xi_tc_supported(const dnbc_ctx *c);
void launch_xi_tc(dnbc_ctx *c, int64_t np);
bool fb_tc_supported(const dnbc_ctx *c);
int64_t fb_tc_wave_sequences(const dnbc_ctx *c);
void launch_tc_tables(dnbc_ctx *c);
int launch_fb_tc(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0, int64_t np, bool backward);
void launch_fb_block(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0, bool backward);
void launch_decode_block_f32(dnbc_ctx *c, int mode, int64_t s0, int64_t s1, int64_t p0, const float *E, uint16_t *path,
                             double *score, uint8_t *tie, void *bp);
void launch_decode_block_f64(dnbc_ctx *c, int mode, int64_t s0, int64_t s1, int64_t p0, const double *E, uint16_t *path,
                             double *score, uint8_t *tie, void *bp);
void launch_supervised_counts(dnbc_ctx *c, int flags, unsigned long long *cnt_pi, unsigned long long *cnt_A,
                              unsigned long long *cnt_B);
void launch_normalize_counts(dnbc_ctx *c, const unsigned long long *cnt_pi, const unsigned long long *cnt_A,
                             const unsigned long long *cnt_B);
// per-edge k-means + EM1D on the device; returns a DNBC_* status
int run_gmm_mle(dnbc_ctx *c, int flags, const double *h_init_centroids, int max_iters, double rel_tol,
                double *h_ll_series, int32_t *h_n_iters);
void launch_generate(dnbc_ctx *c, int64_t n_seq, int32_t seq_len, uint64_t seed);
void launch_f64_to_f32(dnbc_ctx *c, const double *src, float *dst, int64_t n);
void launch_f32_to_f64(dnbc_ctx *c, const float *src, double *dst, int64_t n);

// error plumbing shared by the translation units
int dnbc_fail(dnbc_ctx *c, int code, const char *fmt, ...);
#define CU_TRY(c, expr)                                                                          \
  do {                                                                                           \
    cudaError_t e__ = (expr);                                                                    \
    if (e__ != cudaSuccess) return dnbc_fail((c), DNBC_ECUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

// profiling bracket around a launch family
struct FamTimer {
  dnbc_ctx *c;
  int fam;
  dnbc_ctx::EvPair *ev = nullptr;
  FamTimer(dnbc_ctx *ctx, int f) : c(ctx), fam(f) {
    c->launches++;
    c->fam_launches[fam]++;
    if (c->profile) {
      if (c->ev_used == c->ev_pool.size()) {
        dnbc_ctx::EvPair p{nullptr, nullptr, 0};
        cudaEventCreate(&p.a);
        cudaEventCreate(&p.b);
        c->ev_pool.push_back(p);
      }
      c->ev_pool[c->ev_used].fam = fam;
      cudaEventRecord(c->ev_pool[c->ev_used].a, c->stream);
    }
  }
  ~FamTimer() {
    if (c->profile) {
      cudaEventRecord(c->ev_pool[c->ev_used].b, c->stream);
      c->ev_used++;
    }
  }
};

static inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// ---- device helpers -------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
