// k_generate.cu -- synthetic observations drawn on the device from the current parameters.
//
// Follows the sampling loop of the reference's data generator,
// test-utils/src/main/scala/TestUtils.scala:171-185 (hidden state from the initial edge,
// then from the previous state's transition edge; one symbol per discrete variable and one
// mixture draw per continuous variable), with RandomUtils.nextProbabilityIndex
// (core/src/main/scala/RandomUtils.scala:9-18: first index with acc < u <= acc + p) and
// GaussianUtils.nextGaussianMixture / nextGaussian (core/src/main/scala/GaussianUtils.scala:17-35:
// component by weight, then gaussian * sqrt(variance) + mean).  The reference draws from the
// unseeded global scala.util.Random; here a counter-based Philox4x32-10 keyed by
// (seed, sequence) makes every sequence reproducible and independent of the launch shape.
// 10^9 timesteps of the scale-out configuration (80 GB) cannot be generated on the host and
// shipped, which is why this lives next to the kernels.
#include "dnbc_internal.cuh"

struct Philox {
  uint32_t key0, key1;
  uint32_t ctr[4];
  uint32_t out[4];
  int have;
  __device__ Philox(uint64_t seed, uint64_t stream) : have(0) {
    key0 = (uint32_t)seed;
    key1 = (uint32_t)(seed >> 32);
    ctr[0] = 0; ctr[1] = 0;
    ctr[2] = (uint32_t)stream;
    ctr[3] = (uint32_t)(stream >> 32);
  }
  __device__ void round10() {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key0, k1 = key1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    if (++ctr[0] == 0) ++ctr[1];
  }
  __device__ uint32_t next() {
    if (have == 0) { round10(); have = 4; }
    return out[4 - have--];
  }
  // uniform in (0, 1]
  __device__ double uniform() { return ((double)next() + 1.0) * (1.0 / 4294967296.0); }
  __device__ double gaussian() {
    const double u1 = uniform(), u2 = uniform();
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
  }
};

// first index with acc < u <= acc + p (falls to the last positive entry on round-off)
__device__ __forceinline__ int draw_index(const double *__restrict__ p, int n, int stride, double u) {
  double acc = 0.0;
  int last = 0;
  for (int q = 0; q < n; ++q) {
    const double pq = p[(int64_t)q * stride];
    if (pq > 0.0) last = q;
    if (acc < u && u <= acc + pq) return q;
    acc += pq;
  }
  return last;
}

__global__ void __launch_bounds__(128) k_generate(Dims dm, const int32_t *__restrict__ V, const int32_t *__restrict__ K,
                                                  const double *__restrict__ pi, const double *__restrict__ A,
                                                  const double *__restrict__ B, const double *__restrict__ w,
                                                  const double *__restrict__ mu, const double *__restrict__ var,
                                                  int64_t S, int T, uint64_t seed, uint8_t *disc, float *cont,
                                                  int32_t *labels) {
  const int N = dm.N, D = dm.D, C = dm.C, Vmax = dm.Vmax, Kmax = dm.Kmax;
  const int64_t P = S * (int64_t)T;
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < S; s += (int64_t)gridDim.x * blockDim.x) {
    Philox rng(seed, (uint64_t)s);
    int h = 0;
    for (int t = 0; t < T; ++t) {
      const int64_t n = s * (int64_t)T + t;
      h = t == 0 ? draw_index(pi, N, 1, rng.uniform()) : draw_index(A + (int64_t)h * N, N, 1, rng.uniform());
      labels[n] = h;
      for (int d = 0; d < D; ++d)
        disc[(int64_t)d * P + n] = (uint8_t)draw_index(B + ((int64_t)d * N + h) * Vmax, V[d], 1, rng.uniform());
      for (int c = 0; c < C; ++c) {
        const int64_t base = ((int64_t)c * N + h) * Kmax;
        const int k = draw_index(w + base, K[c], 1, rng.uniform());
        cont[(int64_t)c * P + n] = (float)(rng.gaussian() * sqrt(var[base + k]) + mu[base + k]);
      }
    }
  }
}

void launch_generate(dnbc_ctx *c, int64_t n_seq, int32_t seq_len, uint64_t seed) {
  if (n_seq <= 0) return;
  c->launches++;
  const int grid = (int)std::min<int64_t>((n_seq + 127) / 128, (int64_t)c->sm_count * 16);
  k_generate<<<grid, 128, 0, c->stream>>>(c->dm, c->d_V, c->d_K, c->d_pi, c->d_A, c->d_B, c->d_w, c->d_mu, c->d_var,
                                          n_seq, seq_len, seed, c->d_disc, c->d_cont, c->d_labels);
}
