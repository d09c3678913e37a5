// k_fbs.cu -- K2 for small state spaces and MANY sequences: one LANE per sequence.
//
// The reference has no forward-backward; these kernels implement SURVEY Appendix B like k_fb2.cu and
// are checked against the fp64 oracle (oracle/dnbc_oracle.c fb_one) by the same parity tests.
//
// Why a second mapping: k_fb2.cu gives a sequence a group of N lanes (lane = state).  At N = 10 a lane then does
// 10 multiply-adds per step and ~60 instructions of exchange, shuffle reductions, address arithmetic and stores:
// ncu (profiles/r02t_ncu_n10_summary.txt) counts 26 warp instructions per point in the forward and 50 in the backward
// sweep at 12 points per warp step.  Here a lane owns a whole sequence: the N-vector lives in its registers, a step is
// N^2 / 2 packed FMAs against transition-matrix pairs that every lane reads from the same shared-memory address
// (broadcast), the normaliser is a private sum, and a warp step advances 32 sequences: 12 (forward) / 21 (backward)
// warp instructions per point, no shuffles, no barriers inside a tile, straight-line steps.
// What makes it work is the staging: a lane's rows are 4N bytes apiece and 32 lanes' rows are 32 different cache
// lines (one L1 tag lookup each), so the E / alpha^ / gamma rows of 4 steps x 32 sequences travel through shared
// memory in row-major tiles, copied in with 8-byte cp.async and out with 8-byte stores whose lanes run along a
// sequence's contiguous 16N bytes; outputs overwrite their inputs in the tile (alpha^_t over E_t, gamma_t over
// alpha^_t); tile rows are 2 floats longer than 4N, so the lane-private 64-bit accesses (stride 2N + 1 eight-byte
// units) are bank-conflict free.
// Used for even N <= 12 when the chunk holds at least a wave of 32-sequence warps; the transition statistics are
// accumulated inside the backward sweep (N^2 fp32 registers per lane, flushed in fp64 every ~4096 points per lane).
// Measured at N = 10, 100,000 x T = 200 (profiles/r02d_*): forward 0.81 -> 0.56 ms, backward 1.23 -> 1.11 ms; both are
// now bound by shared-memory latency (ncu: LSU pipe 53 % / 36 %, stalls short_scoreboard) at 12 / 7 warps per SM -- the
// backward sweep's N^2 accumulators cost it the occupancy.  With the transition matrix in registers the forward sweep
// needs 166 registers and is slower (0.84 ms).  Below ~19,000 sequences per GPU the lane-per-state kernels win.
#include "dnbc_internal.cuh"
#include "packed_f32.cuh"

namespace {

constexpr int FBS_WARPS = 4;
constexpr int TB = 4;                                        // steps per staged tile (2 measured slower: 0.60 / 1.17 ms at N = 10)


struct FbsArgs {
  Dims dm;
  DeviceTables tb;
  const int64_t *off;
  int64_t s0, s1, p0;
  const float *E;   // [np][N] natural-log emissions (chunk-relative rows)
  float *AL;        // [np][N] scaled alpha
  float *GA;        // [np][N] gamma
  float *cs;        // [np]    c_t
  double *stats;    // packed statistics ([0] = log-likelihood)
  int64_t st_pi, st_xi;
};

// predicated copies (no branch around them: the tile copies are straight-line code)
__device__ __forceinline__ void cp_async8_if(bool on, void *smem, const void *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.ca.shared.global [%0], [%1], 8;\n\t}" ::"r"(sa), "l"(gmem), "r"((int)on) : "memory");
}
__device__ __forceinline__ void cp_async4_if(bool on, void *smem, const void *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.ca.shared.global [%0], [%1], 4;\n\t}" ::"r"(sa), "l"(gmem), "r"((int)on) : "memory");
}
__device__ __forceinline__ void st_global2_if(bool on, float *dst, float2 v) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\t@p st.global.v2.f32 [%0], {%1, %2};\n\t}" ::"l"(dst), "f"(v.x), "f"(v.y), "r"((int)on) : "memory");
}
__device__ __forceinline__ void st_global1_if(bool on, float *dst, float v) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p st.global.f32 [%0], %1;\n\t}" ::"l"(dst), "f"(v), "r"((int)on) : "memory");
}
// matrix rows are re-read from shared memory at every step (a broadcast LDS.128 each): as ordinary loads the compiler
// hoists all N^2 values into registers, which costs the occupancy the sweep lives on
__device__ __forceinline__ float4 lds128_volatile(const float *p) {
  float4 v;
  const unsigned sa = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sa));
  return v;
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int G>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(G) : "memory"); }
__device__ __forceinline__ u64 ex2_2(u64 a) {
  u64 d;
  asm("{\n\t.reg .f32 lo, hi;\n\tmov.b64 {lo, hi}, %1;\n\tex2.approx.ftz.f32 lo, lo;\n\tex2.approx.ftz.f32 hi, hi;\n\t"
      "mov.b64 %0, {lo, hi};\n\t}" : "=l"(d) : "l"(a));
  return d;
}

template <int N>
struct Geo {
  static constexpr int ROW = TB * N + 2;                     // floats per tile row (one sequence, TB steps)
  static constexpr int PC = TB * N / 2;                      // 8-byte pieces per tile row
  static constexpr int NP2 = N / 2;
  static constexpr int NR = (N + 3) / 4 * 4;                 // matrix row length in shared memory (16-byte rows)
  static constexpr int CROW = TB + 1;                        // floats per row of a normaliser tile
};

// per-warp directory of the 32 sequences of a block
struct SeqDir {
  int64_t row[32];                                           // first element of the sequence's rows (chunk-relative)
  int64_t pt[32];                                            // first point (chunk-relative)
  int T[32];
};

// Tile copies.  An instruction moves SPI whole tile rows (one sequence each, 16N contiguous bytes): lane l owns the
// 8-byte piece l % PC of row l / PC, so the piece index, its step and the lane's activity are per-lane constants and
// a copy costs a directory lookup, one compare and the address arithmetic (the first version derived sequence and
// piece from a running index: ~28 instructions per copy instruction, more than the recursion itself).
template <int N>
struct CopyLane {
  static constexpr int PC = Geo<N>::PC, SPI = 32 / PC > 0 ? 32 / PC : 1;
  static_assert(PC <= 32 && 32 % SPI == 0, "a copy instruction moves whole tile rows and the instructions tile the 32 rows");
  int q0, f, st;                                             // row within an instruction, float offset, step of the piece
  bool on;
  __device__ __forceinline__ CopyLane(int lane) : q0(lane / PC), f(2 * (lane % PC)), st(2 * (lane % PC) / N), on(lane < SPI * PC) {}
};
// rows [t0, t0 + TB) of the warp's 32 sequences -> tile (row-major, ROW floats per sequence); rows at or past a
// sequence's end (or before its start) are not copied
template <int N>
__device__ __forceinline__ void tile_in(float *tile, const float *src, const SeqDir &d, int t0, const CopyLane<N> &cl) {
  constexpr int ROW = Geo<N>::ROW, SPI = CopyLane<N>::SPI;
  const int t = t0 + cl.st;
  const bool on = cl.on && t >= 0;
  float *dst = tile + cl.q0 * ROW + cl.f;
  src += (int64_t)t0 * N + cl.f;
#pragma unroll 8
  for (int m = 0; m < 32 / SPI; ++m) {
    const int q = m * SPI + cl.q0;
    cp_async8_if(on && t < d.T[q], dst + m * SPI * ROW, src + d.row[q]);
  }
}
template <int N>
__device__ __forceinline__ void tile_out(const float *tile, float *dst, const SeqDir &d, int t0, const CopyLane<N> &cl) {
  constexpr int ROW = Geo<N>::ROW, SPI = CopyLane<N>::SPI;
  const int t = t0 + cl.st;
  const float *srcp = tile + cl.q0 * ROW + cl.f;
  dst += (int64_t)t0 * N + cl.f;
#pragma unroll 8
  for (int m = 0; m < 32 / SPI; ++m) {
    const int q = m * SPI + cl.q0;
    st_global2_if(cl.on && t < d.T[q], dst + d.row[q], *reinterpret_cast<const float2 *>(srcp + m * SPI * ROW));
  }
}

// shared memory: matrix rows [N][NR], pi [NR], then per warp {SeqDir, tiles}
template <int N>
__host__ __device__ constexpr size_t fbs_table_bytes() { return (size_t)(N + 1) * Geo<N>::NR * sizeof(float); }
template <int N>
__host__ __device__ constexpr size_t fbs_fwd_warp_bytes() {
  return sizeof(SeqDir) + (size_t)(2 * 32 * Geo<N>::ROW + 32 * Geo<N>::CROW) * sizeof(float);
}
template <int N>
__host__ __device__ constexpr size_t fbs_bwd_warp_bytes() {
  return sizeof(SeqDir) + (size_t)(4 * 32 * Geo<N>::ROW + 2 * 32 * Geo<N>::CROW) * sizeof(float);
}

// ---- forward: alpha^_t(j) = b~_t(j) sum_i alpha^_{t-1}(i) a_ij / c_t ;  LL += ln c_t + m_t
template <int N>
__global__ void __launch_bounds__(FBS_WARPS * 32, 4) k_forward_s(FbsArgs a) {
  using G = Geo<N>;
  constexpr int ROW = G::ROW, NP2 = G::NP2, NR = G::NR, CROW = G::CROW;
  extern __shared__ __align__(16) unsigned char smraw[];
  float *sA = reinterpret_cast<float *>(smraw);             // [N][NR]: row i = a_i0 .. a_i,N-1
  float *sPi = sA + N * NR;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned char *wb = smraw + fbs_table_bytes<N>() + (size_t)w * fbs_fwd_warp_bytes<N>();
  SeqDir &dir = *reinterpret_cast<SeqDir *>(wb);
  float *Ein = reinterpret_cast<float *>(wb + sizeof(SeqDir));   // [2][32][ROW]; alpha^_t overwrites E_t in place
  float *Cout = Ein + 2 * 32 * ROW;                              // [32][CROW]
  const CopyLane<N> cl(lane);
  for (int t = threadIdx.x; t < (N + 1) * NR; t += blockDim.x) {
    const int i = t / NR, j = t % NR;
    sA[t] = j < N ? (i < N ? a.tb.A[i * a.dm.Npad + j] : a.tb.pi[j]) : 0.f;
  }
  __syncthreads();
  u64 pi2[NP2];
#pragma unroll
  for (int jj = 0; jj < NP2; ++jj) pi2[jj] = pk(sPi[2 * jj], sPi[2 * jj + 1]);
  const u64 l2e = pk((float)DNBC_LOG2E, (float)DNBC_LOG2E);
  double ll = 0.0;
  const int nw = blockDim.x >> 5;
  const int64_t sstride = (int64_t)gridDim.x * nw * 32;
  for (int64_t sb = a.s0 + ((int64_t)blockIdx.x * nw + w) * 32; sb < a.s1; sb += sstride) {
    const int64_t s = sb + lane;
    const int T = s < a.s1 ? (int)(a.off[s + 1] - a.off[s]) : 0;
    const int64_t pt = s < a.s1 ? a.off[s] - a.p0 : 0;
    __syncwarp();                                            // the previous block's copies out have read the directory
    dir.row[lane] = pt * N;
    dir.pt[lane] = pt;
    dir.T[lane] = T;
    int Tw = T;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) Tw = max(Tw, __shfl_xor_sync(0xffffffffu, Tw, o));
    __syncwarp();
    tile_in<N>(Ein, a.E, dir, 0, cl);
    cp_commit();
    u64 al2[NP2];
#pragma unroll
    for (int jj = 0; jj < NP2; ++jj) al2[jj] = 0ull;
    for (int tb = 0, b = 0; tb < Tw; tb += TB, b ^= 1) {
      if (tb + TB < Tw) {
        tile_in<N>(Ein + (b ^ 1) * 32 * ROW, a.E, dir, tb + TB, cl);
        cp_commit();
        cp_wait<1>();
      } else {
        cp_wait<0>();
      }
      __syncwarp();                                          // tile b is complete; tile b ^ 1 and Cout were copied out
      float *er = Ein + b * 32 * ROW + lane * ROW;
      float lt = 0.f;
#pragma unroll
      for (int st = 0; st < TB; ++st) {
        // straight-line code: a lane past its sequence's end computes on stale tile contents, which nothing reads
        // (its rows are not copied out, its log-likelihood term is dropped, the next block starts from pi)
        const int t = tb + st;
        u64 acc[NP2];
        if (t == 0) {                                        // (warp-uniform)
#pragma unroll
          for (int jj = 0; jj < NP2; ++jj) acc[jj] = pi2[jj];
        } else {
#pragma unroll
          for (int jj = 0; jj < NP2; ++jj) acc[jj] = 0ull;
#pragma unroll
          for (int i = 0; i < N; ++i) {
            float lo, hi;
            unpk(al2[i / 2], lo, hi);
            const float ai = (i & 1) ? hi : lo;
            const u64 ai2 = pk(ai, ai);
#pragma unroll
            for (int j4 = 0; j4 < NR / 4; ++j4) {
              const float4 v = lds128_volatile(sA + i * NR + 4 * j4);
              if (2 * j4 < NP2) acc[2 * j4] = fma2(ai2, pk(v.x, v.y), acc[2 * j4]);
              if (2 * j4 + 1 < NP2) acc[2 * j4 + 1] = fma2(ai2, pk(v.z, v.w), acc[2 * j4 + 1]);
            }
          }
        }
        // emission factor exp(E_t - max_j E_t)
        u64 e2[NP2];
        float m = -INFINITY;
#pragma unroll
        for (int jj = 0; jj < NP2; ++jj) {
          const float2 v = *reinterpret_cast<const float2 *>(er + st * N + 2 * jj);
          e2[jj] = pk(v.x, v.y);
          m = fmaxf(m, fmaxf(v.x, v.y));
        }
        m = m == -INFINITY ? 0.f : m;
        const u64 nm = pk(-m, -m);
        u64 sum2 = 0ull;
#pragma unroll
        for (int jj = 0; jj < NP2; ++jj) {
          acc[jj] = mul2(acc[jj], ex2_2(mul2(add2(e2[jj], nm), l2e)));
          sum2 = jj == 0 ? acc[0] : add2(sum2, acc[jj]);
        }
        float slo, shi;
        unpk(sum2, slo, shi);
        const float c = slo + shi;
        const float inv = rcp_approx(c);
        const u64 inv2 = pk(inv, inv);
        float *ao = er + st * N;
#pragma unroll
        for (int jj = 0; jj < NP2; ++jj) {
          al2[jj] = mul2(acc[jj], inv2);
          float lo, hi;
          unpk(al2[jj], lo, hi);
          *reinterpret_cast<float2 *>(ao + 2 * jj) = make_float2(lo, hi);
        }
        Cout[lane * CROW + st] = c;
        const float term = lg2_approx(c) * (float)DNBC_LN2 + m;
        lt += t < T ? term : 0.f;
      }
      ll += (double)lt;
      __syncwarp();
      tile_out<N>(Ein + b * 32 * ROW, a.AL, dir, tb, cl);
#pragma unroll
      for (int m = 0; m < TB; ++m) {
        const int idx = m * 32 + lane, q = idx / TB, st = idx % TB;
        st_global1_if(tb + st < dir.T[q], a.cs + dir.pt[q] + tb + st, Cout[q * CROW + st]);
      }
      __syncwarp();                                          // the next request overwrites this tile
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ll += __shfl_xor_sync(0xffffffffu, ll, o);
  if (lane == 0 && ll != 0.0) atomicAdd(a.stats, ll);
}

// ---- backward: beta^_t(i) = sum_j a_ij u_t(j),  u_t(j) = b~_{t+1}(j) beta^_{t+1}(j) / c_{t+1};
//      gamma_t = alpha^_t beta^_t;  pi-stat += gamma_0;  xi-stat[i][j] += a_ij sum_t alpha^_t(i) u_t(j)
template <int N>
__global__ void __launch_bounds__(FBS_WARPS * 32) k_backward_s(FbsArgs a) {
  using G = Geo<N>;
  constexpr int ROW = G::ROW, NP2 = G::NP2, NR = G::NR, CROW = G::CROW;
  extern __shared__ __align__(16) unsigned char smraw[];
  float *sAt = reinterpret_cast<float *>(smraw);            // [N][NR]: row j = a_0j .. a_N-1,j (transposed)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned char *wb = smraw + fbs_table_bytes<N>() + (size_t)w * fbs_bwd_warp_bytes<N>();
  SeqDir &dir = *reinterpret_cast<SeqDir *>(wb);
  float *ALin = reinterpret_cast<float *>(wb + sizeof(SeqDir));   // [2][32][ROW]; gamma_t overwrites alpha^_t in place
  float *Ein = ALin + 2 * 32 * ROW;                               // [2][32][ROW]  rows t + 1
  float *Cin = Ein + 2 * 32 * ROW;                                // [2][32][CROW]  c_{t+1}
  const CopyLane<N> cl(lane);
  float *xs = ALin;                                               // flush scratch [(N + 1) N][33] (no copies in flight then)
  static_assert((size_t)(N + 1) * N * 33 <= (size_t)4 * 32 * ROW, "flush scratch must fit the input tiles");
  for (int t = threadIdx.x; t < N * NR; t += blockDim.x) {
    const int j = t / NR, i = t % NR;
    sAt[t] = i < N ? a.tb.A[i * a.dm.Npad + j] : 0.f;
  }
  __syncthreads();
  const u64 l2e = pk((float)DNBC_LOG2E, (float)DNBC_LOG2E);
  u64 xacc[N][NP2];
  float piacc[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    piacc[i] = 0.f;
#pragma unroll
    for (int jj = 0; jj < NP2; ++jj) xacc[i][jj] = 0ull;
  }
  int xcount = 0;
  auto flush = [&]() {                                      // per-lane fp32 windows -> fp64 statistics
    __syncwarp();
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
      for (int jj = 0; jj < NP2; ++jj) {
        float lo, hi;
        unpk(xacc[i][jj], lo, hi);
        xs[(i * N + 2 * jj) * 33 + lane] = lo;
        xs[(i * N + 2 * jj + 1) * 33 + lane] = hi;
        xacc[i][jj] = 0ull;
      }
      xs[(N * N + i) * 33 + lane] = piacc[i];
      piacc[i] = 0.f;
    }
    __syncwarp();
    for (int e = lane; e < (N + 1) * N; e += 32) {
      double sum = 0.0;
      for (int k = 0; k < 32; ++k) sum += (double)xs[e * 33 + k];
      if (sum != 0.0) {
        if (e < N * N) atomicAdd(a.stats + a.st_xi + e, sum * (double)sAt[(e % N) * NR + e / N]);
        else atomicAdd(a.stats + a.st_pi + (e - N * N), sum);
      }
    }
    __syncwarp();
    xcount = 0;
  };
  const int nw = blockDim.x >> 5;
  const int64_t sstride = (int64_t)gridDim.x * nw * 32;
  for (int64_t sb = a.s0 + ((int64_t)blockIdx.x * nw + w) * 32; sb < a.s1; sb += sstride) {
    const int64_t s = sb + lane;
    const int T = s < a.s1 ? (int)(a.off[s + 1] - a.off[s]) : 0;
    const int64_t pt = s < a.s1 ? a.off[s] - a.p0 : 0;
    __syncwarp();
    dir.row[lane] = pt * N;
    dir.pt[lane] = pt;
    dir.T[lane] = T;
    int Tw = T;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) Tw = max(Tw, __shfl_xor_sync(0xffffffffu, Tw, o));
    __syncwarp();
    if (Tw == 0) continue;
    // tiles from the top: rows [tb, tb + TB) of alpha^, rows [tb + 1, tb + TB + 1) of E and c
    auto request = [&](int tb, int b) {
      tile_in<N>(ALin + b * 32 * ROW, a.AL, dir, tb, cl);
      tile_in<N>(Ein + b * 32 * ROW, a.E, dir, tb + 1, cl);
#pragma unroll
      for (int m = 0; m < TB; ++m) {
        const int idx = m * 32 + lane, q = idx / TB, st = idx % TB;
        cp_async4_if(tb + 1 + st < dir.T[q], Cin + b * 32 * CROW + q * CROW + st, a.cs + dir.pt[q] + tb + 1 + st);
      }
      cp_commit();
    };
    const int tbtop = (Tw - 1) / TB * TB;
    request(tbtop, 0);
    u64 be2[NP2];                                            // beta^_{t+1}
#pragma unroll
    for (int jj = 0; jj < NP2; ++jj) be2[jj] = 0ull;
    for (int tb = tbtop, b = 0; tb >= 0; tb -= TB, b ^= 1) {
      if (tb - TB >= 0) {
        request(tb - TB, b ^ 1);
        cp_wait<1>();
      } else {
        cp_wait<0>();
      }
      __syncwarp();
      float *alr = ALin + b * 32 * ROW + lane * ROW;
      const float *er = Ein + b * 32 * ROW + lane * ROW;
      const float *cr = Cin + b * 32 * CROW + lane * CROW;
      float *go = alr;
#pragma unroll 1
      for (int st = TB - 1; st >= 0; --st) {
        // straight-line code; a lane outside its sequence (t >= T) runs on zeros and writes zeros to tile rows
        // that are not copied out
        const int t = tb + st;
        const bool live = t < T, hasu = t < T - 1;
        u64 al2[NP2];
#pragma unroll
        for (int jj = 0; jj < NP2; ++jj) {
          const float2 v = *reinterpret_cast<const float2 *>(alr + st * N + 2 * jj);
          al2[jj] = live ? pk(v.x, v.y) : 0ull;
        }
        // u_j = exp(E_{t+1}(j) - max) beta^_{t+1}(j) / c_{t+1}   (none at the sequence's last step)
        u64 e2[NP2], u2[NP2], bn2[NP2];
        float m = -INFINITY;
#pragma unroll
        for (int jj = 0; jj < NP2; ++jj) {
          const float2 v = *reinterpret_cast<const float2 *>(er + st * N + 2 * jj);
          e2[jj] = pk(v.x, v.y);
          m = fmaxf(m, fmaxf(v.x, v.y));
        }
        m = m == -INFINITY ? 0.f : m;
        const u64 nm = pk(-m, -m);
        const float ic = rcp_approx(cr[st]);
        const u64 ic2 = pk(ic, ic);
#pragma unroll
        for (int jj = 0; jj < NP2; ++jj) {
          const u64 uu = mul2(mul2(ex2_2(mul2(add2(e2[jj], nm), l2e)), ic2), be2[jj]);
          u2[jj] = hasu ? uu : 0ull;
          bn2[jj] = 0ull;
        }
        // beta^_t(i) = sum_j a_ij u_j (pairs over i from the transposed rows); xi += alpha^_t(i) u_j
#pragma unroll
        for (int j = 0; j < N; ++j) {
          float lo, hi;
          unpk(u2[j / 2], lo, hi);
          const float uj = (j & 1) ? hi : lo;
          const u64 uj2 = pk(uj, uj);
#pragma unroll
          for (int i4 = 0; i4 < NR / 4; ++i4) {
            const float4 v = lds128_volatile(sAt + j * NR + 4 * i4);
            if (2 * i4 < NP2) bn2[2 * i4] = fma2(uj2, pk(v.x, v.y), bn2[2 * i4]);
            if (2 * i4 + 1 < NP2) bn2[2 * i4 + 1] = fma2(uj2, pk(v.z, v.w), bn2[2 * i4 + 1]);
          }
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
          float lo, hi;
          unpk(al2[i / 2], lo, hi);
          const float ai = (i & 1) ? hi : lo;
          const u64 ai2 = pk(ai, ai);
#pragma unroll
          for (int jj = 0; jj < NP2; ++jj) xacc[i][jj] = fma2(ai2, u2[jj], xacc[i][jj]);
        }
        if (t == T - 1) {
#pragma unroll
          for (int jj = 0; jj < NP2; ++jj) bn2[jj] = pk(1.f, 1.f);
        }
        // sum_i alpha^_t(i) beta^_t(i) is 1 in exact arithmetic; rescale beta^_t so that it is (k_fb2.cu)
        u64 g2 = mul2(al2[0], bn2[0]);
#pragma unroll
        for (int jj = 1; jj < NP2; ++jj) g2 = fma2(al2[jj], bn2[jj], g2);
        float glo, ghi;
        unpk(g2, glo, ghi);
        const float gs = glo + ghi;
        const float fix = gs > 0.f ? rcp_approx(gs) : 1.f;
        const u64 fix2 = pk(fix, fix);
        const bool first = t == 0 && live;
#pragma unroll
        for (int jj = 0; jj < NP2; ++jj) {
          be2[jj] = mul2(bn2[jj], fix2);
          const u64 gg = mul2(al2[jj], be2[jj]);
          float lo, hi;
          unpk(gg, lo, hi);
          *reinterpret_cast<float2 *>(go + st * N + 2 * jj) = make_float2(lo, hi);
          piacc[2 * jj] += first ? lo : 0.f;
          piacc[2 * jj + 1] += first ? hi : 0.f;
        }
      }
      __syncwarp();
      tile_out<N>(ALin + b * 32 * ROW, a.GA, dir, tb, cl);
      __syncwarp();                                          // the next request overwrites this tile
    }
    xcount += Tw;
    if (xcount >= 4096) flush();
  }
  flush();
}

template <int N>
void launch_fbs_t(dnbc_ctx *c, const FbsArgs &a, bool backward) {
  // warps per CTA: the backward sweep stages 27 KB per warp (five tiles), so CTAs of two warps pack an SM better
  const int nw = backward ? 2 : FBS_WARPS;
  const size_t smem = fbs_table_bytes<N>() + nw * (backward ? fbs_bwd_warp_bytes<N>() : fbs_fwd_warp_bytes<N>());
  const int64_t nseq = a.s1 - a.s0;
  const int64_t per_block = (int64_t)nw * 32;
  int per_sm = (int)((224 * 1024) / (smem + 1024));
  per_sm = per_sm > 16 / nw ? 16 / nw : (per_sm < 1 ? 1 : per_sm);
  int64_t blocks = (nseq + per_block - 1) / per_block;
  const int64_t cap = (int64_t)c->sm_count * per_sm;
  blocks = blocks > cap ? cap : blocks;
  if (backward) {
    cudaFuncSetAttribute(k_backward_s<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_backward_s<N><<<(int)blocks, nw * 32, smem, c->stream>>>(a);
  } else {
    cudaFuncSetAttribute(k_forward_s<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_forward_s<N><<<(int)blocks, nw * 32, smem, c->stream>>>(a);
  }
}

}  // namespace

// even state counts up to 12 and enough sequences to give every SM scheduler a warp of 32 of them
bool fbs_supported(const dnbc_ctx *c, int64_t nseq) {
  const int N = c->dm.N;
  if (N < 2 || N > 12 || (N & 1)) return false;
  if (c->opt_fb_seq_lanes == 2) return false;
  if (c->opt_fb_seq_lanes == 1) return true;
  return nseq >= (int64_t)c->sm_count * 4 * 32;
}

void launch_fbs(dnbc_ctx *c, int64_t s0, int64_t s1, int64_t p0, bool backward) {
  StatsLayout sl = stats_layout(c->dm);
  FbsArgs a{c->dm, c->tables(), c->d_off, s0, s1, p0, c->E, c->AL, c->GA, c->cs, c->d_stats, sl.pi, sl.xi};
  switch (c->dm.N) {
    case 2: launch_fbs_t<2>(c, a, backward); break;
    case 4: launch_fbs_t<4>(c, a, backward); break;
    case 6: launch_fbs_t<6>(c, a, backward); break;
    case 8: launch_fbs_t<8>(c, a, backward); break;
    case 10: launch_fbs_t<10>(c, a, backward); break;
    default: launch_fbs_t<12>(c, a, backward); break;
  }
}
