// On-device post-processing of the draws (SURVEY.md 8(f) row f2 / native inventory k5): what the reference's reports do with
// rstan::extract output on the host, done where the draws are, over EVERY sampling iteration of every chain:
//   README.Rmd:206-220, final_2016.R:708-760   per-state mean / sd / quantiles / P(win) of predicted_score[, T, s]
//   README.Rmd:230-248                         national vote per draw = state_weights-weighted mean of the state shares
//   README.Rmd:271-300                         electoral college: dem_ev = sum(ev * (share > 0.5)) per draw
//   generated quantities (poll_model_2020.stan:134-140)  predicted_score = inv_logit(mu_b)' of the kept draws, in extract()'s layout
// plus the ingredients of Stan's effective sample size / split R-hat of the monitored scalars (chain means, variances, and the
// chain-averaged autocovariance at every lag), and the run statistics that used to need the whole sampler_params table on the host.
// Everything is deterministic: fixed reduction trees, no floating-point atomics; quantiles are EXACT order statistics
// (bisection on the ordered bit pattern of the fp32 values, counting passes), interpolated like numpy / R type 7.
#include <cuda_runtime.h>
#include <cstdint>

namespace potus {

constexpr int PNT = 1024;

__device__ __forceinline__ double block_sum_d(double v, double* sh) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
  return t;
}
__device__ __forceinline__ unsigned long long block_sum_u(unsigned long long v, unsigned long long* sh) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  unsigned long long t = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
  return t;
}

// monitor [R][S+1] (logit scale; last column national_mu_b_average) -> quantity-major shares sh[S+2][R]:
//   rows 0..S-1 inv_logit(mu_b[s,T]); row S = national vote (weighted mean of the state shares); row S+1 = democratic electoral votes
extern "C" __global__ void __launch_bounds__(256) potus_post_shares_kernel(const float* __restrict__ mon, const float* __restrict__ w,
                                                                           const float* __restrict__ ev, int S, long long R, float* __restrict__ sh) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < R; r += nwarp) {
    float nat = 0.f, e = 0.f;
    for (int s = lane; s < S; s += 32) {
      const float p = 1.0f / (1.0f + __expf(-mon[r * (S + 1) + s]));
      sh[(long long)s * R + r] = p;
      nat = fmaf(w[s], p, nat);
      if (ev != nullptr && p > 0.5f) e += ev[s];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { nat += __shfl_xor_sync(0xffffffffu, nat, off); e += __shfl_xor_sync(0xffffffffu, e, off); }
    if (lane == 0) { sh[(long long)S * R + r] = nat; sh[(long long)(S + 1) * R + r] = e; }
  }
}

// per quantity (one block each): mean, sd (n-1), P(x > thr[q])   -> out[q][3]
extern "C" __global__ void __launch_bounds__(PNT) potus_post_moments_kernel(const float* __restrict__ sh, long long R, const float* __restrict__ thr,
                                                                            double* __restrict__ out) {
  __shared__ double shd[32];
  __shared__ unsigned long long shu[32];
  const float* x = sh + (long long)blockIdx.x * R;
  double s1 = 0;
  unsigned long long cnt = 0;
  const float t = thr[blockIdx.x];
  for (long long i = threadIdx.x; i < R; i += blockDim.x) { const float v = x[i]; s1 += v; cnt += v > t; }
  const double mean = block_sum_d(s1, shd) / (double)R;
  const unsigned long long c = block_sum_u(cnt, shu);
  double s2 = 0;
  for (long long i = threadIdx.x; i < R; i += blockDim.x) { const double d = (double)x[i] - mean; s2 += d * d; }
  const double var = block_sum_d(s2, shd) / (double)(R > 1 ? R - 1 : 1);
  if (threadIdx.x == 0) { out[blockIdx.x * 3 + 0] = mean; out[blockIdx.x * 3 + 1] = sqrt(var); out[blockIdx.x * 3 + 2] = (double)c / (double)R; }
}

// order-preserving map of fp32 to uint32 (all finite values, either sign)
__device__ __forceinline__ uint32_t fkey(float f) { const uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float fkey_inv(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// exact order statistics: block (q, j) finds the rank-th smallest (1-based rank[j]) of quantity q and the next one up:
// out[q][j][2].  32 counting passes of bisection over the key space, then one pass for the successor.
extern "C" __global__ void __launch_bounds__(PNT) potus_post_select_kernel(const float* __restrict__ sh, long long R, const long long* __restrict__ rank,
                                                                           int nrank, double* __restrict__ out) {
  __shared__ unsigned long long shu[32];
  const int q = blockIdx.x / nrank, j = blockIdx.x % nrank;
  const float* x = sh + (long long)q * R;
  const unsigned long long k = (unsigned long long)rank[j];
  uint32_t lo = 0u, hi = 0xffffffffu;   // smallest key v with count(key <= v) >= k
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    unsigned long long c = 0;
    for (long long i = threadIdx.x; i < R; i += blockDim.x) c += fkey(x[i]) <= mid;
    c = block_sum_u(c, shu);
    if (c >= k) hi = mid; else lo = mid + 1;
  }
  // successor in sorted order: the same value if it has duplicates beyond rank k, else the smallest larger key
  unsigned long long cle = 0;
  uint32_t nxt = 0xffffffffu;
  for (long long i = threadIdx.x; i < R; i += blockDim.x) { const uint32_t key = fkey(x[i]); cle += key <= lo; if (key > lo && key < nxt) nxt = key; }
  cle = block_sum_u(cle, shu);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) nxt = min(nxt, __shfl_xor_sync(0xffffffffu, nxt, off));
  __shared__ uint32_t shm[32];
  __syncthreads();
  if ((threadIdx.x & 31) == 0) shm[threadIdx.x >> 5] = nxt;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t m = 0xffffffffu;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = min(m, shm[w]);
    const double a = (double)fkey_inv(lo);
    const double b = (cle > k || k >= (unsigned long long)R) ? a : (double)fkey_inv(m);
    out[((long long)q * nrank + j) * 2 + 0] = a;
    out[((long long)q * nrank + j) * 2 + 1] = b;
  }
}

// Stan ESS / split R-hat ingredients of monitored scalar q (logit scale, as monitored), one block per (q, group of chains):
//   cstat[q][c][6]   = mean, biased variance (1/n), and the same for the first and the second half of chain c
//   acov[q][g][lag]  = sum over the chains of group g of the biased autocovariance at `lag` (1/n sum_i xc_i xc_{i+lag})
extern "C" __global__ void __launch_bounds__(512) potus_post_acov_kernel(const float* __restrict__ mon, int C, int n, int S1, int cpg,
                                                                         double* __restrict__ cstat, double* __restrict__ acov) {
  extern __shared__ float xs[];   // [n] centred chain
  __shared__ double shd[32];
  const int q = blockIdx.x, g = blockIdx.y, tid = threadIdx.x;
  const int nl = (n + (int)blockDim.x - 1) / (int)blockDim.x;   // lags per thread (<= 4 for n <= 2048)
  double acc[4] = {0, 0, 0, 0};
  for (int c = g * cpg; c < min(C, (g + 1) * cpg); ++c) {
    const float* x = mon + ((long long)c * n) * S1 + q;
    double s = 0, s_a = 0, s_b = 0;
    const int h = n / 2;
    for (int i = tid; i < n; i += blockDim.x) { const double v = x[(long long)i * S1]; s += v; if (i < h) s_a += v; if (i >= n - h) s_b += v; }
    const double mean = block_sum_d(s, shd) / n;
    const double mean_a = block_sum_d(s_a, shd) / (h > 0 ? h : 1), mean_b = block_sum_d(s_b, shd) / (h > 0 ? h : 1);
    __syncthreads();
    double v2 = 0, v2a = 0, v2b = 0;
    for (int i = tid; i < n; i += blockDim.x) {
      const double v = x[(long long)i * S1];
      const double d = v - mean;
      xs[i] = (float)d;
      v2 += d * d;
      if (i < h) v2a += (v - mean_a) * (v - mean_a);
      if (i >= n - h) v2b += (v - mean_b) * (v - mean_b);
    }
    const double var = block_sum_d(v2, shd) / n, var_a = block_sum_d(v2a, shd) / (h > 0 ? h : 1), var_b = block_sum_d(v2b, shd) / (h > 0 ? h : 1);
    if (tid == 0) {
      double* o = cstat + ((long long)q * C + c) * 6;
      o[0] = mean; o[1] = var; o[2] = mean_a; o[3] = var_a; o[4] = mean_b; o[5] = var_b;
    }
    __syncthreads();
    for (int j = 0; j < nl && j < 4; ++j) {
      const int lag = tid + j * blockDim.x;
      if (lag < n) {
        double a = 0;
        for (int i = 0; i + lag < n; ++i) a += (double)xs[i] * (double)xs[i + lag];
        acc[j] += a / n;
      }
    }
    __syncthreads();
  }
  for (int j = 0; j < nl && j < 4; ++j) {
    const int lag = tid + j * blockDim.x;
    if (lag < n) acov[((long long)q * gridDim.y + g) * n + lag] = acc[j];
  }
}

// predicted_score of the kept draws in rstan::extract's layout: out[r + R*(t + T*s)] = inv_logit(mu_b[s,t]) (fp32)
extern "C" __global__ void __launch_bounds__(256) potus_post_pscore_kernel(const float* __restrict__ draws, long long R, int draw_len, int S, int T,
                                                                           float* __restrict__ out) {
  const long long n = R * (long long)S * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i % R, ts = i / R;
    const int t = (int)(ts % T), s = (int)(ts / T);
    out[i] = 1.0f / (1.0f + __expf(-draws[r * draw_len + s + (long long)S * t]));
  }
}

// run statistics from the per-iteration diagnostics [C][nt][8] (what potus_run used to copy to the host in full):
// out[0..5] = total leapfrogs, sampling leapfrogs, sampling divergences, sum accept (sampling), sum depth (sampling), (unused)
extern "C" __global__ void __launch_bounds__(PNT) potus_post_runstats_kernel(const float* __restrict__ sp, int C, int nt, int nw, double* __restrict__ out) {
  __shared__ double shd[32];
  double a[5] = {0, 0, 0, 0, 0};
  const long long n = (long long)C * nt;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float* p = sp + i * 8;
    const int it = (int)(i % nt);
    a[0] += (double)(long long)p[4];
    if (it >= nw) { a[1] += (double)(long long)p[4]; a[2] += (double)(long long)p[5]; a[3] += p[1]; a[4] += p[3]; }
  }
  for (int k = 0; k < 5; ++k) { const double t = block_sum_d(a[k], shd); if (threadIdx.x == 0) out[k] = t; }
}

}  // namespace potus
