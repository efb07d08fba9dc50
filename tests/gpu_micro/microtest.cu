// GPU micro-validation of the sm_100a building blocks the sampler kernel relies on.
//   (1) fp16 hi/lo split GEMM  D[256x64] = A[256x64] * B^T  via tcgen05.mma (M=128,N=64,K=16),
//       hand-built SWIZZLE_NONE descriptors, K-major B and MN-major B, checked against fp64.
//   (2) TMEM used as thread-private storage (tcgen05.st / tcgen05.ld 32x32b.x32 round trip).
//   (3) cycle counts for the MMA sequence, the TMEM epilogue and the private ld/st.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/microtest tests/gpu_micro/microtest.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../us-potus-model_b200/csrc/ptx_sm100.cuh"

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e = (x);                                                              \
    if (e != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);  \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

constexpr int ROWS = 256, KD = 64, ND = 64;
constexpr uint32_t A_SBO = 128, B_SBO = 128, B_LBO = 1024;  // K-major, 8-row groups contiguous

struct Smem {
  alignas(128) __half a_hi[ROWS * KD + 512];
  alignas(128) __half a_lo[ROWS * KD + 512];
  alignas(128) __half b_hi[ND * KD];
  alignas(128) __half b_lo[ND * KD];
  alignas(8) uint64_t bar;
  uint32_t tmem_base;
};

// variant bit0: B consumed MN-major (computes A * Y with Y[k][n] = tile[row=k][col=n])
// variant bit1: padded A LBO (4096+16) to kill store bank conflicts
__global__ void __launch_bounds__(512, 1) k_mma(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                int variant, long long* cyc, int reps) {
  extern __shared__ __align__(128) unsigned char raw[];
  Smem& sm = *reinterpret_cast<Smem*>(raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_lbo = (variant & 2) ? 4096 + 16 : 4096;
  if (warp == 0) {
    ptx::tmem_alloc(&sm.tmem_base, 512);
    ptx::tmem_relinquish();
  }
  if (tid == 0) {
    ptx::mbar_init(&sm.bar, 1);
    ptx::fence_mbar_init();
  }
  // operands -> smem (generic proxy)
  for (int i = tid; i < ROWS * KD; i += 512) {
    int t = i / KD, k = i % KD;
    __half hi, lo;
    ptx::split_f16(A[i], hi, lo);
    uint32_t off = (k / 8) * a_lbo + (t / 8) * A_SBO + (t % 8) * 16 + (k % 8) * 2;
    *reinterpret_cast<__half*>(reinterpret_cast<unsigned char*>(sm.a_hi) + off) = hi;
    *reinterpret_cast<__half*>(reinterpret_cast<unsigned char*>(sm.a_lo) + off) = lo;
  }
  for (int i = tid; i < ND * KD; i += 512) {
    int r = i / KD, c = i % KD;
    __half hi, lo;
    ptx::split_f16(B[i] * 256.0f, hi, lo);
    uint32_t off = (c / 8) * B_LBO + (r / 8) * B_SBO + (r % 8) * 16 + (c % 8) * 2;
    *reinterpret_cast<__half*>(reinterpret_cast<unsigned char*>(sm.b_hi) + off) = hi;
    *reinterpret_cast<__half*>(reinterpret_cast<unsigned char*>(sm.b_lo) + off) = lo;
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tbase = sm.tmem_base;
  long long t0 = 0, t1 = 0, t2 = 0;
  uint32_t phase = 0;
  for (int rep = 0; rep < reps; ++rep) {
    if (tid == 0) {
      t0 = clock64();
      const uint32_t idesc = ptx::make_idesc_f16(128, 64, 0, (variant & 1));
      for (int tile = 0; tile < 2; ++tile) {
        for (int prod = 0; prod < 3; ++prod) {
          const __half* ap = (prod == 2) ? sm.a_lo : sm.a_hi;
          const __half* bp = (prod == 1) ? sm.b_lo : sm.b_hi;
          uint32_t dcol = tbase + tile * 64 + (prod == 0 ? 0 : 128);
          for (int ks = 0; ks < 4; ++ks) {
            uint32_t aaddr = ptx::smem_u32(ap) + tile * 16 * A_SBO + ks * 2 * a_lbo;
            uint64_t ad = ptx::make_smem_desc(aaddr, a_lbo, A_SBO);
            uint64_t bd;
            if (variant & 1) {  // MN-major view of the same tile: N=col, K=row
              uint32_t baddr = ptx::smem_u32(bp) + ks * 2 * B_SBO;  // K advances along rows: 16 rows = 2 groups
              bd = ptx::make_smem_desc(baddr, /*LBO (k/8)*/ B_SBO, /*SBO (mn/8)*/ B_LBO);
            } else {
              uint32_t baddr = ptx::smem_u32(bp) + ks * 2 * B_LBO;
              bd = ptx::make_smem_desc(baddr, B_LBO, B_SBO);
            }
            uint32_t acc = (prod == 0 || prod == 1) ? (ks > 0) : 1u;
            ptx::mma_f16_ss(dcol, ad, bd, idesc, acc);
          }
        }
      }
      ptx::mma_commit(&sm.bar);
    }
    ptx::mbar_wait(&sm.bar, phase);
    phase ^= 1;
    ptx::tc_fence_after();
    if (tid == 0) t1 = clock64();
    // epilogue: warp w -> lane quarter w%4, tile (w/4)>>1, column half (w/4)&1
    {
      const int q = warp & 3, g = warp >> 2, tile = g >> 1, half = g & 1;
      uint32_t taddr = tbase + ((uint32_t)(q * 32) << 16) + tile * 64 + half * 32;
      uint32_t d1[32], d2[32];
      ptx::tmem_ld32(taddr, d1);
      ptx::tmem_ld32(taddr + 128, d2);
      ptx::tmem_wait_ld();
      const int row = tile * 128 + q * 32 + lane;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float v = (__uint_as_float(d1[j]) + __uint_as_float(d2[j]) * (1.0f / 2048.0f)) * (1.0f / 256.0f);
        D[row * ND + half * 32 + j] = v;
      }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (tid == 0) t2 = clock64();
  }
  if (tid == 0 && blockIdx.x == 0) {
    cyc[0] = t1 - t0;
    cyc[1] = t2 - t1;
  }
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tbase, 512);
}

// TMEM as thread-private storage: 2 vectors x 32 floats per thread in columns [256,384) and [384,512)
__global__ void __launch_bounds__(512, 1) k_tmem_private(float* out, long long* cyc, int reps) {
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_base, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t my = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256 + (warp >> 2) * 32;
  uint32_t v[32], u[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __float_as_uint((float)(tid * 32 + j));
  ptx::tmem_st32(my, v);
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(-(float)(tid * 32 + j));
  ptx::tmem_st32(my + 128, v);
  ptx::tmem_wait_st();
  __syncthreads();
  long long t0 = clock64();
  float acc = 0.f;
  for (int r = 0; r < reps; ++r) {
    ptx::tmem_ld32(my, u);
    ptx::tmem_ld32(my + 128, v);
    ptx::tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float a = __uint_as_float(u[j]), b = __uint_as_float(v[j]);
      acc += a + b;  // must stay 0
      u[j] = __float_as_uint(a + 1.0f);
      v[j] = __float_as_uint(b - 1.0f);
    }
    ptx::tmem_st32(my, u);
    ptx::tmem_st32(my + 128, v);
    ptx::tmem_wait_st();
  }
  __syncthreads();
  long long t1 = clock64();
  ptx::tmem_ld32(my, u);
  ptx::tmem_wait_ld();
#pragma unroll
  for (int j = 0; j < 32; ++j) out[tid * 32 + j] = __uint_as_float(u[j]) + acc;
  if (tid == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 512);
}

int main() {
  std::vector<float> hA(ROWS * KD), hB(ND * KD), hD(ROWS * ND);
  srand(1843);
  for (auto& x : hA) x = ((rand() % 20001) - 10000) / 2500.0f;  // [-4,4]
  for (auto& x : hB) x = ((rand() % 20001) - 3000) / 250000.0f;  // ~[-0.012,0.068]
  float *dA, *dB, *dD;
  long long* dc;
  CK(cudaMalloc(&dA, hA.size() * 4));
  CK(cudaMalloc(&dB, hB.size() * 4));
  CK(cudaMalloc(&dD, hD.size() * 4));
  CK(cudaMalloc(&dc, 64));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem) + 128));
  int fails = 0;
  for (int variant = 0; variant < 4; ++variant) {
    CK(cudaMemset(dD, 0, hD.size() * 4));
    k_mma<<<1, 512, sizeof(Smem) + 128>>>(dA, dB, dD, variant, dc, 3);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    long long hc[2];
    CK(cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int t = 0; t < ROWS; ++t)
      for (int n = 0; n < ND; ++n) {
        double ref = 0;
        for (int k = 0; k < KD; ++k) {
          double b = (variant & 1) ? hB[k * KD + n] : hB[n * KD + k];
          ref += (double)hA[t * KD + k] * b;
        }
        maxerr = fmax(maxerr, fabs(ref - hD[t * ND + n]));
        maxref = fmax(maxref, fabs(ref));
      }
    bool ok = maxerr < 2e-6 * maxref + 1e-7;
    printf("mma variant %d (B %s-major, A lbo %s): max|err| %.3e (max|ref| %.3f) %s | cycles mma %lld epi+sync %lld\n", variant,
           (variant & 1) ? "MN" : "K", (variant & 2) ? "padded" : "4096", maxerr, maxref, ok ? "OK" : "FAIL", hc[0], hc[1]);
    fails += !ok;
  }
  // full-chip timing of the MMA kernel (148 CTAs)
  {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int reps = 200;
    k_mma<<<148, 512, sizeof(Smem) + 128>>>(dA, dB, dD, 2, dc, 10);
    CK(cudaEventRecord(e0));
    k_mma<<<148, 512, sizeof(Smem) + 128>>>(dA, dB, dD, 2, dc, reps);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("148 CTAs x %d reps of (24 MMAs + epilogue): %.3f ms -> %.2f us per rep per CTA\n", reps, ms, ms * 1e3 / reps);
  }
  // TMEM private storage
  {
    float* dout;
    CK(cudaMalloc(&dout, 512 * 32 * 4));
    const int reps = 100;
    k_tmem_private<<<1, 512>>>(dout, dc, reps);
    CK(cudaDeviceSynchronize());
    std::vector<float> ho(512 * 32);
    CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
    long long hc;
    CK(cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int i = 0; i < 512 * 32; ++i) bad += (ho[i] != (float)(i + reps));
    printf("tmem private storage: %d mismatches of %d %s | %lld cycles for %d x (2 ld32 + 2 st32 per thread, 16 warps) = %.1f cyc/iter\n",
           bad, 512 * 32, bad ? "FAIL" : "OK", hc, reps, (double)hc / reps);
    fails += (bad != 0);
  }
  printf(fails ? "MICROTEST FAILED (%d)\n" : "MICROTEST PASSED\n", fails);
  return fails;
}
