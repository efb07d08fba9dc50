// Resident NUTS kernel for poll_model_2020 (sm_100a).  One CTA = one chain; the chain's position
// lives in shared memory, its half-step momentum and sqrt(metric) in TMEM (thread-private columns),
// and every leapfrog is:  scan -> tcgen05 GEMM (L0 W) -> poll residuals -> tcgen05 GEMM (L0^T G)
// -> scan -> momentum/position update + NUTS bookkeeping, without touching HBM for the state.
//
// What it restates (reference file:line; the sampler itself is Stan 2.24.1, not in the tree):
//   poll_model_2020.stan:70-113  transformed parameters   -> eval_point() forward half
//   poll_model_2020.stan:115-132 log density               -> eval_point() energy + hand-derived gradient
//   poll_model_2020.stan:134-140 generated quantities      -> eval_point() emit hooks + host (predicted_score)
//   Stan base_nuts::transition / build_tree (multinomial NUTS, generalised U-turn with the two
//   extra sub-tree checks), expl_leapfrog, diag_e_metric, stepsize/var/windowed adaptation,
//   init_stepsize -> transition(), find_stepsize(), the adaptation block of potus_nuts_kernel.
// The CPU statement of exactly this algorithm (iterative tree, Philox streams) is
// oracle/potus_oracle.c (tree_mode 1); tests compare the two decision by decision.
//
// Conventions that keep the hot loops mask-free: every internal vector is zero in its padding
// slots (walk rows >= T, state columns >= S, nz slots >= NZ), in shared memory, TMEM and global
// memory alike, so element-wise updates need no validity tests.
#include <cuda_runtime.h>
#include <math_constants.h>
#include "ptx_sm100.cuh"
#include "potus_layout.h"

namespace potus {

extern __shared__ __align__(128) unsigned char smem_raw[];

struct Ctl {
  uint64_t bar_mma[2];
  uint64_t bar_load;
  uint32_t tmem_base;
  int chain;
  double U;        // potential energy of the last evaluated point (centred: -lp + lp_const)
  double u_extra;  // -(log-density terms of rho_e_bias)
  float rho, mu_e, sig_rho, rho_term;  // rho_term = sig_e*rho/sqrt(1-rho^2)
  float rn_total;
  float pad0;
  alignas(16) float kred[2][NWARP];
  alignas(16) float mred[NWARP][8];
  alignas(16) float lred[2][3][NWARP];   // leaf reduction: |P|^2 and the two level-0 U-turn sums
  ChainState cs;
  // per-transition statistics kept by thread 0 (not needed for control flow)
  double U_samp, H_samp, U_prop, H_prop;
  float sum_metro;
  int pad1;
#ifdef POTUS_PROF
  unsigned long long prof[40];
#endif
};
static_assert(sizeof(Ctl) <= 2048, "control block");
static_assert(NT == 512 && EPT == 32 && DPW == 16, "this file is written for 16 warps x 32 elements");

#define SMP(T_, off) (reinterpret_cast<T_*>(smem_raw + (off)))
// optional phase clocks (development builds, -DPOTUS_PROF): thread 0 accumulates cycles per phase
#ifdef POTUS_PROF
#define PROF_DECL long long prof_t_ = clock64()
#define PROF(i) do { if (threadIdx.x == 0) { long long n_ = clock64(); CTL().prof[i] += (unsigned long long)(n_ - prof_t_); prof_t_ = n_; } } while (0)
#define PROF_RESET prof_t_ = clock64()
#define PROF_COUNT CTL().prof[39] += 1
#else
#define PROF_DECL
#define PROF(i)
#define PROF_RESET
#define PROF_COUNT
#endif
__device__ __forceinline__ const ModelDev& MD() { return *SMP(const ModelDev, SM_MODEL); }
__device__ __forceinline__ Ctl& CTL() { return *SMP(Ctl, SM_CTL); }
__device__ __forceinline__ float* sQZ() { return SMP(float, SM_QZ); }
__device__ __forceinline__ float* sQNZ() { return SMP(float, SM_QNZ); }
__device__ __forceinline__ float* sGNZ() { return SMP(float, SM_GNZ); }
__device__ __forceinline__ float* sSCR() { return SMP(float, SM_A); }
__device__ __forceinline__ float* sRR() { return SMP(float, SM_RR); }
__device__ __forceinline__ float* sPSUM() { return SMP(float, SM_PSUM); }
__device__ __forceinline__ float* sE() { return SMP(float, SM_E); }
__device__ __forceinline__ float* sEBAR() { return SMP(float, SM_E) + 256; }
__device__ __forceinline__ float* sTOT() { return SMP(float, SM_TOT); }
__device__ __forceinline__ float* sPRIOR() { return SMP(float, SM_PRIOR); }
__device__ __forceinline__ double* sRED() { return SMP(double, SM_RED); }
__device__ __forceinline__ uint32_t* sPKI() { return SMP(uint32_t, SM_PK); }
__device__ __forceinline__ float* sPKF(int which) { return SMP(float, SM_PK + which * NPOLL_CAP * 4); }

// thread-private TMEM window (32 columns) of the calling thread, column offset 0
__device__ __forceinline__ uint32_t tpriv() {
  const uint32_t w = threadIdx.x >> 5;
  return CTL().tmem_base + (((w & 3u) * 32u) << 16) + (w >> 2) * 32u;
}
// owned position elements come in 16 float2 pairs: pair d of a walk lane = states (2l,2l+1) on day 16w+d;
// pair d of an nz lane = nz slots nz_slot(w, ln, 2d), +1 (see potus_layout.h).
__device__ __forceinline__ float2* qpair(int d) {
  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  if (l < ZLANES) return reinterpret_cast<float2*>(sQZ() + (16 * w + d) * QZ_PITCH + 2 * l);
  return reinterpret_cast<float2*>(sQNZ() + nz_slot(w, l - ZLANES, 2 * d));
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10, same streams as oracle/potus_oracle.c
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t h0 = __umulhi(0xD2511F53u, c[0]), l0 = 0xD2511F53u * c[0];
    uint32_t h1 = __umulhi(0xCD9E8D57u, c[2]), l1 = 0xCD9E8D57u * c[2];
    uint32_t n0 = h1 ^ c[1] ^ k0, n2 = h0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = l1; c[2] = n2; c[3] = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ float u01(uint32_t w) { return ((float)(w >> 9) + 0.5f) * (1.0f / 8388608.0f); }
__device__ __forceinline__ void rng_words(unsigned long long seed, uint32_t chain, uint32_t idx, uint32_t iter, uint32_t stream,
                                          uint32_t sub, uint32_t (&w)[4]) {
  w[0] = idx; w[1] = iter; w[2] = stream | (sub << 8); w[3] = chain;
  philox(w, (uint32_t)seed, (uint32_t)(seed >> 32));
}
__device__ __forceinline__ float rng_normal(unsigned long long seed, uint32_t chain, uint32_t idx, uint32_t iter, uint32_t stream,
                                            uint32_t sub) {
  uint32_t w[4];
  rng_words(seed, chain, idx, iter, stream, sub, w);
  return sqrtf(-2.0f * logf(u01(w[0]))) * cospif(2.0f * u01(w[1]));
}

// ------------------------------------------------------------------------------------------------
// thread-private TMEM vectors and owner-layout global vectors (element e of thread tid at [e*512+tid])
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tm_ld16(uint32_t tp, uint32_t col, float (&v)[16]) {
  ptx::tmem_ld16f(tp + col, v);
  ptx::tmem_wait_ld();
}
__device__ __forceinline__ void tm_st16(uint32_t tp, uint32_t col, const float (&v)[16]) { ptx::tmem_st16f(tp + col, v); }
// issue-only variants: several loads share one tcgen05.wait::ld (each ld+wait round trip costs a few hundred cycles)
__device__ __forceinline__ void tm_ld16_nowait(uint32_t tp, uint32_t col, float (&v)[16]) { ptx::tmem_ld16f(tp + col, v); }
__device__ __forceinline__ float* slot_ptr(float* ws, int slot) { return ws + (size_t)slot * VEC; }

// ---- owner-layout vectors in global memory (potus_layout.h: oslot): float4 group k of the calling thread
__device__ __forceinline__ float4* own4(float* vec) { return reinterpret_cast<float4*>(vec) + threadIdx.x; }
__device__ __forceinline__ const float4* own4(const float* vec) { return reinterpret_cast<const float4*>(vec) + threadIdx.x; }
// (group k of the thread = own4(vec)[k * NT], k = 0..7)
__device__ __forceinline__ void st_groups4(float4* o, int k0, const float* v, int n) {   // n floats (multiple of 4) -> groups k0..
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (4 * i < n) o[(k0 + i) * NT] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void ld_groups4(const float4* o, int k0, float* v, int n) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (4 * i < n) { const float4 t = o[(k0 + i) * NT]; v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w; }
}

__device__ __forceinline__ void tm_to_global(uint32_t tp, uint32_t col, float* g) {
  float4* o = own4(g);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float v[16];
    tm_ld16(tp, col + 16 * h, v);
    st_groups4(o, 4 * h, v, 16);
  }
}
__device__ __forceinline__ void global_to_tm(uint32_t tp, uint32_t col, const float* g) {
  const float4* o = own4(g);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float v[16];
    ld_groups4(o, 4 * h, v, 16);
    tm_st16(tp, col + 16 * h, v);
  }
  ptx::tmem_wait_st();
}
__device__ __forceinline__ void q_to_global(float* g) {
  float4* o = own4(g);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float2 a = *qpair(2 * k), b = *qpair(2 * k + 1);
    o[k * NT] = make_float4(a.x, a.y, b.x, b.y);
  }
}
__device__ __forceinline__ void global_to_q(const float* g) {
  const float4* o = own4(g);
  float4 v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = o[k * NT];
#pragma unroll
  for (int k = 0; k < 8; ++k) { *qpair(2 * k) = make_float2(v[k].x, v[k].y); *qpair(2 * k + 1) = make_float2(v[k].z, v[k].w); }
}

// inclusive scan of affine maps x -> A x + B over the lanes of a warp (AR(1) recurrences)
__device__ __forceinline__ void affine_scan(float& A, float& B, int l) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    float Ap = __shfl_up_sync(0xffffffffu, A, off), Bp = __shfl_up_sync(0xffffffffu, B, off);
    if (l >= off) { B = A * Bp + B; A = A * Ap; }
  }
}

// One poll: centred log-likelihood term f = ll(eta) - ll(eta_hat) and residual r = d f / d eta, with d = eta - eta_hat,
// p = p_hat = inv_logit(eta_hat), y/n = p + rho_hat:   f = n [ rho_hat d - g(d) ],  r = n [ rho_hat - g'(d) ],
// g(d) = log(1 - p + p e^d) - p d  =  sum_{k>=2} kappa_k d^k / k!   (the cumulant series of Bernoulli(p); kappa_2 = v = p(1-p),
// kappa_3 = v w, kappa_4 = v (1 - 6v), kappa_5 = v w (1 - 12 v), ... with w = 1 - 2p).
//   |d| < 0.4 (every poll of a chain near its typical set): the series through d^10 and ITS OWN derivative -- an exactly consistent
//     energy / gradient pair, truncation error < 7e-10 v, ~45 FMAs, no divergence inside a warp;
//   0.4 <= |d| < 12: the closed form with expm1f / log1pf (warm-up transients);
//   beyond: direct softplus difference (accuracy irrelevant out there).
// (the series branch alone: branch-free, so that two polls of one thread can be scheduled into each other)
__device__ __forceinline__ void poll_term_series(float d, float n, float ph, float rh, float& f, float& r) {
  const float v = ph * (1.0f - ph), w = 1.0f - 2.0f * ph, x2 = d * d;
  const float e4 = fmaf(-6.0f, v, 1.0f) * (1.0f / 24.0f);
  const float e6 = fmaf(v, fmaf(120.0f, v, -30.0f), 1.0f) * (1.0f / 720.0f);
  const float e8 = fmaf(v, fmaf(v, fmaf(-5040.0f, v, 1680.0f), -126.0f), 1.0f) * (1.0f / 40320.0f);
  const float e10 = fmaf(v, fmaf(v, fmaf(v, fmaf(362880.0f, v, -151200.0f), 17640.0f), -510.0f), 1.0f) * (1.0f / 3628800.0f);
  const float o5 = fmaf(-12.0f, v, 1.0f) * (1.0f / 120.0f);
  const float o7 = fmaf(v, fmaf(360.0f, v, -60.0f), 1.0f) * (1.0f / 5040.0f);
  const float o9 = fmaf(v, fmaf(v, fmaf(-20160.0f, v, 5040.0f), -252.0f), 1.0f) * (1.0f / 362880.0f);
  const float ev = x2 * fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, e10, e8), e6), e4), 0.5f);
  const float od = x2 * fmaf(x2, fmaf(x2, fmaf(x2, o9, o7), o5), 1.0f / 6.0f);
  const float g = v * fmaf(w * d, od, ev);
  const float dev = d * fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, 10.0f * e10, 8.0f * e8), 6.0f * e6), 4.0f * e4), 1.0f);
  const float dod = x2 * fmaf(x2, fmaf(x2, fmaf(x2, 9.0f * o9, 7.0f * o7), 5.0f * o5), 0.5f);
  const float gp = v * fmaf(w, dod, dev);
  f = n * fmaf(rh, d, -g);
  r = n * (rh - gp);
}
// (|d| >= 0.4: warm-up transients and the far tail)
__device__ __forceinline__ void poll_term_wide(float eta, float n, float eh, float ph, float rh, float& f, float& r) {
  const float d = eta - eh;
  if (fabsf(d) < 12.0f) {
    // ll(eta) - ll(eta_hat) = n [ (y/n) d - log1p(p_hat expm1(d)) ]
    const float em1 = expm1f(d);
    const float uu = ph * em1;
    f = n * (rh * d + (ph * d - log1pf(uu)));
    r = n * (rh - ph * (1.0f - ph) * __fdividef(em1, 1.0f + uu));
  } else {  // far tail: direct, stable softplus difference
    const float sp = fmaxf(eta, 0.f) + log1pf(__expf(-fabsf(eta)));
    const float sph = fmaxf(eh, 0.f) + log1pf(__expf(-fabsf(eh)));
    const float sg = 1.0f / (1.0f + __expf(-eta));
    f = n * ((ph + rh) * d - (sp - sph));
    r = n * ((ph + rh) - sg);
  }
}
// (out-of-line copy for callers that want the common path to stay one basic block)
__device__ __noinline__ void poll_term_wide_cold(float eta, float n, float eh, float ph, float rh, float& f, float& r) {
  poll_term_wide(eta, n, eh, ph, rh, f, r);
}
__device__ __forceinline__ void poll_term(float eta, float n, float eh, float ph, float rh, float& f, float& r) {
  const float d = eta - eh;
  const float ad = fabsf(d);
  if (ad < 0.4f) {
    poll_term_series(d, n, ph, rh, f, r);
  } else if (ad < 12.0f) {
    const float em1 = expm1f(d);
    const float uu = ph * em1;
    f = n * (rh * d + (ph * d - log1pf(uu)));
    r = n * (rh - ph * (1.0f - ph) * __fdividef(em1, 1.0f + uu));
  } else {
    const float sp = fmaxf(eta, 0.f) + log1pf(__expf(-fabsf(eta)));
    const float sph = fmaxf(eh, 0.f) + log1pf(__expf(-fabsf(eh)));
    const float sg = 1.0f / (1.0f + __expf(-eta));
    f = n * ((ph + rh) * d - (sp - sph));
    r = n * ((ph + rh) - sg);
  }
}

// Accumulator read-out shared by both GEMMs: (D1 + D2/2048) * scale (+ prior for day rows) -> fp32
// scratch [row][53].  The scratch aliases the operand planes, so nothing is written before the LAST
// commit (bar_mma[1]) has completed; tile-0 warps still overlap their TMEM loads with tile 1's MMAs.
__device__ __forceinline__ void tmem_epilogue(float scale, bool add_prior, uint32_t parity) {
  Ctl& ctl = CTL();
  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31, T = MD().T;
  const int g = w >> 2, tile = g >> 1, half = g & 1, qd = w & 3;
  ptx::mbar_wait(&ctl.bar_mma[tile], parity);
  ptx::tc_fence_after();
  const int row = tile * 128 + qd * 32 + l;
  const uint32_t taddr = ctl.tmem_base + ((uint32_t)(qd * 32) << 16) + tile * 64 + half * 32;
  const bool rowok = (row < T) || (row == PB_ROW);
  const bool prior_row = add_prior && row < T;
  float* out = sSCR() + row * SCR_PITCH + half * 32;
  const float* pr = sPRIOR() + half * 32;
  float d1[16], d2[16];
  ptx::tmem_ld16f(taddr + TM_D1, d1);
  ptx::tmem_ld16f(taddr + TM_D2, d2);
  ptx::tmem_wait_ld();
  if (tile == 0) ptx::mbar_wait(&ctl.bar_mma[1], parity);
  float2* out2 = reinterpret_cast<float2*>(out);   // (row*54 + 32*half) is even: 8-byte aligned
  if (rowok) {
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      float v0 = fmaf(d2[j], 1.0f / 2048.0f, d1[j]) * scale, v1 = fmaf(d2[j + 1], 1.0f / 2048.0f, d1[j + 1]) * scale;
      if (prior_row) { v0 += pr[j]; v1 += pr[j + 1]; }
      out2[j >> 1] = make_float2(v0, v1);
    }
  }
  // second 16 columns: half 0 -> cols 16..31, half 1 -> cols 48..63 of which only 48..51 exist
  ptx::tmem_ld16f(taddr + TM_D1 + 16, d1);
  ptx::tmem_ld16f(taddr + TM_D2 + 16, d2);
  ptx::tmem_wait_ld();
  if (rowok) {
    if (half == 0) {
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        float v0 = fmaf(d2[j], 1.0f / 2048.0f, d1[j]) * scale, v1 = fmaf(d2[j + 1], 1.0f / 2048.0f, d1[j + 1]) * scale;
        if (prior_row) { v0 += pr[16 + j]; v1 += pr[16 + j + 1]; }
        out2[8 + (j >> 1)] = make_float2(v0, v1);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; j += 2) {
        float v0 = fmaf(d2[j], 1.0f / 2048.0f, d1[j]) * scale, v1 = fmaf(d2[j + 1], 1.0f / 2048.0f, d1[j + 1]) * scale;
        if (prior_row) { v0 += pr[16 + j]; v1 += pr[16 + j + 1]; }
        out2[8 + (j >> 1)] = make_float2(v0, v1);
      }
    }
  }
  ptx::tc_fence_before();
}

struct Emit {
  float* draw;     // [draw_len] or null
  float* monitor;  // [S+1] or null
  float* mu;       // [S*T] or null (test hook)
};

// issue one split-precision GEMM: D1 = A_hi B_hi ; D2 = A_hi B_lo + A_lo B_hi   (2 M-tiles, K = 64)
__device__ __forceinline__ void issue_gemm(bool b_mn_major) {
  Ctl& ctl = CTL();
  ptx::tc_fence_after();
  const uint32_t idesc = ptx::make_idesc_f16(128, 64, 0, b_mn_major ? 1 : 0);
  const uint32_t a0 = ptx::smem_u32(smem_raw + SM_A), b0 = ptx::smem_u32(smem_raw + SM_B), tb = ctl.tmem_base;
#pragma unroll
  for (int tile = 0; tile < 2; ++tile) {
#pragma unroll
    for (int prod = 0; prod < 3; ++prod) {
      const uint32_t ap = a0 + (prod == 2 ? A_PLANE : 0) + tile * 16 * A_SBO;
      const uint32_t bp = b0 + (prod == 1 ? B_PLANE : 0);
      const uint32_t dcol = tb + tile * 64 + (prod == 0 ? TM_D1 : TM_D2);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t ad = ptx::make_smem_desc(ap + ks * 2 * A_LBO, A_LBO, A_SBO);
        const uint64_t bd = b_mn_major ? ptx::make_smem_desc(bp + ks * 2 * B_SBO, /*LBO: k groups*/ B_SBO, /*SBO: n groups*/ B_LBO)
                                       : ptx::make_smem_desc(bp + ks * 2 * B_LBO, B_LBO, B_SBO);
        ptx::mma_f16_ss(dcol, ad, bd, idesc, (prod == 2) ? 1u : (uint32_t)(ks > 0));
      }
    }
    ptx::mma_commit(&ctl.bar_mma[tile]);
  }
}

// ================================================================================================
// eval_point: potential U (-> ctl.U) and its gradient (-> TMEM TM_G, owner layout) at the position
// held in shared memory.  Called by all 512 threads.  Both mbarriers complete exactly twice per call,
// so their phase parity is 0 for GEMM 1 and 1 for GEMM 2 on every call.
// ================================================================================================
// Leaf mode (the NUTS inner loop): the gradient never goes to TMEM; P11 continues straight into the momentum
// update and the per-leaf bookkeeping that only needs this thread's own elements (see transition()).
struct LeafTail {
  float hs;          // signed half step
  // (owner-layout float4 views of the calling thread, see own4())
  const float4* Lr;  // odd leaf: momentum of the previous (even) leaf = Left_0 of the level-0 merge, else null
  float4* Fs;        // even leaf: FIRST[...] slot that receives this leaf's momentum, else null
  float4* Es;        // leaf closing a level-1 left half: its e slot, else null
  float kk, c1a, c1b;  // out: this thread's partial |P|^2 and the two level-0 U-turn dot products
};

template <bool LEAF>
__device__ __forceinline__ void eval_body(const Emit em, LeafTail& lt) {
  const ModelDev& m = MD();
  Ctl& ctl = CTL();
  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  const int T = m.T, S = m.S;
  const bool zlane = l < ZLANES;
  const int nd = min(max(T - 16 * w, 0), 16);       // day rows of this warp that exist
  const int ndw = min(max(T - 1 - 16 * w, 0), 16);  // ... that carry a walk innovation (t <= T-2)
  float qsq = 0.f;
  PROF_DECL;

  // ---------------- P1: reverse scan of the walk innovations (poll_model_2020.stan:86 collapsed)
  // (the running sums are recomputed from shared memory in P2 / P11 instead of being carried in 32 registers
  //  across the barriers: this kernel runs at the 128-register limit)
  if (zlane) {
    float run0 = 0.f, run1 = 0.f;
    const float* qz = sQZ() + (16 * w) * QZ_PITCH + 2 * l;
#pragma unroll
    for (int d = 15; d >= 0; --d) {
      const float2 z = *reinterpret_cast<const float2*>(qz + d * QZ_PITCH);
      qsq = fmaf(z.x, z.x, fmaf(z.y, z.y, qsq));
      if (d < ndw) { run0 += z.x; run1 += z.y; }
    }
    *reinterpret_cast<float2*>(sTOT() + w * 52 + 2 * l) = make_float2(run0, run1);
  } else {
    const float* qn = sQNZ() + nz_slot(w, l - ZLANES, 0);
#pragma unroll
    for (int d = 0; d < 16; ++d) {
      const float2 v = *reinterpret_cast<const float2*>(qn - d * (2 * NZ_LANES));
      qsq = fmaf(v.x, v.x, fmaf(v.y, v.y, qsq));
    }
    // rho's unconstrained value carries no N(0,1) term.  Only the OWNER of that slot may read it here: we are before
    // barrier S1 and the owner updated it in advance_q with no barrier since (any other reader races; seen by racecheck
    // and as a run-to-run flake, profiles/r01_d_racecheck_fused_kernel.log)
    if (tid == m.urho_owner) { const float v = sQNZ()[m.nz_urho]; qsq -= v * v; }
  }
  __syncthreads();  // S1
  PROF(0);
  // ---------------- P2: W -> fp16 hi/lo operand planes (K-major, SWIZZLE_NONE)
  {
    float carry0 = 0.f, carry1 = 0.f, zt0 = 0.f, zt1 = 0.f, zb0 = 0.f, zb1 = 0.f;
    const bool act = l < m.npair;
    if (act) {
#pragma unroll
      for (int w2 = 1; w2 < NWARP; ++w2) {   // fixed trip count: all loads issue back to back
        const float2 t2 = *reinterpret_cast<const float2*>(sTOT() + w2 * 52 + 2 * l);
        if (w2 > w) { carry0 += t2.x; carry1 += t2.y; }
      }
      zt0 = sQNZ()[m.nz_zT + 2 * l];
      zb0 = sQNZ()[m.nz_zb + 2 * l];
      if (2 * l + 1 < S) { zt1 = sQNZ()[m.nz_zT + 2 * l + 1]; zb1 = sQNZ()[m.nz_zb + 2 * l + 1]; }
    }
    const float base0 = m.a_T * zt0 + m.a_w * carry0, base1 = m.a_T * zt1 + m.a_w * carry1;
    unsigned char* ap = smem_raw + SM_A + (uint32_t)(l >> 2) * A_LBO + (uint32_t)(l & 3) * 4 + (uint32_t)(2 * w) * A_SBO;
    const float* qz = sQZ() + (16 * w) * QZ_PITCH + 2 * l;
    float run0 = 0.f, run1 = 0.f;
#pragma unroll
    for (int d = 15; d >= 0; --d) {
      float v0 = 0.f, v1 = 0.f;
      if (d < ndw) { const float2 z = *reinterpret_cast<const float2*>(qz + d * QZ_PITCH); run0 += z.x; run1 += z.y; }
      if (act) {
        if (d < nd) { v0 = fmaf(m.a_w, run0, base0); v1 = fmaf(m.a_w, run1, base1); }
        else if (16 * w + d == PB_ROW) { v0 = m.a_b * zb0; v1 = m.a_b * zb1; }
        if (2 * l + 1 >= S) v1 = 0.f;
      }
      const __half2 hi = __floats2half2_rn(v0, v1);
      const float2 hf = __half22float2(hi);
      const __half2 lo = __floats2half2_rn((v0 - hf.x) * 2048.0f, (v1 - hf.y) * 2048.0f);
      const uint32_t off = (uint32_t)(d >> 3) * A_SBO + (uint32_t)(d & 7) * 16;
      *reinterpret_cast<__half2*>(ap + off) = hi;
      *reinterpret_cast<__half2*>(ap + A_PLANE + off) = lo;
    }
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();  // S2
  PROF(1);
  // ---------------- P3: mu_b^T = W^T X^T on the tensor core
  if (tid == 0) issue_gemm(false);
  // overlapped with the MMA: AR(1) partisan non-response bias, poll_model_2020.stan:91-93 (warp 1)
  if (w == 1) {
    if (m.full) {
      const float* ze = sQNZ() + m.nz_ze;
      const float u_rho = sQNZ()[m.nz_urho], u_mu = sQNZ()[m.nz_umu];
      const float rho = 1.0f / (1.0f + __expf(-u_rho));
      const float mu_e = 0.02f * u_mu;
      const float s2 = sqrtf(fmaxf(1.0f - rho * rho, 0.f));
      const float sig_rho = s2 * m.sig_e;
      const float cst = mu_e * (1.0f - rho);
      float A = 1.f, B = 0.f;
      float u[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int t = 8 * l + j;
        float uj = 0.f, aj = 1.f;
        if (t < T) { aj = rho; uj = (t == 0) ? m.sig_e * ze[0] : cst + sig_rho * ze[t]; }
        u[j] = uj;
        B = aj * B + uj; A = aj * A;
      }
      affine_scan(A, B, l);
      float ein = __shfl_up_sync(0xffffffffu, B, 1);
      if (l == 0) ein = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int t = 8 * l + j;
        if (t < T) { ein = rho * ein + u[j]; sE()[t] = ein; }
      }
      if (l == 0) {
        ctl.rho = rho; ctl.mu_e = mu_e; ctl.sig_rho = sig_rho; ctl.rho_term = m.sig_e * rho / fmaxf(s2, 1e-20f);
        // -(normal(0.7,0.1) prior + log-Jacobian of the (0,1) transform), fp64
        const double r = 1.0 / (1.0 + exp(-(double)u_rho));
        ctl.u_extra = 0.5 * ((r - 0.7) / 0.1) * ((r - 0.7) / 0.1) - log(r) - log1p(-r);
      }
    } else if (l == 0) {
      ctl.u_extra = 0.0;
    }
  }
  PROF(2);
  // ---------------- P4: epilogue 1: TMEM -> mu_b (+prior) in the fp32 scratch
  tmem_epilogue(1.0f / 256.0f, true, 0);
  __syncthreads();  // S3
  PROF(3);
  // ---------------- optional outputs (transformed parameters / generated quantities of this point)
  if (em.draw != nullptr) {
    float* o = em.draw;
    for (int i = tid; i < S * T; i += NT) { int t = i / S, s = i - t * S; o[i] = sSCR()[t * SCR_PITCH + s]; }
    o += S * T;
    for (int i = tid; i < m.P; i += NT) o[i] = m.sig_c * sQNZ()[m.nz_c + i];
    o += m.P;
    for (int i = tid; i < m.M; i += NT) o[i] = m.full ? m.sig_m * sQNZ()[m.nz_m + i] : 0.f;
    o += m.M;
    for (int i = tid; i < m.Pop; i += NT) o[i] = m.full ? m.sig_pop * sQNZ()[m.nz_pop + i] : 0.f;
    o += m.Pop;
    for (int i = tid; i < T; i += NT) o[i] = m.full ? sE()[i] : 0.f;
    o += T;
    for (int i = tid; i < S; i += NT) o[i] = sSCR()[PB_ROW * SCR_PITCH + i];
    o += S;
#pragma unroll 4
    for (int d = 0; d < 16; ++d) {
      const float2 v = *qpair(d);
      const int s0 = m.map_i2s[oslot(2 * d, tid)], s1 = m.map_i2s[oslot(2 * d + 1, tid)];
      if (s0 >= 0) o[s0] = v.x;
      if (s1 >= 0) o[s1] = v.y;
    }
  }
  if (em.monitor != nullptr) {
    for (int i = tid; i < S; i += NT) em.monitor[i] = sSCR()[(T - 1) * SCR_PITCH + i];
    if (tid == 0) em.monitor[S] = sSCR()[(T - 1) * SCR_PITCH + NAT_COL];
  }
  if (em.mu != nullptr)
    for (int i = tid; i < S * T; i += NT) { int t = i / S, s = i - t * S; em.mu[i] = sSCR()[t * SCR_PITCH + s]; }

  // level-2 final of this thread (at most one per thread), fetched now so that its L2 latency hides behind the polls
  uint2 t2d = make_uint2(0u, 0u);
  if (tid < m.n_t2) t2d = __ldg(m.t2 + tid);
  // ---------------- P5: polls: linear predictor (stan:95-112), centred binomial_logit (stan:130-131), residuals
  {
    float fsum = 0.f, rnat = 0.f, gm[MAX_MODE] = {0.f, 0.f, 0.f, 0.f}, gp[MAX_MODE] = {0.f, 0.f, 0.f, 0.f};
    const float* scr = sSCR();
    const float* qn = sQNZ();
    const float* pbrow = scr + PB_ROW * SCR_PITCH;
    const bool full = m.full;
    // one poll: linear predictor, centred log-likelihood term f and residual r
    auto poll_eval = [&](int k, uint32_t ix, float& f, float& r, float& sigx) {
      const int s = ix & 63, d = (ix >> 6) & 255, p = (ix >> 14) & 1023, mo = (ix >> 24) & 7, po = (ix >> 27) & 7;
      sigx = (s == NAT_COL) ? m.sig_n : m.sig_s;
      float eta = scr[d * SCR_PITCH + s] + pbrow[s] + m.sig_c * qn[m.nz_c + p] + sigx * qn[m.nz_x + k];
      if (full) {
        eta += m.sig_m * qn[m.nz_m + mo] + m.sig_pop * qn[m.nz_pop + po];
        if ((ix >> 30) & 1) eta += sE()[d];
      }
      poll_term(eta, sPKF(1)[k], sPKF(2)[k], sPKF(3)[k], sPKF(4)[k], f, r);
    };
    auto poll_accum = [&](int k, uint32_t ix, float f, float r, float sigx) {
      const int s = ix & 63, mo = (ix >> 24) & 7, po = (ix >> 27) & 7;
      fsum += f;
      sRR()[k] = r;
      sGNZ()[m.nz_x + k] = sigx * r;
      if (s == NAT_COL) rnat += r;
      if (full) {
#pragma unroll
        for (int j = 0; j < MAX_MODE - 1; ++j) { gm[j] += (mo == j) ? r : 0.f; gp[j] += (po == j) ? r : 0.f; }
        gm[MAX_MODE - 1] += r;  // total; the last class follows by difference
      }
    };
    for (int k = tid; k < m.N; k += NT) {
      const uint32_t ix = sPKI()[k];
      float f, r, sx;
      poll_eval(k, ix, f, r, sx);
      poll_accum(k, ix, f, r, sx);
    }
    PROF(16);
    double v0 = 0.5 * (double)qsq - (double)fsum;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, off);
      rnat += __shfl_xor_sync(0xffffffffu, rnat, off);
    }
    if (full) {
#pragma unroll
      for (int j = 0; j < MAX_MODE; ++j) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          gm[j] += __shfl_xor_sync(0xffffffffu, gm[j], off);
          if (j < MAX_MODE - 1) gp[j] += __shfl_xor_sync(0xffffffffu, gp[j], off);
        }
      }
    }
    if (l == 0) {
      double* red = sRED() + w * 12;
      red[0] = v0; red[1] = (double)rnat;
#pragma unroll
      for (int j = 0; j < MAX_MODE; ++j) { red[2 + j] = (double)gm[j]; red[6 + j] = (double)gp[j]; }
    }
  }
  PROF(17);
  __syncthreads();  // S4
  PROF(4);
  // ---------------- P6: level-1 segment sums of residuals (days, pollsters, states); descriptors live in shared memory
  {
    const float* rr = sRR();
    const uint32_t* t1 = SMP(uint32_t, SM_T1);
    const uint16_t* idl = SMP(uint16_t, SM_IDS);
    for (int i = tid; i < m.n_t1; i += NT) {
      const uint32_t td = t1[i];
      const int start = td & 8191, cnt = (td >> 13) & 31, type = (td >> 18) & 3, slot = td >> 20;
      // loads are unconditional (arrays are padded so that reading SEG entries is always in bounds) and only the
      // accumulation is predicated: all SEG loads issue back to back instead of SEG serialized latencies.  Neighbouring
      // threads hold neighbouring segments of one list (starts differ by SEG = 16 words): each lane walks its segment
      // rotated by its lane id so a warp's 32 accesses fall into 32 different banks instead of 2.  NOTE: the summation
      // order inside a segment therefore depends on the lane, but is fixed for a given task -> still deterministic.
      float acc = 0.f;
      if (type == 1) {
        float v[SEG];
        uint32_t ixs[SEG];
#pragma unroll
        for (int j = 0; j < SEG; ++j) { const int jj = (j + l) & (SEG - 1); v[j] = rr[start + jj]; ixs[j] = sPKI()[start + jj]; }
#pragma unroll
        for (int j = 0; j < SEG; ++j) acc += ((((j + l) & (SEG - 1)) < cnt) && ((ixs[j] >> 30) & 1)) ? v[j] : 0.f;
      } else {
        const uint16_t* ids = idl + start;
        int id[SEG];
#pragma unroll
        for (int j = 0; j < SEG; ++j) id[j] = ids[(j + l) & (SEG - 1)];
        float v[SEG];
#pragma unroll
        for (int j = 0; j < SEG; ++j) v[j] = rr[id[j]];
#pragma unroll
        for (int j = 0; j < SEG; ++j) acc += (((j + l) & (SEG - 1)) < cnt) ? v[j] : 0.f;
      }
      sPSUM()[slot] = acc;
    }
  }
  PROF(18);
  if (w < 10) {  // finalize the block reduction: warp v reduces value v (fixed shuffle tree => deterministic)
    double sv = (l < NWARP) ? sRED()[l * 12 + w] : 0.0;
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) sv += __shfl_xor_sync(0xffffffffu, sv, off);
    if (l == 0) {
      if (w == 0) ctl.U = sv + ctl.u_extra;
      else if (w == 1) ctl.rn_total = (float)sv;
      else sRED()[NWARP * 12 + w] = sv;   // mode / population class sums, combined after barrier S5
    }
  }
  PROF(19);
  // zero the operand planes (scratch reads finished at S4)
  {
    uint4* a4 = SMP(uint4, SM_A);
    const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < (int)(A_REGION / 16 + NT - 1) / NT; ++i) {
      const int k = tid + i * NT;
      if (k < (int)(A_REGION / 16)) a4[k] = z;
    }
  }
  PROF(20);
  __syncthreads();  // S5
  PROF(5);
  // ---------------- P7: G operand cells (direct sums of residuals) and level-2 finals -> pollster gradients / g_e / g_pb row
  {
    const float* rr = sRR();
    const uint32_t* cells = SMP(uint32_t, SM_CELL);
    unsigned char* ahi = smem_raw + SM_A;
    for (int i = tid; i < m.n_cell; i += NT) {
      const uint32_t cd = cells[i];
      const int start = cd & 4095, cnt = (cd >> 12) & 63, t = cd >> 24, s = (cd >> 18) & 63;
      float acc = rr[start];
      for (int j = 1; j < cnt; ++j) acc += rr[start + j];
      __half hi, lo;
      ptx::split_f16(acc * m.scale_G, hi, lo);
      const uint32_t off = (uint32_t)(s >> 3) * A_LBO + (uint32_t)(t >> 3) * A_SBO + (uint32_t)(t & 7) * 16 + (uint32_t)(s & 7) * 2;
      *reinterpret_cast<__half*>(ahi + off) = hi;
      *reinterpret_cast<__half*>(ahi + A_PLANE + off) = lo;
    }
    const float* ps_ = sPSUM();
    if (tid < m.n_t2) {
      const uint2 td = t2d;   // prefetched before the poll phase
      const int ps = td.x & 0xffff, pc = (td.x >> 16) & 0xff, kind = td.x >> 24;
      float acc = 0.f;
#pragma unroll 4
      for (int j = 0; j < pc; ++j) acc += ps_[ps + j];
      if (kind == 0) {
        const int t = td.y >> 6, s = td.y & 63;
        __half hi, lo;
        ptx::split_f16(acc * m.scale_G, hi, lo);
        const uint32_t off = (uint32_t)(s >> 3) * A_LBO + (uint32_t)(t >> 3) * A_SBO + (uint32_t)(t & 7) * 16 + (uint32_t)(s & 7) * 2;
        *reinterpret_cast<__half*>(ahi + off) = hi;
        *reinterpret_cast<__half*>(ahi + A_PLANE + off) = lo;
      } else if (kind == 2) {
        sEBAR()[td.y] = acc;
      } else {
        const float sc = (kind == 1) ? m.sig_c : (kind == 4 ? m.sig_m : m.sig_pop);
        sGNZ()[td.y] = sc * acc;
      }
    }
    if (m.full && tid >= 32 && tid < 32 + 2 * MAX_MODE) {  // mode / population gradients from the class sums
      const int j = (tid - 32) % MAX_MODE;
      const bool ispop = (tid - 32) >= MAX_MODE;
      const int ncls = ispop ? m.Pop : m.M;
      const double* cls = sRED() + NWARP * 12 + (ispop ? 6 : 2);
      const double tot = sRED()[NWARP * 12 + 2 + (MAX_MODE - 1)];   // sum over all polls
      if (j < ncls) {
        double v = cls[j];
        if (j == MAX_MODE - 1) v = tot - sRED()[NWARP * 12 + (ispop ? 6 : 2)] - sRED()[NWARP * 12 + (ispop ? 7 : 3)] - sRED()[NWARP * 12 + (ispop ? 8 : 4)];
        sGNZ()[(ispop ? m.nz_pop : m.nz_m) + j] = (ispop ? m.sig_pop : m.sig_m) * (float)v;
      }
    }
    if (tid == 0) {  // polling-bias row, national K-slot: sum of all national residuals
      __half hi, lo;
      ptx::split_f16(ctl.rn_total * m.scale_G, hi, lo);
      const int t = PB_ROW, s = NAT_COL;
      const uint32_t off = (uint32_t)(s >> 3) * A_LBO + (uint32_t)(t >> 3) * A_SBO + (uint32_t)(t & 7) * 16 + (uint32_t)(s & 7) * 2;
      *reinterpret_cast<__half*>(smem_raw + SM_A + off) = hi;
      *reinterpret_cast<__half*>(smem_raw + SM_A + A_PLANE + off) = lo;
    }
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();  // S6
  PROF(6);
  // ---------------- P8: H^T = G^T X  (B consumed MN-major: the same X planes, transposed view)
  if (tid == 0) issue_gemm(true);
  // overlapped: adjoint of the AR(1) recurrence (warp 1) -> gradients of raw_e_bias, mu_e_bias, rho_e_bias
  if (m.full && w == 1) {
    const float rho = ctl.rho, mu_e = ctl.mu_e, sig_rho = ctl.sig_rho, rterm = ctl.rho_term;
    const float* ze = sQNZ() + m.nz_ze;
    float* gze = sGNZ() + m.nz_ze;
    float A = 1.f, B = 0.f, u[8];  // reversed order: lane l handles t = T-1-8l-j
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int t = T - 1 - (8 * l + j);
      float uj = 0.f, aj = 1.f;
      if (t >= 0) { aj = rho; uj = sEBAR()[t]; }
      u[j] = uj; B = aj * B + uj; A = aj * A;
    }
    affine_scan(A, B, l);
    float ein = __shfl_up_sync(0xffffffffu, B, 1);
    if (l == 0) ein = 0.f;
    float s_mu = 0.f, s_rho = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int t = T - 1 - (8 * l + j);
      if (t >= 0) {
        ein = rho * ein + u[j];  // ebar[t]
        if (t >= 1) {
          gze[t] = sig_rho * ein;
          s_mu += ein;
          s_rho += ein * ((sE()[t - 1] - mu_e) - ze[t] * rterm);
        } else {
          gze[0] = m.sig_e * ein;
        }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { s_mu += __shfl_xor_sync(0xffffffffu, s_mu, off); s_rho += __shfl_xor_sync(0xffffffffu, s_rho, off); }
    if (l == 0) {
      sGNZ()[m.nz_umu] = 0.02f * (1.0f - rho) * s_mu;
      const float d_rho = s_rho - (rho - 0.7f) * 100.0f;
      // stored as (d lp/d u_rho + u_rho) so that the generic "theta - gnz" form yields -d lp/d u_rho
      sGNZ()[m.nz_urho] = rho * (1.0f - rho) * d_rho + (1.0f - 2.0f * rho) + sQNZ()[m.nz_urho];
    }
  }
  PROF(7);
  // ---------------- P9: epilogue 2: H -> scratch
  tmem_epilogue(m.inv_scale_G * (1.0f / 256.0f), false, 1);
  __syncthreads();  // S7
  PROF(8);
  // ---------------- P10: forward cumulative sum of H over days; row 254 = L0^T g_pb = sum_t H[:,t]
  if (zlane) {
    float run0 = 0.f, run1 = 0.f;
    const float* hz = sSCR() + (16 * w) * SCR_PITCH + 2 * l;
#pragma unroll
    for (int d = 0; d < 16; ++d)
      if (d < nd) { const float2 hh = *reinterpret_cast<const float2*>(hz + d * SCR_PITCH); run0 += hh.x; run1 += hh.y; }
    *reinterpret_cast<float2*>(sTOT() + w * 52 + 2 * l) = make_float2(run0, run1);
  }
  if (tid < S) {  // gradient sources of raw_mu_b_T and raw_polling_bias share the spare GEMM row
    const float h = sSCR()[PB_ROW * SCR_PITCH + tid];
    sGNZ()[m.nz_zT + tid] = m.a_T * h;
    sGNZ()[m.nz_zb + tid] = m.a_b * h;
  }
  // odd leaf: the previous leaf's momentum (level-0 merge partner) is fetched from L2 now, so that its latency hides
  // behind barrier S8 and the carry loads instead of stalling the sweep below
  float lrv[EPT];
  if (LEAF) {
    if (lt.Lr != nullptr) {
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float4 t4 = lt.Lr[k * NT]; lrv[4 * k] = t4.x; lrv[4 * k + 1] = t4.y; lrv[4 * k + 2] = t4.z; lrv[4 * k + 3] = t4.w; }
    }
  }
  __syncthreads();  // S8
  PROF(9);
  // ---------------- P11: gradient of U in owner layout
  {
    const uint32_t tp = tpriv();
    float carry0 = 0.f, carry1 = 0.f;
    if (zlane) {
#pragma unroll
      for (int w2 = 0; w2 < NWARP - 1; ++w2) {
        const float2 t2 = *reinterpret_cast<const float2*>(sTOT() + w2 * 52 + 2 * l);
        if (w2 < w) { carry0 += t2.x; carry1 += t2.y; }
      }
    }
    const float* qn = sQNZ() + nz_slot(w, l - ZLANES, 0);
    const float* gn = sGNZ() + nz_slot(w, l - ZLANES, 0);
    const float* qz = sQZ() + (16 * w) * QZ_PITCH + 2 * l;
    const float* hz = sSCR() + (16 * w) * SCR_PITCH + 2 * l;
    float pre0 = carry0, pre1 = carry1;   // running prefix of H over days (this warp's rows)
    float kk = 0.f, c1a = 0.f, c1b = 0.f;
    // 8-element chunks: four operand sets (g, p_half, s, previous momentum) live at once stay inside the register budget
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float g[8], p[8], sm[8], lr[8];
      if (LEAF) {
        ptx::tmem_ld8f(tp + TM_P + 8 * c, p);
        ptx::tmem_ld8f(tp + TM_S + 8 * c, sm);
#pragma unroll
        for (int j = 0; j < 8; ++j) lr[j] = lrv[c * 8 + j];
      }
      if (zlane) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const int d = (c * 8 + j) >> 1;
          const float2 z = *reinterpret_cast<const float2*>(qz + d * QZ_PITCH);
          if (d < nd) { const float2 hh = *reinterpret_cast<const float2*>(hz + d * SCR_PITCH); pre0 += hh.x; pre1 += hh.y; }
          const bool walk = d < ndw;
          g[j] = walk ? fmaf(-m.a_w, pre0, z.x) : z.x;
          g[j + 1] = walk ? fmaf(-m.a_w, pre1, z.y) : z.y;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const int d = (c * 8 + j) >> 1;
          const float2 th = *reinterpret_cast<const float2*>(qn - d * (2 * NZ_LANES));
          const float2 gs = *reinterpret_cast<const float2*>(gn - d * (2 * NZ_LANES));
          g[j] = th.x - gs.x; g[j + 1] = th.y - gs.y;
        }
      }
      __syncwarp();
      if (!LEAF) {
        ptx::tmem_st8f(tp + TM_G + 8 * c, g);
      } else {
        // full-step momentum P = p_half - hs*s*g (kept in TM_TMP for the U-turn merges), next half-step momentum
        // p_half' = 2P - p_half, |P|^2, and -- for an odd leaf -- the level-0 merge with the previous leaf
        ptx::tmem_wait_ld();
        const bool odd = lt.Lr != nullptr;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float P = fmaf(-lt.hs * sm[j], g[j], p[j]);
          kk = fmaf(P, P, kk);
          if (odd) { const float x = lr[j] + P; c1a = fmaf(lr[j], x, c1a); c1b = fmaf(P, x, c1b); }
          p[j] = 2.0f * P - p[j];
          g[j] = P;
        }
        ptx::tmem_st8f(tp + TM_TMP + 8 * c, g);
        ptx::tmem_st8f(tp + TM_P + 8 * c, p);
        if (odd) ptx::tmem_st8f(tp + TM_G + 8 * c, lr);   // running sum of left rhos for the higher merges
        if (lt.Fs != nullptr) {
#pragma unroll
          for (int k = 0; k < 2; ++k) lt.Fs[(2 * c + k) * NT] = make_float4(g[4 * k], g[4 * k + 1], g[4 * k + 2], g[4 * k + 3]);
        }
        if (lt.Es != nullptr) {
#pragma unroll
          for (int k = 0; k < 2; ++k)   // (its r = b + e is re-formed by the level-1 merge)
            lt.Es[(2 * c + k) * NT] = make_float4(g[4 * k], g[4 * k + 1], g[4 * k + 2], g[4 * k + 3]);
        }
      }
    }
    ptx::tmem_wait_st();
    if (LEAF) { lt.kk = kk; lt.c1a = c1a; lt.c1b = c1b; }
  }
  PROF(10);
  // note: callers synchronise before reading ctl.U
}

// out-of-line copy for the cold call sites (initial point of a transition, step-size search, inits, final draw);
// the leaf loop of transition() inlines eval_body so that registers are allocated across the whole loop body
__device__ __noinline__ void eval_point(const Emit em) { LeafTail none{}; eval_body<false>(em, none); }

// ================================================================================================
// small block-wide helpers
// ================================================================================================
// block sum, every thread gets the total (fixed summation order); `buf` alternates between uses that
// are not separated by another barrier
__device__ __forceinline__ float block_sum_f(float v, int buf) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  float* kred = CTL().kred[buf];
  if ((tid & 31) == 0) kred[tid >> 5] = v;
  __syncthreads();
  const float4 a = *reinterpret_cast<const float4*>(kred), b = *reinterpret_cast<const float4*>(kred + 4);
  const float4 c = *reinterpret_cast<const float4*>(kred + 8), d = *reinterpret_cast<const float4*>(kred + 12);
  return (((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w))) + (((c.x + c.y) + (c.z + c.w)) + ((d.x + d.y) + (d.z + d.w)));
}

// momentum refresh: P ~ N(0, I) in whitened coordinates -> TMEM column block `col`; returns |P|^2 partial
__device__ __forceinline__ float draw_momentum(const RunArgs& a, uint32_t tp, uint32_t chain_gid, uint32_t iter, uint32_t stream,
                                               uint32_t sub, uint32_t col) {
  float ss = 0.f;
  const int32_t* map = a.m.map_i2s;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int si = map[oslot(h * 16 + j, threadIdx.x)];
      v[j] = (si >= 0) ? rng_normal(a.seed, chain_gid, (uint32_t)si, iter, stream, sub) : 0.f;
      ss += v[j] * v[j];
    }
    tm_st16(tp, col + 16 * h, v);
  }
  ptx::tmem_wait_st();
  return ss;
}

// Phase A of a leaf: P = p_half - hs*s*g (full-step momentum) -> TM_TMP; returns |P|^2 partial.
__device__ __forceinline__ float full_step_momentum(uint32_t tp, float hs) {
  float ss = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float p[16], s[16], g[16];
    tm_ld16_nowait(tp, TM_P + 16 * h, p);
    tm_ld16_nowait(tp, TM_S + 16 * h, s);
    tm_ld16_nowait(tp, TM_G + 16 * h, g);
    ptx::tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) { p[j] = fmaf(-hs * s[j], g[j], p[j]); ss = fmaf(p[j], p[j], ss); }
    tm_st16(tp, TM_TMP + 16 * h, p);
  }
  ptx::tmem_wait_st();
  return ss;
}
// position update of a leaf: q' = q + eps_signed * s * p_half'   (p_half' was written to TM_P by the fused leaf tail)
__device__ __forceinline__ void advance_q(uint32_t tp, float eps_signed) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float p[16], s[16];
    tm_ld16_nowait(tp, TM_P + 16 * h, p);
    tm_ld16_nowait(tp, TM_S + 16 * h, s);
    ptx::tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      float2* q = qpair((h * 16 + j) >> 1);
      float2 v = *q;
      v.x = fmaf(eps_signed * s[j], p[j], v.x);
      v.y = fmaf(eps_signed * s[j + 1], p[j + 1], v.y);
      *q = v;
    }
  }
}

// One U-turn merge (Stan's three criteria) between the completed left subtree L = {b,e,r} and the
// implicit right subtree R = {b: rb (or P if null), r: P + S, e: P}; S (TM_G) += L.r afterwards.
// `first` : S is implicitly zero.  (Level 0, where L is a single leaf, is fused into the leaf tail.)
// `level1`: both halves are two leaves: L.r = L.b + L.e (the same fp32 sum that would have been stored) and R.b = S
// (the previous leaf, put there by the level-0 merge), so only two vectors come from global memory.
__device__ __forceinline__ bool merge_check(uint32_t tp, const float* Lb, const float* Le, const float* Lr, const float* Rb, bool first,
                                            bool level1 = false) {
  const int tid = threadIdx.x;
  const float4 *Lb4 = own4(Lb), *Le4 = own4(Le), *Lr4 = Lr ? own4(Lr) : nullptr, *Rb4 = Rb ? own4(Rb) : nullptr;
  float c1a = 0.f, c1b = 0.f, c2a = 0.f, c2b = 0.f, c3a = 0.f, c3b = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float P[16], S[16];
    tm_ld16_nowait(tp, TM_TMP + 16 * h, P);
    if (first) {
#pragma unroll
      for (int j = 0; j < 16; ++j) S[j] = 0.f;
    } else {
      tm_ld16_nowait(tp, TM_G + 16 * h, S);
    }
    ptx::tmem_wait_ld();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int gk = (4 * h + k) * NT;
      const float4 lb4 = Lb4[gk], le4 = Le4[gk];
      const float4 lr4 = level1 ? make_float4(0.f, 0.f, 0.f, 0.f) : Lr4[gk];
      const float4 rb4 = (level1 || Rb == nullptr) ? make_float4(0.f, 0.f, 0.f, 0.f) : Rb4[gk];
      const float lbv[4] = {lb4.x, lb4.y, lb4.z, lb4.w}, lev[4] = {le4.x, le4.y, le4.z, le4.w};
      const float lrv[4] = {lr4.x, lr4.y, lr4.z, lr4.w}, rbv[4] = {rb4.x, rb4.y, rb4.z, rb4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = 4 * k + i;
        const float lb = lbv[i], le = lev[i];
        const float lr = level1 ? le + lb : lrv[i];
        const float rb = level1 ? S[j] : ((Rb == nullptr) ? P[j] : rbv[i]);
        const float x = lr + S[j] + P[j];
        c1a = fmaf(lb, x, c1a); c1b = fmaf(P[j], x, c1b);
        const float y = lr + rb;
        c2a = fmaf(lb, y, c2a); c2b = fmaf(rb, y, c2b);
        const float z = S[j] + P[j] + le;
        c3a = fmaf(le, z, c3a); c3b = fmaf(P[j], z, c3b);
        S[j] += lr;
      }
    }
    tm_st16(tp, TM_G + 16 * h, S);
  }
  ptx::tmem_wait_st();
  float v[6] = {c1a, c1b, c2a, c2b, c3a, c3b};
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
  }
  Ctl& ctl = CTL();
  if ((tid & 31) == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) ctl.mred[tid >> 5][i] = v[i];
  }
  __syncthreads();
  bool pred = true;
  if (tid < 6) {
    float s = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < NWARP; ++w2) s += ctl.mred[w2][tid];
    pred = s > 0.f;
  }
  return __syncthreads_and(pred) != 0;
}

__device__ __forceinline__ float logaddexp_f(float a, float b) {
  if (a == -CUDART_INF_F) return b;
  if (b == -CUDART_INF_F) return a;
  const float mx = fmaxf(a, b), mn = fminf(a, b);
  return mx + log1pf(__expf(mn - mx));
}

struct TransStats { float lp, accept, eps, depth, nleap, divergent, energy; };

// ================================================================================================
// one NUTS transition (Stan base_nuts::transition, iterative tree).  On entry q(smem) = current point,
// TM_S = sqrt(inverse metric).  On exit q(smem) = new sample.  `em`: outputs for the ENTRY point
// (the previous iteration's draw), emitted during the initial gradient evaluation.
// ================================================================================================
__device__ __noinline__ void transition(const RunArgs& a, float* ws, uint32_t chain_gid, uint32_t iter, float eps, const Emit em,
                                        TransStats& st) {
  Ctl& ctl = CTL();
  const uint32_t tp = tpriv();
  const int tid = threadIdx.x;
  // fresh whitened momentum into TM_P, gradient at the current point
  float ksq = draw_momentum(a, tp, chain_gid, iter, 1, 0, TM_P);
  eval_point(em);
  ksq = block_sum_f(ksq, 0);  // (its barrier also publishes ctl.U)
  const double U0 = ctl.U;
  const double H0 = U0 + 0.5 * (double)ksq;
  // the two trajectory ends, stored mid-leapfrog: (q +- eps s p_half, p_half), and the tree summary
  {
    const float hs = 0.5f * eps;
    float4* g4 = own4(ws);
    constexpr int V4 = VEC / 4;   // float4s per slot
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float P[16], s[16], gr[16];
      tm_ld16_nowait(tp, TM_P + 16 * h, P);
      tm_ld16_nowait(tp, TM_S + 16 * h, s);
      tm_ld16_nowait(tp, TM_G + 16 * h, gr);
      ptx::tmem_wait_ld();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 qa = *qpair(8 * h + 2 * k), qb = *qpair(8 * h + 2 * k + 1);
        const float qq[4] = {qa.x, qa.y, qb.x, qb.y};
        float pf[4], pb[4], qf[4], qbk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = 4 * k + i;
          pf[i] = fmaf(-hs * s[j], gr[j], P[j]);
          pb[i] = fmaf(hs * s[j], gr[j], P[j]);
          qf[i] = fmaf(eps * s[j], pf[i], qq[i]);
          qbk[i] = fmaf(-eps * s[j], pb[i], qq[i]);
        }
        const int gk = (4 * h + k) * NT;
        const float4 P4 = make_float4(P[4 * k], P[4 * k + 1], P[4 * k + 2], P[4 * k + 3]);
        g4[SLOT_ENDF_P * V4 + gk] = make_float4(pf[0], pf[1], pf[2], pf[3]);
        g4[SLOT_ENDB_P * V4 + gk] = make_float4(pb[0], pb[1], pb[2], pb[3]);
        g4[SLOT_ENDF_Q * V4 + gk] = make_float4(qf[0], qf[1], qf[2], qf[3]);
        g4[SLOT_ENDB_Q * V4 + gk] = make_float4(qbk[0], qbk[1], qbk[2], qbk[3]);
        g4[SLOT_TOP_BB * V4 + gk] = P4;
        g4[SLOT_TOP_FF * V4 + gk] = P4;
        g4[SLOT_TOP_RHO * V4 + gk] = P4;
        g4[SLOT_CAND_A * V4 + gk] = make_float4(qq[0], qq[1], qq[2], qq[3]);
      }
    }
  }
  int samp = SLOT_CAND_A, prop = SLOT_CAND_B;
  if (tid == 0) { ctl.U_samp = U0; ctl.H_samp = H0; ctl.U_prop = U0; ctl.H_prop = H0; ctl.sum_metro = 0.f; }
  float lsw = 0.f;
  int n_leap = 0, depth = 0, loaded = 0;
  bool divergent = false;

  while (depth < a.max_depth) {
    uint32_t rw[4];
    rng_words(a.seed, chain_gid, (uint32_t)depth, iter, 2, 0, rw);
    const int dir = (rw[0] >> 31) ? 1 : -1;  // u > 0.5
    if (loaded != dir) {
      if (loaded != 0) {
        q_to_global(slot_ptr(ws, loaded > 0 ? SLOT_ENDF_Q : SLOT_ENDB_Q));
        tm_to_global(tp, TM_P, slot_ptr(ws, loaded > 0 ? SLOT_ENDF_P : SLOT_ENDB_P));
      }
      global_to_q(slot_ptr(ws, dir > 0 ? SLOT_ENDF_Q : SLOT_ENDB_Q));   // owner-only access: no barrier needed before
      global_to_tm(tp, TM_P, slot_ptr(ws, dir > 0 ? SLOT_ENDF_P : SLOT_ENDB_P));
      loaded = dir;
      ptx::tc_fence_before();
      __syncthreads();
    }
    const float eps_s = dir > 0 ? eps : -eps;
    const float hs = 0.5f * eps_s;
    float lsw_sub = -CUDART_INF_F;
    bool ok = true, persist = true;
    const int nleaf = 1 << depth;
    for (int n = 0; n < nleaf; ++n) {
      const Emit none{nullptr, nullptr, nullptr};
      PROF_DECL;
      // FIRST[z] holds the momentum of the first leaf of every live subtree: leaf m goes to slot z = tz(m)
      // (m = 0: z = depth).  It is not overwritten before leaf m + 2^(z+1), i.e. after every subtree that
      // starts at m has been merged -- so left summaries need no copy of their first momentum.
      auto first_slot = [&](int mleaf) -> float* {
        const int z = mleaf ? (__ffs(mleaf) - 1) : depth;
        return slot_ptr(ws, SLOT_LEFT + 3 * (z - 1));
      };
      const int t = __ffs(~n) - 1;  // trailing ones of n = number of subtrees this leaf completes
      const bool last = (n == nleaf - 1);
      // selection uniforms of the next 256 leaves (stream 3, index = leaf number within the transition): one Philox per leaf
      // computed by one thread each, instead of every thread recomputing the same number at every leaf
      if ((n & 255) == 0) {
        if (n != 0) __syncthreads();   // (the previous table is still being read by slower threads)
        if (tid < 256 && n + tid < nleaf) {
          uint32_t sw[4];
          rng_words(a.seed, chain_gid, (uint32_t)(n_leap + 1 + tid), iter, 3, 0, sw);
          SMP(float, SM_UTAB)[tid] = u01(sw[0]);
        }
      }
      // gradient + momentum update + everything that needs only this thread's elements, in one sweep:
      //   even leaf: its momentum starts subtrees at levels 0..tz(n) -> FIRST slot;
      //   odd leaf : level-0 U-turn sums against the previous leaf; if it closes a level-1 left half (t == 1), its {e, r}
#ifndef POTUS_NO_PREF
      // checkpoints of levels >= 2 were written >= 4 leaves ago and may have left L2 (148 CTAs x 2.4 MB of slots):
      // pull the ones this leaf's merges will read back in while the gradient is computed
      if (tid == 0 && t >= 3) {
        for (int k = 2; k < t; ++k) {
          const float* Lk = slot_ptr(ws, SLOT_LEFT + 3 * (k - 1));
          for (int part = 0; part < 4; ++part) {
            ptx::prefetch_l2_bulk(Lk + VEC + part * (VEC / 4), VEC);
            ptx::prefetch_l2_bulk(Lk + 2 * VEC + part * (VEC / 4), VEC);
            ptx::prefetch_l2_bulk(first_slot(n - (2 << k) + 1) + part * (VEC / 4), VEC);
            ptx::prefetch_l2_bulk(first_slot(n - (1 << k) + 1) + part * (VEC / 4), VEC);
          }
        }
      }
#endif
      LeafTail lt;
      lt.hs = hs;
      lt.Lr = (t > 0) ? own4((const float*)first_slot(n - 1)) : nullptr;
      lt.Fs = (t == 0 && !last) ? own4(first_slot(n)) : nullptr;
      lt.Es = (t == 1 && !last) ? own4(slot_ptr(ws, SLOT_LEFT) + VEC) : nullptr;
      eval_body<true>(none, lt);
      PROF_RESET;
      // one block reduction for |P|^2 and the two level-0 criteria (its barrier also publishes ctl.U)
      float kk = lt.kk, c1a = lt.c1a, c1b = lt.c1b;
      {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          kk += __shfl_xor_sync(0xffffffffu, kk, off);
          c1a += __shfl_xor_sync(0xffffffffu, c1a, off);
          c1b += __shfl_xor_sync(0xffffffffu, c1b, off);
        }
        float* r3 = &ctl.lred[n & 1][0][0];
        if ((tid & 31) == 0) { r3[tid >> 5] = kk; r3[NWARP + (tid >> 5)] = c1a; r3[2 * NWARP + (tid >> 5)] = c1b; }
        __syncthreads();
        float tot[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float4 a4 = *reinterpret_cast<const float4*>(r3 + i * NWARP), b4 = *reinterpret_cast<const float4*>(r3 + i * NWARP + 4);
          const float4 c4 = *reinterpret_cast<const float4*>(r3 + i * NWARP + 8), d4 = *reinterpret_cast<const float4*>(r3 + i * NWARP + 12);
          tot[i] = (((a4.x + a4.y) + (a4.z + a4.w)) + ((b4.x + b4.y) + (b4.z + b4.w))) +
                   (((c4.x + c4.y) + (c4.z + c4.w)) + ((d4.x + d4.y) + (d4.z + d4.w)));
        }
        kk = tot[0]; c1a = tot[1]; c1b = tot[2];
      }
      PROF(11);
      double h = ctl.U + 0.5 * (double)kk;
      if (!(h == h)) h = CUDART_INF;
      ++n_leap;
      const float dH = (float)(H0 - h);  // energy differences are O(1): fp32 is ample for the weights
      if (h - H0 > 1000.0) divergent = true;
      lsw_sub = logaddexp_f(lsw_sub, dH);
      if (tid == 0) ctl.sum_metro += (dH > 0.f) ? 1.0f : __expf(dH);
      if (divergent) { ok = false; break; }
      // multinomial selection inside the new subtree (reservoir form of Stan's pairwise rule)
      {
        // (the table entry was written before this leaf's gradient, i.e. behind at least eight barriers)
        if (n == 0 || SMP(float, SM_UTAB)[n & 255] < __expf(dH - lsw_sub)) {
          q_to_global(slot_ptr(ws, prop));
          if (tid == 0) { ctl.U_prop = ctl.U; ctl.H_prop = h; }
        }
      }
      PROF(12);
      // position update q' = q + eps*s*p_half' (after the candidate copy above).  No barrier follows: this is safe only
      // because, until barrier S1 of the next gradient, every thread reads nothing of q but the elements it owns.
      // (Fusing this pass into the leaf tail -- writing every leaf's position to a spare candidate slot so that selection
      //  becomes a pointer swap -- was built and measured in round 2: bit-identical chains, 5 % SLOWER (the eight extra
      //  128-bit global stores per thread per leaf cost more than the two TMEM re-reads they save); profiles/r02_ab_resident_fusion.log)
      advance_q(tp, eps_s);
      PROF(15);
      // U-turn checks for every subtree this leaf completes (level 0 came with the reduction above)
      if (t > 0) ok = (c1a > 0.f) && (c1b > 0.f);
      if (t > 1 && ok) ok = merge_check(tp, first_slot(n - 3), slot_ptr(ws, SLOT_LEFT) + VEC, nullptr, nullptr, false, true);
      for (int k = 2; k < t && ok; ++k) {
        const float* Lk = slot_ptr(ws, SLOT_LEFT + 3 * (k - 1));
        ok = merge_check(tp, first_slot(n - (2 << k) + 1), Lk + VEC, Lk + 2 * VEC, first_slot(n - (1 << k) + 1), false);
      }
      if (!ok) break;
      PROF(13);
      if (!last) {
        if (t >= 2) {
          // this subtree becomes the stored left half at level t: {e, r} (its b is FIRST[...]); t == 1 was stored above
          float* Lt = slot_ptr(ws, SLOT_LEFT + 3 * (t - 1));
          float4 *e4 = own4(Lt + VEC), *r4 = own4(Lt + 2 * VEC);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            float P[16], S[16];
            tm_ld16_nowait(tp, TM_TMP + 16 * hh, P);
            tm_ld16_nowait(tp, TM_G + 16 * hh, S);
            ptx::tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) S[j] += P[j];
            st_groups4(e4, 4 * hh, P, 16);
            st_groups4(r4, 4 * hh, S, 16);
          }
        }
      } else {
        // last leaf: merge the finished subtree with the existing trajectory (top level of
        // base_nuts::transition) and fold it into the tree summary
        const float* F = slot_ptr(ws, dir > 0 ? SLOT_TOP_BB : SLOT_TOP_FF);
        const float* A = slot_ptr(ws, dir > 0 ? SLOT_TOP_FF : SLOT_TOP_BB);
        const float* Rb = (depth == 0) ? nullptr : first_slot(0);
        persist = merge_check(tp, F, A, slot_ptr(ws, SLOT_TOP_RHO), Rb, depth == 0);
        float4* rho4 = own4(slot_ptr(ws, SLOT_TOP_RHO));
        float4* end4 = own4(slot_ptr(ws, dir > 0 ? SLOT_TOP_FF : SLOT_TOP_BB));
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float P[16], S[16];
          tm_ld16_nowait(tp, TM_TMP + 16 * hh, P);
          tm_ld16_nowait(tp, TM_G + 16 * hh, S);  // = old rho_top + sum of lower lefts
          ptx::tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) S[j] += P[j];
          st_groups4(rho4, 4 * hh, S, 16);
          st_groups4(end4, 4 * hh, P, 16);
        }
      }
      PROF(14);
      if (threadIdx.x == 0) { PROF_COUNT; }
    }
    if (!ok) break;
    ++depth;
    // biased progressive sampling at the top level
    if (lsw_sub > lsw || u01(rw[1]) < __expf(lsw_sub - lsw)) {
      const int tmp = samp; samp = prop; prop = tmp;
      if (tid == 0) { ctl.U_samp = ctl.U_prop; ctl.H_samp = ctl.H_prop; }
    }
    lsw = logaddexp_f(lsw, lsw_sub);
    if (!persist) break;
  }
  global_to_q(slot_ptr(ws, samp));
  ptx::tc_fence_before();
  __syncthreads();
  // (thread 0 wrote the statistics before the barrier above; every thread returns the same values)
  st.lp = (float)(-ctl.U_samp);  // centred; the host adds lp_const in fp64 (fp32 cannot hold -1.2e6 to 1e-2)
  st.accept = ctl.sum_metro / (float)(n_leap > 0 ? n_leap : 1);
  st.eps = eps; st.depth = (float)depth; st.nleap = (float)n_leap; st.divergent = divergent ? 1.f : 0.f;
  st.energy = (float)ctl.H_samp;
  if (tid == 0) { ctl.cs.n_leapfrog += n_leap; ctl.cs.U = (float)ctl.U_samp; }
}

// Stan base_hmc::init_stepsize: double / halve eps until the one-step acceptance crosses 0.8
__device__ __noinline__ float find_stepsize(const RunArgs& a, float* ws, uint32_t chain_gid, uint32_t iter_tag, float eps) {
  Ctl& ctl = CTL();
  const uint32_t tp = tpriv();
  if (!(eps > 0.f) || eps > 1e7f) return eps;
  q_to_global(slot_ptr(ws, SLOT_TMPQ));
  int direction = 0;
  const Emit none{nullptr, nullptr, nullptr};
  for (uint32_t attempt = 0; attempt < 200; ++attempt) {
    __syncthreads();
    global_to_q(slot_ptr(ws, SLOT_TMPQ));
    float k0 = draw_momentum(a, tp, chain_gid, iter_tag, 5, attempt, TM_P);
    ptx::tc_fence_before();
    __syncthreads();
    eval_point(none);
    k0 = block_sum_f(k0, 0);
    const double H0 = ctl.U + 0.5 * (double)k0;
    // first half step + position step
    full_step_momentum(tp, 0.5f * eps);  // TM_TMP = p - eps/2 s g  (= p_half)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float p[16], s[16];
      tm_ld16_nowait(tp, TM_TMP + 16 * h, p);
      tm_ld16_nowait(tp, TM_S + 16 * h, s);
      ptx::tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        float2* q = qpair((h * 16 + j) >> 1);
        float2 v = *q;
        v.x = fmaf(eps * s[j], p[j], v.x);
        v.y = fmaf(eps * s[j + 1], p[j + 1], v.y);
        *q = v;
      }
      tm_st16(tp, TM_P + 16 * h, p);
    }
    ptx::tmem_wait_st();
    ptx::tc_fence_before();
    __syncthreads();
    eval_point(none);
    float k1 = full_step_momentum(tp, 0.5f * eps);
    k1 = block_sum_f(k1, 1);
    double h = ctl.U + 0.5 * (double)k1;
    if (!(h == h)) h = CUDART_INF;
    const double dH = H0 - h;
    const double thr = log(0.8);
    if (attempt == 0) { direction = dH > thr ? 1 : -1; continue; }
    if (direction == 1 && !(dH > thr)) break;
    if (direction == -1 && !(dH < thr)) break;
    eps = direction == 1 ? 2.0f * eps : 0.5f * eps;
    if (eps > 1e7f || eps == 0.f) break;
  }
  __syncthreads();
  global_to_q(slot_ptr(ws, SLOT_TMPQ));
  ptx::tc_fence_before();
  __syncthreads();
  return eps;
}

// ================================================================================================
// kernels
// ================================================================================================
__device__ __forceinline__ void cta_setup(const ModelDev& mg) {
  const int tid = threadIdx.x, w = tid >> 5;
  Ctl& ctl = CTL();
  if (w == 0) { ptx::tmem_alloc(&ctl.tmem_base, 512); ptx::tmem_relinquish(); }
  if (tid == 0) {
    ptx::mbar_init(&ctl.bar_mma[0], 1);
    ptx::mbar_init(&ctl.bar_mma[1], 1);
    ptx::mbar_init(&ctl.bar_load, 1);
    ptx::fence_mbar_init();
  }
  // zero every padded vector in shared memory; copy the model block
  for (int i = tid; i < (int)((SM_CTL - SM_QZ) / 4); i += NT) SMP(float, SM_QZ)[i] = 0.f;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&mg);
    for (int i = tid; i < (int)(sizeof(ModelDev) / 4); i += NT) SMP(uint32_t, SM_MODEL)[i] = src[i];
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  // constants: TMA bulk copies (global -> shared), one mbarrier
  if (tid == 0) {
    const uint32_t bytes = 2 * B_PLANE + 5 * NPOLL_CAP * 4 + 64 * 4 + NT1_CAP * 4 + NCELL_CAP * 4 + NIDS_CAP * 2;
    ptx::mbar_expect_tx(&ctl.bar_load, bytes);
    ptx::bulk_g2s(smem_raw + SM_T1, mg.t1, NT1_CAP * 4, &ctl.bar_load);
    ptx::bulk_g2s(smem_raw + SM_CELL, mg.cells, NCELL_CAP * 4, &ctl.bar_load);
    ptx::bulk_g2s(smem_raw + SM_IDS, mg.ids, NIDS_CAP * 2, &ctl.bar_load);
    ptx::bulk_g2s(smem_raw + SM_B, mg.btiles, 2 * B_PLANE, &ctl.bar_load);
    ptx::bulk_g2s(smem_raw + SM_PK, mg.pk, 5 * NPOLL_CAP * 4, &ctl.bar_load);
    ptx::bulk_g2s(smem_raw + SM_PRIOR, mg.prior, 64 * 4, &ctl.bar_load);
  }
  ptx::mbar_wait(&ctl.bar_load, 0);
  // zero the resident TMEM vectors (padding slots must hold zeros)
  {
    float z[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) z[j] = 0.f;
    const uint32_t tp = tpriv();
    tm_st16(tp, TM_P, z); tm_st16(tp, TM_P + 16, z); tm_st16(tp, TM_S, z); tm_st16(tp, TM_S + 16, z);
    ptx::tmem_wait_st();
  }
  ptx::tc_fence_before();
  __syncthreads();
}
__device__ __forceinline__ void cta_teardown() {
  ptx::tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) ptx::tmem_dealloc(CTL().tmem_base, 512);
}

// test hook: log density + gradient for n positions (potus_logp_grad)
extern "C" __global__ void __launch_bounds__(NT, 1) potus_eval_kernel(const __grid_constant__ EvalArgs a) {
  cta_setup(a.m);
  const uint32_t tp = tpriv();
  for (int i = blockIdx.x; i < a.n; i += gridDim.x) {
    global_to_q(a.q_in + (size_t)i * VEC);
    __syncthreads();
    const Emit em{nullptr, nullptr, a.mu_out ? a.mu_out + (size_t)i * a.m.S * a.m.T : nullptr};
    eval_point(em);
    __syncthreads();
    tm_to_global(tp, TM_G, a.g_out + (size_t)i * VEC);
    if (threadIdx.x == 0) a.u_out[i] = CTL().U;
    ptx::tc_fence_before();
    __syncthreads();
  }
  cta_teardown();
}

extern "C" __global__ void __launch_bounds__(NT, 1) potus_nuts_kernel(const __grid_constant__ RunArgs a) {
  cta_setup(a.m);
  const ModelDev& m = MD();
  Ctl& ctl = CTL();
  const int tid = threadIdx.x;
  const uint32_t tp = tpriv();
  float* ws = a.workspace + (size_t)blockIdx.x * NSLOT * VEC;
  const int n_iter_total = a.iter_warmup + a.iter_sampling;
  const Emit none{nullptr, nullptr, nullptr};

#ifdef POTUS_PROF
  if (tid < 40) ctl.prof[tid] = 0ull;
#endif
  for (;;) {
    __syncthreads();
    if (tid == 0) ctl.chain = atomicAdd(a.queue, 1);
    __syncthreads();
    const int chain = ctl.chain;
    if (chain >= a.n_chains) break;
    const uint32_t gid = (uint32_t)(a.chain_id_offset + chain);
    float* qg = a.q + (size_t)chain * VEC;
    float* sg = a.sqrt_m + (size_t)chain * VEC;
    float* wmean = a.wf_mean + (size_t)chain * VEC;   // owner layout: element e of this thread at oslot(e, tid)
    float* wm2 = a.wf_m2 + (size_t)chain * VEC;
    const int32_t* map = a.m.map_i2s;
    if (tid == 0) ctl.cs = a.cs[chain];
    __syncthreads();

    if (a.do_init) {
      // unit metric (zero in padding slots so that padding never moves)
      {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float one[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) one[j] = (map[oslot(h * 16 + j, tid)] >= 0) ? 1.0f : 0.f;
          tm_st16(tp, TM_S + 16 * h, one);
        }
        ptx::tmem_wait_st();
      }
      // ---- random inits U(-r, r) on the unconstrained scale; retry while U / gradient are not finite
      bool good = false;
      for (uint32_t attempt = 0; attempt < 100 && !good; ++attempt) {
#pragma unroll 2
        for (int d = 0; d < 16; ++d) {
          float2 v = make_float2(0.f, 0.f);
          const int s0 = map[oslot(2 * d, tid)], s1 = map[oslot(2 * d + 1, tid)];
          uint32_t rw[4];
          if (s0 >= 0) { rng_words(a.seed, gid, (uint32_t)s0, 0, 0, attempt, rw); v.x = a.init_radius * (2.0f * u01(rw[0]) - 1.0f); }
          if (s1 >= 0) { rng_words(a.seed, gid, (uint32_t)s1, 0, 0, attempt, rw); v.y = a.init_radius * (2.0f * u01(rw[0]) - 1.0f); }
          *qpair(d) = v;
        }
        ptx::tc_fence_before();
        __syncthreads();
        eval_point(none);
        int bad = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float g[16];
          tm_ld16(tp, TM_G + 16 * h, g);
#pragma unroll
          for (int j = 0; j < 16; ++j) bad |= !isfinite(g[j]);
        }
        __syncthreads();
        if (!isfinite(ctl.U)) bad = 1;
        good = __syncthreads_or(bad) == 0;
      }
      if (tid == 0) {
        ChainState& cs = ctl.cs;
        cs.status = good ? 0 : -1;
        cs.eps = 1.0f; cs.da_counter = 0; cs.da_sbar = 0; cs.da_xbar = 0;
        cs.w_counter = 0; cs.w_size = a.w_base_window; cs.w_nsamp = 0;
        // Stan's windowed_adaptation leaves adapt_next_window_ at UINT_MAX for num_warmup < 20: no metric update at all
        cs.w_next = a.w_base_window > 0 ? a.w_init_buffer + a.w_base_window - 1 : -1;
        cs.iter = 0; cs.n_leapfrog = 0;
      }
#pragma unroll 8
      for (int e = 0; e < EPT; ++e) { wmean[oslot(e, tid)] = 0.f; wm2[oslot(e, tid)] = 0.f; }
      __syncthreads();
      const float e0 = find_stepsize(a, ws, gid, 0xFFFFFFFFu, 1.0f);
      if (tid == 0) { ctl.cs.eps = e0; ctl.cs.da_mu = log(10.0 * (double)e0); }
      __syncthreads();
    } else {
      global_to_q(qg);
      global_to_tm(tp, TM_S, sg);
      ptx::tc_fence_before();
      __syncthreads();
    }

    // ---- iterations
    for (int it = a.iter_begin; it < a.iter_end; ++it) {
      const float eps = ctl.cs.eps;
      // outputs of the ENTRY point = draw of iteration it-1 (if it was a sampling iteration)
      Emit em = none;
      const int kprev = it - 1 - a.iter_warmup;
      if (kprev >= 0) {
        em.monitor = a.monitor + ((size_t)chain * a.iter_sampling + kprev) * (m.S + 1);
        if (a.keep_per_chain > 0 && (kprev % a.keep_every) == 0) {
          const int slot = kprev / a.keep_every;
          if (slot < a.keep_per_chain) em.draw = a.draws + ((size_t)chain * a.keep_per_chain + slot) * a.draw_len;
        }
      }
      TransStats st;
      transition(a, ws, gid, (uint32_t)it, eps, em, st);
      if (tid == 0) {
        float* sp = a.sampler_params + ((size_t)chain * n_iter_total + it) * 8;
        sp[0] = st.lp; sp[1] = st.accept; sp[2] = st.eps; sp[3] = st.depth; sp[4] = st.nleap; sp[5] = st.divergent; sp[6] = st.energy; sp[7] = 0.f;
      }
      if (it < a.iter_warmup) {
        // ---- Stan stepsize_adaptation::learn_stepsize (dual averaging)
        if (tid == 0) {
          ChainState& cs = ctl.cs;
          cs.da_counter++;
          const double as = st.accept > 1.f ? 1.0 : (double)st.accept;
          const double eta = 1.0 / (cs.da_counter + 10.0);
          cs.da_sbar = (1.0 - eta) * cs.da_sbar + eta * ((double)a.adapt_delta - as);
          const double x = cs.da_mu - cs.da_sbar * sqrt((double)cs.da_counter) / 0.05;
          const double xe = pow((double)cs.da_counter, -0.75);
          cs.da_xbar = (1.0 - xe) * cs.da_xbar + xe * x;
          cs.eps = (float)exp(x);
        }
        __syncthreads();
        // ---- Stan var_adaptation::learn_variance with windowed_adaptation
        const int wc = ctl.cs.w_counter;
        const bool in_window = wc >= a.w_init_buffer && wc < a.iter_warmup - a.w_term_buffer && wc != a.iter_warmup;
        const bool end_window = wc == ctl.cs.w_next && wc != a.iter_warmup;
        int nsamp = ctl.cs.w_nsamp;
        if (in_window) {  // Welford update, owner layout (padding: q = 0 keeps mean = m2 = 0)
          ++nsamp;
          const float inv = 1.0f / (float)nsamp;
#pragma unroll 4
          for (int d = 0; d < 16; ++d) {
            const float2 q = *qpair(d);
#pragma unroll
            for (int b = 0; b < 2; ++b) {
              const int gi = oslot(2 * d + b, tid);
              const float qq = b ? q.y : q.x;
              const float mu = wmean[gi], dd = qq - mu, mu2 = fmaf(dd, inv, mu);
              wmean[gi] = mu2;
              wm2[gi] = fmaf(qq - mu2, dd, wm2[gi]);
            }
          }
        }
        __syncthreads();
        if (end_window) {
          const float n = (float)nsamp;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float s[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int gi = oslot(h * 16 + j, tid);
              float v = 0.f;
              if (map[gi] >= 0) {
                const float var = wm2[gi] / (n - 1.0f);
                v = sqrtf((n / (n + 5.0f)) * var + 1e-3f * (5.0f / (n + 5.0f)));
              }
              s[j] = v;
              wmean[gi] = 0.f; wm2[gi] = 0.f;
            }
            tm_st16(tp, TM_S + 16 * h, s);
          }
          ptx::tmem_wait_st();
          if (tid == 0) {
            ChainState& cs = ctl.cs;
            const int last = a.iter_warmup - a.w_term_buffer - 1;
            if (cs.w_next != last) {
              cs.w_size *= 2;
              cs.w_next = cs.w_counter + cs.w_size;
              if (cs.w_next != last) {
                const int boundary = cs.w_next + 2 * cs.w_size;
                if (boundary >= a.iter_warmup - a.w_term_buffer) cs.w_next = last;
              }
            }
          }
          nsamp = 0;
          ptx::tc_fence_before();
          __syncthreads();
          const float e1 = find_stepsize(a, ws, gid, (uint32_t)it, ctl.cs.eps);
          if (tid == 0) {
            ChainState& cs = ctl.cs;
            cs.eps = e1; cs.da_mu = log(10.0 * (double)e1); cs.da_counter = 0; cs.da_sbar = 0; cs.da_xbar = 0;
          }
        }
        if (tid == 0) {
          ctl.cs.w_nsamp = nsamp;
          ctl.cs.w_counter = wc + 1;
          if (it == a.iter_warmup - 1) ctl.cs.eps = (float)exp(ctl.cs.da_xbar);
        }
        __syncthreads();
      }
    }
    // ---- the last iteration's draw needs one more evaluation at the final point
    if (a.iter_end == n_iter_total && a.iter_end > a.iter_begin) {
      Emit em = none;
      const int kprev = n_iter_total - 1 - a.iter_warmup;
      if (kprev >= 0) {
        em.monitor = a.monitor + ((size_t)chain * a.iter_sampling + kprev) * (m.S + 1);
        if (a.keep_per_chain > 0 && (kprev % a.keep_every) == 0) {
          const int slot = kprev / a.keep_every;
          if (slot < a.keep_per_chain) em.draw = a.draws + ((size_t)chain * a.keep_per_chain + slot) * a.draw_len;
        }
        eval_point(em);
        __syncthreads();
      }
    }
    // ---- persist the chain
    q_to_global(qg);
    tm_to_global(tp, TM_S, sg);
    if (tid == 0) { ctl.cs.iter = a.iter_end; a.cs[chain] = ctl.cs; }
    ptx::tc_fence_before();
    __syncthreads();
  }
#ifdef POTUS_PROF
  if (a.prof != nullptr && tid < 40) atomicAdd(a.prof + tid, ctl.prof[tid]);
#endif
  cta_teardown();
}

}  // namespace potus
