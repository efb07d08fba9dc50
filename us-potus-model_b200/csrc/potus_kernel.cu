// Resident NUTS kernel for poll_model_2020 (sm_100a).  One CTA = one chain; the chain's position
// lives in shared memory, its half-step momentum and sqrt(metric) in TMEM (thread-private columns),
// and every leapfrog is:  scan -> tcgen05 GEMM (L0 W) -> poll residuals -> tcgen05 GEMM (L0^T G)
// -> scan -> momentum/position update + NUTS bookkeeping, without touching HBM for the state.
//
// What it restates (reference file:line; the sampler itself is Stan 2.24.1, not in the tree):
//   poll_model_2020.stan:70-113  transformed parameters   -> eval_point() forward half
//   poll_model_2020.stan:115-132 log density               -> eval_point() energy + hand-derived gradient
//   poll_model_2020.stan:134-140 generated quantities      -> emit_draw()
//   Stan base_nuts::transition / build_tree (multinomial NUTS, generalised U-turn with the two
//   extra sub-tree checks), expl_leapfrog, diag_e_metric, stepsize/var/windowed adaptation,
//   init_stepsize -> transition(), adapt_*(), find_stepsize().
// The CPU statement of exactly this algorithm (iterative tree, Philox streams) is
// oracle/potus_oracle.c (tree_mode 1); tests compare the two.
#include <cuda_runtime.h>
#include <math_constants.h>
#include "ptx_sm100.cuh"
#include "potus_layout.h"

namespace potus {

struct Ctl {
  uint64_t bar_mma[2];
  uint64_t bar_load;
  uint32_t tmem_base;
  int chain;
  double U;        // potential energy of the last evaluated point (centred: -lp + lp_const)
  double u_extra;  // -(log-density terms of rho_e_bias)
  float rho, mu_e, sig_rho, rho_term;  // rho_term = sig_e*rho/sqrt(1-rho^2)
  float rn_total;
  float kred[16];
  float mred[16][8];
  int nonfinite;
  ChainState cs;
};

struct Smem {
  unsigned char* a;      // operand planes / fp32 scratch
  unsigned char* b;      // X hi, X lo
  float* qz;             // [254][52]
  float* qnz;            // [NZ_CAP]
  float* gnz;            // [NZ_CAP]
  uint32_t* pk_idx;
  float *pk_n, *pk_eta, *pk_p, *pk_rho;
  float* rr;
  float* psum;
  float *e, *ebar;
  float* tot;
  float* prior;
  double* red;
  Ctl* ctl;
  float* scr;            // = (float*)a
};

__device__ __forceinline__ Smem carve(unsigned char* base) {
  Smem s;
  s.a = base + SM_A;
  s.b = base + SM_B;
  s.qz = reinterpret_cast<float*>(base + SM_QZ);
  s.qnz = reinterpret_cast<float*>(base + SM_QNZ);
  s.gnz = reinterpret_cast<float*>(base + SM_GNZ);
  s.pk_idx = reinterpret_cast<uint32_t*>(base + SM_PK);
  s.pk_n = reinterpret_cast<float*>(base + SM_PK + 1 * NPOLL_CAP * 4);
  s.pk_eta = reinterpret_cast<float*>(base + SM_PK + 2 * NPOLL_CAP * 4);
  s.pk_p = reinterpret_cast<float*>(base + SM_PK + 3 * NPOLL_CAP * 4);
  s.pk_rho = reinterpret_cast<float*>(base + SM_PK + 4 * NPOLL_CAP * 4);
  s.rr = reinterpret_cast<float*>(base + SM_RR);
  s.psum = reinterpret_cast<float*>(base + SM_PSUM);
  s.e = reinterpret_cast<float*>(base + SM_E);
  s.ebar = s.e + 256;
  s.tot = reinterpret_cast<float*>(base + SM_TOT);
  s.prior = reinterpret_cast<float*>(base + SM_PRIOR);
  s.red = reinterpret_cast<double*>(base + SM_RED);
  s.ctl = reinterpret_cast<Ctl*>(base + SM_CTL);
  s.scr = reinterpret_cast<float*>(base + SM_A);
  return s;
}

struct TC {  // per-thread constants
  int tid, w, l;
  bool zlane, nzlane;
  uint32_t tpriv;  // TMEM address of this thread's private 32-column window (column offset 0)
  int nz0;         // first nz slot owned (nz lanes)
  uint32_t ph;     // mbarrier phase parity (both MMA barriers flip once per GEMM)
};

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
__device__ __forceinline__ void tm_ld16(const TC& tc, uint32_t col, float (&v)[16]) {
  uint32_t u[16];
  ptx::tmem_ld16(tc.tpriv + col, u);
  ptx::tmem_wait_ld();
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(u[j]);
}
__device__ __forceinline__ void tm_st16(const TC& tc, uint32_t col, const float (&v)[16]) {
  uint32_t u[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) u[j] = __float_as_uint(v[j]);
  ptx::tmem_st16(tc.tpriv + col, u);
}
__device__ __forceinline__ float* slot_ptr(float* ws, int slot) { return ws + (size_t)slot * VEC; }

// copy: TMEM private vector <-> global
__device__ __forceinline__ void tm_to_global(const TC& tc, uint32_t col, float* g) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float v[16];
    tm_ld16(tc, col + 16 * h, v);
#pragma unroll
    for (int j = 0; j < 16; ++j) g[(h * 16 + j) * NT + tc.tid] = v[j];
  }
}
__device__ __forceinline__ void global_to_tm(const TC& tc, uint32_t col, const float* g) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = g[(h * 16 + j) * NT + tc.tid];
    tm_st16(tc, col + 16 * h, v);
  }
  ptx::tmem_wait_st();
}
// position element access in shared memory for owned element e
__device__ __forceinline__ float* q_elem(const Smem& sm, const TC& tc, int e) {
  if (tc.zlane) return sm.qz + (16 * tc.w + (e >> 1)) * QZ_PITCH + 2 * tc.l + (e & 1);
  return sm.qnz + tc.nz0 + e;
}
// is owned element e a real parameter (not padding)?
__device__ __forceinline__ bool elem_valid(const ModelDev& m, const TC& tc, int e) {
  if (tc.zlane) return (16 * tc.w + (e >> 1)) < m.T && (2 * tc.l + (e & 1)) < m.S;
  if (tc.nzlane) return tc.nz0 + e < m.NZ;
  return false;
}
__device__ __forceinline__ void q_to_global(const ModelDev& m, const Smem& sm, const TC& tc, float* g) {
#pragma unroll
  for (int e = 0; e < EPT; ++e) g[e * NT + tc.tid] = elem_valid(m, tc, e) ? *q_elem(sm, tc, e) : 0.0f;
}
__device__ __forceinline__ void global_to_q(const ModelDev& m, const Smem& sm, const TC& tc, const float* g) {
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    float v = g[e * NT + tc.tid];
    if (elem_valid(m, tc, e)) *q_elem(sm, tc, e) = v;
  }
}

// ------------------------------------------------------------------------------------------------
// affine-map warp scan helpers (AR(1) recurrences)
// ------------------------------------------------------------------------------------------------
// inclusive scan of maps x -> A x + B, composition "later after earlier"; lane order ascending
__device__ __forceinline__ void affine_scan(float& A, float& B, int l) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    float Ap = __shfl_up_sync(0xffffffffu, A, off), Bp = __shfl_up_sync(0xffffffffu, B, off);
    if (l >= off) { B = A * Bp + B; A = A * Ap; }
  }
}


// Accumulator read-out shared by both GEMMs: (D1 + D2/2048) * scale (+ prior for day rows) -> fp32
// scratch [row][53].  The scratch aliases the operand planes, so nothing is written before the LAST
// commit (bar_mma[1]) has completed; tile-0 warps still overlap their TMEM loads with tile 1's MMAs.
__device__ __forceinline__ void tmem_epilogue(const ModelDev& m, const Smem& sm, const TC& tc, float scale, bool add_prior) {
  Ctl* ctl = sm.ctl;
  const int w = tc.w, l = tc.l, T = m.T;
  const int g = w >> 2, tile = g >> 1, half = g & 1, qd = w & 3;
  ptx::mbar_wait(&ctl->bar_mma[tile], tc.ph);
  ptx::tc_fence_after();
  const int row = tile * 128 + qd * 32 + l;
  const uint32_t taddr = ctl->tmem_base + ((uint32_t)(qd * 32) << 16) + tile * 64 + half * 32;
  const bool rowok = (row < T) || (row == PB_ROW);
  uint32_t d1[16], d2[16];
  ptx::tmem_ld16(taddr + TM_D1, d1);
  ptx::tmem_ld16(taddr + TM_D2, d2);
  ptx::tmem_wait_ld();
  if (tile == 0) ptx::mbar_wait(&ctl->bar_mma[1], tc.ph);
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    if (cc == 1) {
      ptx::tmem_ld16(taddr + TM_D1 + 16, d1);
      ptx::tmem_ld16(taddr + TM_D2 + 16, d2);
      ptx::tmem_wait_ld();
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int col = half * 32 + cc * 16 + j;
      if (col < 52 && rowok) {
        float v = (__uint_as_float(d1[j]) + __uint_as_float(d2[j]) * (1.0f / 2048.0f)) * scale;
        if (add_prior && row < T) v += sm.prior[col];
        sm.scr[row * SCR_PITCH + col] = v;
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

// ================================================================================================
// eval_point: potential U (-> ctl->U) and its gradient (-> TMEM TM_G, owner layout) at the position
// held in shared memory.  Must be called by all 512 threads.
// ================================================================================================
__device__ __noinline__ void eval_point(const ModelDev& m, const Smem& sm, TC& tc, const Emit em) {
  const int T = m.T, S = m.S, w = tc.w, l = tc.l, tid = tc.tid;
  Ctl* ctl = sm.ctl;
  float qsq = 0.f;  // sum of squares of owned parameters (prior energy)

  // ---------------- P1: reverse scan of the walk innovations (poll_model_2020.stan:86 collapsed)
  float c[32];
  {
    float run0 = 0.f, run1 = 0.f;
    if (tc.zlane) {
#pragma unroll
      for (int d = 15; d >= 0; --d) {
        const int t = 16 * w + d;
        float2 z = make_float2(0.f, 0.f);
        if (t < T) z = *reinterpret_cast<const float2*>(sm.qz + t * QZ_PITCH + 2 * l);
        qsq += z.x * z.x + z.y * z.y;
        if (t <= T - 2) { run0 += z.x; run1 += z.y; }
        c[2 * d] = run0; c[2 * d + 1] = run1;
      }
      sm.tot[w * 52 + 2 * l] = run0;
      sm.tot[w * 52 + 2 * l + 1] = run1;
    } else if (tc.nzlane) {
#pragma unroll
      for (int e = 0; e < EPT; ++e) { float v = sm.qnz[tc.nz0 + e]; qsq += v * v; }
      if (m.full) {  // rho's unconstrained value carries no N(0,1) term
        int k = m.nz_urho - tc.nz0;
        if (k >= 0 && k < EPT) { float v = sm.qnz[m.nz_urho]; qsq -= v * v; }
      }
    }
  }
  __syncthreads();  // S1
  // ---------------- P2: W -> fp16 hi/lo operand planes (K-major, SWIZZLE_NONE)
  {
    float carry0 = 0.f, carry1 = 0.f, zt0 = 0.f, zt1 = 0.f, zb0 = 0.f, zb1 = 0.f;
    const bool act = tc.zlane && l < m.npair;
    if (act) {
      for (int w2 = w + 1; w2 < NWARP; ++w2) { carry0 += sm.tot[w2 * 52 + 2 * l]; carry1 += sm.tot[w2 * 52 + 2 * l + 1]; }
      zt0 = sm.qnz[m.nz_zT + 2 * l]; zb0 = sm.qnz[m.nz_zb + 2 * l];
      if (2 * l + 1 < S) { zt1 = sm.qnz[m.nz_zT + 2 * l + 1]; zb1 = sm.qnz[m.nz_zb + 2 * l + 1]; }
    }
    const uint32_t colofs = (uint32_t)(l >> 2) * A_LBO + (uint32_t)(l & 3) * 4;
#pragma unroll
    for (int d = 0; d < 16; ++d) {
      const int t = 16 * w + d;
      float v0 = 0.f, v1 = 0.f;
      if (act) {
        if (t < T) { v0 = m.a_T * zt0 + m.a_w * (c[2 * d] + carry0); v1 = m.a_T * zt1 + m.a_w * (c[2 * d + 1] + carry1); }
        else if (t == PB_ROW) { v0 = m.a_b * zb0; v1 = m.a_b * zb1; }
        if (2 * l + 1 >= S) v1 = 0.f;
      }
      __half h0, l0, h1, l1;
      ptx::split_f16(v0, h0, l0);
      ptx::split_f16(v1, h1, l1);
      const uint32_t off = colofs + (uint32_t)(t >> 3) * A_SBO + (uint32_t)(t & 7) * 16;
      *reinterpret_cast<__half2*>(sm.a + off) = __halves2half2(h0, h1);
      *reinterpret_cast<__half2*>(sm.a + A_PLANE + off) = __halves2half2(l0, l1);
    }
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();  // S2
  // ---------------- P3: mu_b^T = W^T X^T on the tensor core (2 M-tiles x 3 split products x 4 K-steps)
  if (tid == 0) {
    ptx::tc_fence_after();
    const uint32_t idesc = ptx::make_idesc_f16(128, 64, 0, 0);
    const uint32_t a0 = ptx::smem_u32(sm.a), b0 = ptx::smem_u32(sm.b), tb = ctl->tmem_base;
    for (int tile = 0; tile < 2; ++tile) {
#pragma unroll
      for (int prod = 0; prod < 3; ++prod) {
        const uint32_t ap = a0 + (prod == 2 ? A_PLANE : 0) + tile * 16 * A_SBO;
        const uint32_t bp = b0 + (prod == 1 ? B_PLANE : 0);
        const uint32_t dcol = tb + tile * 64 + (prod == 0 ? TM_D1 : TM_D2);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint64_t ad = ptx::make_smem_desc(ap + ks * 2 * A_LBO, A_LBO, A_SBO);
          uint64_t bd = ptx::make_smem_desc(bp + ks * 2 * B_LBO, B_LBO, B_SBO);
          ptx::mma_f16_ss(dcol, ad, bd, idesc, (prod == 2) ? 1u : (ks > 0));
        }
      }
      ptx::mma_commit(&ctl->bar_mma[tile]);
    }
  }
  // overlapped with the MMA: AR(1) partisan non-response bias, poll_model_2020.stan:91-93 (warp 1)
  if (m.full && w == 1) {
    const float u_rho = sm.qnz[m.nz_urho], u_mu = sm.qnz[m.nz_umu];
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
      if (t < T) { aj = rho; uj = (t == 0) ? m.sig_e * sm.qnz[m.nz_ze] : cst + sig_rho * sm.qnz[m.nz_ze + t]; }
      u[j] = uj;
      B = aj * B + uj; A = aj * A;
    }
    affine_scan(A, B, l);
    float ein = __shfl_up_sync(0xffffffffu, B, 1);
    if (l == 0) ein = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int t = 8 * l + j;
      if (t < T) { ein = rho * ein + u[j]; sm.e[t] = ein; }
    }
    if (l == 0) {
      ctl->rho = rho; ctl->mu_e = mu_e; ctl->sig_rho = sig_rho; ctl->rho_term = m.sig_e * rho / fmaxf(s2, 1e-20f);
      // -(normal(0.7,0.1) prior + log-Jacobian of the (0,1) transform), fp64
      const double r = 1.0 / (1.0 + exp(-(double)u_rho));
      ctl->u_extra = 0.5 * ((r - 0.7) / 0.1) * ((r - 0.7) / 0.1) - log(r) - log1p(-r);
    }
  }
  if (!m.full && tid == 32) ctl->u_extra = 0.0;
  // ---------------- P4: epilogue 1: TMEM -> mu_b (+prior) in the fp32 scratch
  tmem_epilogue(m, sm, tc, (1.0f / 256.0f), true);
  __syncthreads();  // S3
  // ---------------- optional outputs (transformed parameters / generated quantities of this point)
  if (em.draw != nullptr) {
    float* o = em.draw;
    for (int i = tid; i < S * T; i += NT) { int t = i / S, s = i - t * S; o[i] = sm.scr[t * SCR_PITCH + s]; }
    o += S * T;
    for (int i = tid; i < m.P; i += NT) o[i] = m.sig_c * sm.qnz[m.nz_c + i];
    o += m.P;
    for (int i = tid; i < m.M; i += NT) o[i] = m.full ? m.sig_m * sm.qnz[m.nz_m + i] : 0.f;
    o += m.M;
    for (int i = tid; i < m.Pop; i += NT) o[i] = m.full ? m.sig_pop * sm.qnz[m.nz_pop + i] : 0.f;
    o += m.Pop;
    for (int i = tid; i < T; i += NT) o[i] = m.full ? sm.e[i] : 0.f;
    o += T;
    for (int i = tid; i < S; i += NT) o[i] = sm.scr[PB_ROW * SCR_PITCH + i];
    o += S;
#pragma unroll 4
    for (int e = 0; e < EPT; ++e) {
      int si = m.map_i2s[e * NT + tid];
      if (si >= 0) o[si] = *q_elem(sm, tc, e);
    }
  }
  if (em.monitor != nullptr) {
    for (int i = tid; i < S; i += NT) em.monitor[i] = sm.scr[(T - 1) * SCR_PITCH + i];
    if (tid == 0) em.monitor[S] = sm.scr[(T - 1) * SCR_PITCH + NAT_COL];
  }
  if (em.mu != nullptr)
    for (int i = tid; i < S * T; i += NT) { int t = i / S, s = i - t * S; em.mu[i] = sm.scr[t * SCR_PITCH + s]; }

  // ---------------- P5: polls: linear predictor (stan:95-112), centred binomial_logit (stan:130-131), residuals
  float fsum = 0.f, gm[MAX_MODE] = {0.f, 0.f, 0.f, 0.f}, gp[MAX_MODE] = {0.f, 0.f, 0.f, 0.f}, rnat = 0.f;
  for (int k = tid; k < m.N; k += NT) {
    const uint32_t ix = sm.pk_idx[k];
    const int s = ix & 63, d = (ix >> 6) & 255, p = (ix >> 14) & 1023, mo = (ix >> 24) & 7, po = (ix >> 27) & 7;
    const float un = (float)((ix >> 30) & 1);
    const bool nat = (s == NAT_COL);
    const float sigx = nat ? m.sig_n : m.sig_s;
    float eta = sm.scr[d * SCR_PITCH + s] + sm.scr[PB_ROW * SCR_PITCH + s] + m.sig_c * sm.qnz[m.nz_c + p] +
                sigx * sm.qnz[m.nz_x + k];
    if (m.full) eta += m.sig_m * sm.qnz[m.nz_m + mo] + m.sig_pop * sm.qnz[m.nz_pop + po] + un * sm.e[d];
    const float n = sm.pk_n[k], eh = sm.pk_eta[k], ph = sm.pk_p[k], rh = sm.pk_rho[k];
    const float dl = eta - eh;
    float f, r;
    if (fabsf(dl) < 12.0f) {
      // ll(eta) - ll(eta_hat) = n [ (y/n) dl - log1p(p_hat expm1(dl)) ],  y/n = p_hat + rho_hat
      const float em1 = expm1f(dl);
      const float uu = ph * em1;
      f = n * (rh * dl + (ph * dl - log1pf(uu)));
      r = n * (rh - ph * (1.0f - ph) * em1 / (1.0f + uu));
    } else {  // far tail: direct, stable softplus difference (accuracy irrelevant out here)
      const float sp = fmaxf(eta, 0.f) + log1pf(__expf(-fabsf(eta)));
      const float sph = fmaxf(eh, 0.f) + log1pf(__expf(-fabsf(eh)));
      const float sg = 1.0f / (1.0f + __expf(-eta));
      f = n * ((ph + rh) * dl - (sp - sph));
      r = n * ((ph + rh) - sg);
    }
    fsum += f;
    sm.rr[k] = r;
    sm.gnz[m.nz_x + k] = sigx * r;
    if (nat) rnat += r;
    if (m.full) {
#pragma unroll
      for (int j = 0; j < MAX_MODE; ++j) { gm[j] += (mo == j) ? r : 0.f; gp[j] += (po == j) ? r : 0.f; }
    }
  }
  {
    double v[10];
    v[0] = 0.5 * (double)qsq - (double)fsum;
    v[1] = (double)rnat;
#pragma unroll
    for (int j = 0; j < MAX_MODE; ++j) { v[2 + j] = (double)gm[j]; v[6 + j] = (double)gp[j]; }
#pragma unroll
    for (int i = 0; i < 10; ++i) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
    }
    if (l == 0) {
#pragma unroll
      for (int i = 0; i < 10; ++i) sm.red[w * 12 + i] = v[i];
    }
  }
  __syncthreads();  // S4
  // ---------------- P6: level-1 segment sums of residuals (cells, days, pollsters, states)
  for (int i = tid; i < m.n_t1; i += NT) {
    const uint32_t td = __ldg(m.t1 + i);
    const int start = td & 0xffff, cnt = (td >> 16) & 0xff, type = td >> 24;
    float acc = 0.f;
    if (type == 0) {
      for (int j = 0; j < cnt; ++j) acc += sm.rr[start + j];
    } else if (type == 1) {
      for (int j = 0; j < cnt; ++j) acc += ((sm.pk_idx[start + j] >> 30) & 1) ? sm.rr[start + j] : 0.f;
    } else {
      for (int j = 0; j < cnt; ++j) acc += sm.rr[__ldg(m.ids + start + j)];
    }
    sm.psum[i] = acc;
  }
  if (w == 0 && l < 10) {  // finalize the block reduction
    double s = 0;
    for (int w2 = 0; w2 < NWARP; ++w2) s += sm.red[w2 * 12 + l];
    if (l == 0) ctl->U = s;  // u_extra added below (written by warp 1 before S3)
    else if (l == 1) ctl->rn_total = (float)s;
    else if (m.full) {
      if (l < 2 + MAX_MODE) { if (l - 2 < m.M) sm.gnz[m.nz_m + l - 2] = m.sig_m * (float)s; }
      else if (l - 6 < m.Pop) sm.gnz[m.nz_pop + l - 6] = m.sig_pop * (float)s;
    }
  }
  // zero the operand planes (scratch reads finished at S4)
  {
    uint4* a4 = reinterpret_cast<uint4*>(sm.a);
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < (int)(A_REGION / 16); i += NT) a4[i] = z;
  }
  __syncthreads();  // S5
  // ---------------- P7: level-2 finals -> G operand cells / pollster gradients / g_e
  for (int i = tid; i < m.n_t2; i += NT) {
    const uint2 td = __ldg(m.t2 + i);
    const int ps = td.x & 0xffff, pc = (td.x >> 16) & 0xff, kind = td.x >> 24;
    float acc = 0.f;
    for (int j = 0; j < pc; ++j) acc += sm.psum[ps + j];
    if (kind == 0) {
      const int t = td.y >> 6, s = td.y & 63;
      __half hi, lo;
      ptx::split_f16(acc * m.scale_G, hi, lo);
      const uint32_t off = (uint32_t)(s >> 3) * A_LBO + (uint32_t)(t >> 3) * A_SBO + (uint32_t)(t & 7) * 16 + (uint32_t)(s & 7) * 2;
      *reinterpret_cast<__half*>(sm.a + off) = hi;
      *reinterpret_cast<__half*>(sm.a + A_PLANE + off) = lo;
    } else if (kind == 1) {
      sm.gnz[td.y] = m.sig_c * acc;
    } else {
      sm.ebar[td.y] = acc;
    }
  }
  if (tid == 0) {  // polling-bias row, national K-slot: sum of all national residuals
    __half hi, lo;
    ptx::split_f16(ctl->rn_total * m.scale_G, hi, lo);
    const int t = PB_ROW, s = NAT_COL;
    const uint32_t off = (uint32_t)(s >> 3) * A_LBO + (uint32_t)(t >> 3) * A_SBO + (uint32_t)(t & 7) * 16 + (uint32_t)(s & 7) * 2;
    *reinterpret_cast<__half*>(sm.a + off) = hi;
    *reinterpret_cast<__half*>(sm.a + A_PLANE + off) = lo;
    ctl->U += ctl->u_extra;
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();  // S6
  // ---------------- P8: H^T = G^T X  (B consumed MN-major: the same X planes, transposed view)
  if (tid == 0) {
    ptx::tc_fence_after();
    const uint32_t idesc = ptx::make_idesc_f16(128, 64, 0, 1);
    const uint32_t a0 = ptx::smem_u32(sm.a), b0 = ptx::smem_u32(sm.b), tb = ctl->tmem_base;
    for (int tile = 0; tile < 2; ++tile) {
#pragma unroll
      for (int prod = 0; prod < 3; ++prod) {
        const uint32_t ap = a0 + (prod == 2 ? A_PLANE : 0) + tile * 16 * A_SBO;
        const uint32_t bp = b0 + (prod == 1 ? B_PLANE : 0);
        const uint32_t dcol = tb + tile * 64 + (prod == 0 ? TM_D1 : TM_D2);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint64_t ad = ptx::make_smem_desc(ap + ks * 2 * A_LBO, A_LBO, A_SBO);
          uint64_t bd = ptx::make_smem_desc(bp + ks * 2 * B_SBO, /*LBO: k groups*/ B_SBO, /*SBO: n groups*/ B_LBO);
          ptx::mma_f16_ss(dcol, ad, bd, idesc, (prod == 2) ? 1u : (ks > 0));
        }
      }
      ptx::mma_commit(&ctl->bar_mma[tile]);
    }
  }
  // overlapped: adjoint of the AR(1) recurrence (warp 1) -> gradients of raw_e_bias, mu_e_bias, rho_e_bias
  if (m.full && w == 1) {
    const float rho = ctl->rho, mu_e = ctl->mu_e, sig_rho = ctl->sig_rho, rterm = ctl->rho_term;
    // reversed order: lane l handles t = T-1-8l-j
    float A = 1.f, B = 0.f, u[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int t = T - 1 - (8 * l + j);
      float uj = 0.f, aj = 1.f;
      if (t >= 0) { aj = rho; uj = sm.ebar[t]; }
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
          sm.gnz[m.nz_ze + t] = sig_rho * ein;
          s_mu += ein;
          s_rho += ein * ((sm.e[t - 1] - mu_e) - sm.qnz[m.nz_ze + t] * rterm);
        } else {
          sm.gnz[m.nz_ze] = m.sig_e * ein;
        }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { s_mu += __shfl_xor_sync(0xffffffffu, s_mu, off); s_rho += __shfl_xor_sync(0xffffffffu, s_rho, off); }
    if (l == 0) {
      sm.gnz[m.nz_umu] = 0.02f * (1.0f - rho) * s_mu;
      const float d_rho = s_rho - (rho - 0.7f) * 100.0f;
      // stored as (d lp/d u_rho + u_rho) so that the generic "theta - gnz" form yields -d lp/d u_rho
      sm.gnz[m.nz_urho] = rho * (1.0f - rho) * d_rho + (1.0f - 2.0f * rho) + sm.qnz[m.nz_urho];
    }
  }
  tc.ph ^= 1;  // bar_mma phases of GEMM 1 consumed
  // ---------------- P9: epilogue 2: H -> scratch
  tmem_epilogue(m, sm, tc, m.inv_scale_G * (1.0f / 256.0f), false);
  tc.ph ^= 1;
  __syncthreads();  // S7
  // ---------------- P10/P11: forward cumulative sum of H over days -> gradient of the walk block
  {
    float run0 = 0.f, run1 = 0.f;
    if (tc.zlane) {
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        const int t = 16 * w + d;
        if (t < T) { run0 += sm.scr[t * SCR_PITCH + 2 * l]; run1 += sm.scr[t * SCR_PITCH + 2 * l + 1]; }
        c[2 * d] = run0; c[2 * d + 1] = run1;
      }
      sm.tot[w * 52 + 2 * l] = run0;
      sm.tot[w * 52 + 2 * l + 1] = run1;
    }
  }
  __syncthreads();  // S8
  {
    // every lane computes its 16-element half in (lane-divergent) arithmetic, then the whole warp
    // converges for the .sync.aligned TMEM store
    float carry0 = 0.f, carry1 = 0.f;
    if (tc.zlane)
      for (int w2 = 0; w2 < w; ++w2) { carry0 += sm.tot[w2 * 52 + 2 * l]; carry1 += sm.tot[w2 * 52 + 2 * l + 1]; }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float g[16];
      if (tc.zlane) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int e = h * 16 + j, d = e >> 1, b = e & 1, t = 16 * w + d;
          float v = 0.f;
          if (t < T && 2 * l + b < S) {
            const float z = sm.qz[t * QZ_PITCH + 2 * l + b];
            v = (t <= T - 2) ? z - m.a_w * (c[e] + (b ? carry1 : carry0)) : z;
          }
          g[j] = v;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int k = tc.nz0 + h * 16 + j;
          float v = 0.f;
          if (k < m.NZ) {
            const float th = sm.qnz[k];
            float gs;
            if (k >= m.nz_zT && k < m.nz_zT + S) {
              float s = 0.f;
              for (int w2 = 0; w2 < NWARP; ++w2) s += sm.tot[w2 * 52 + (k - m.nz_zT)];
              gs = m.a_T * s;
            } else if (k >= m.nz_zb && k < m.nz_zb + S) {
              gs = m.a_b * sm.scr[PB_ROW * SCR_PITCH + (k - m.nz_zb)];
            } else {
              gs = sm.gnz[k];
            }
            v = th - gs;
          }
          g[j] = v;
        }
      }
      __syncwarp();
      tm_st16(tc, TM_G + 16 * h, g);
    }
    ptx::tmem_wait_st();
  }
  // note: callers synchronise before reading ctl->U
}

// ================================================================================================
// small block-wide helpers
// ================================================================================================
// sum of squares reduction -> every thread gets the total (fixed summation order)
__device__ __forceinline__ float block_sum_f(const Smem& sm, const TC& tc, float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();  // protect kred against the previous use
  if (tc.l == 0) sm.ctl->kred[tc.w] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NWARP; ++i) s += sm.ctl->kred[i];
  return s;
}

// momentum refresh: P ~ N(0, I) in whitened coordinates -> TMEM column block `col`; returns |P|^2 partial
__device__ __forceinline__ float draw_momentum(const RunArgs& a, const TC& tc, uint32_t chain_gid, uint32_t iter, uint32_t stream,
                                               uint32_t sub, uint32_t col) {
  float ss = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int si = a.m.map_i2s[(h * 16 + j) * NT + tc.tid];
      v[j] = (si >= 0) ? rng_normal(a.seed, chain_gid, (uint32_t)si, iter, stream, sub) : 0.f;
      ss += v[j] * v[j];
    }
    tm_st16(tc, col + 16 * h, v);
  }
  ptx::tmem_wait_st();
  return ss;
}

// Phase A of a leaf: P = p_half - h*s*g (full-step momentum) -> TM_TMP; returns |P|^2 partial.
__device__ __forceinline__ float full_step_momentum(const TC& tc, float hs) {
  float ss = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float p[16], s[16], g[16];
    tm_ld16(tc, TM_P + 16 * h, p);
    tm_ld16(tc, TM_S + 16 * h, s);
    tm_ld16(tc, TM_G + 16 * h, g);
#pragma unroll
    for (int j = 0; j < 16; ++j) { p[j] = p[j] - hs * s[j] * g[j]; ss += p[j] * p[j]; }
    tm_st16(tc, TM_TMP + 16 * h, p);
  }
  ptx::tmem_wait_st();
  return ss;
}
// Phase C: p_half' = 2P - p_half ; q' = q + eps_signed * s * p_half'
__device__ __forceinline__ void advance(const ModelDev& m, const Smem& sm, const TC& tc, float eps_signed) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float p[16], s[16], P[16];
    tm_ld16(tc, TM_P + 16 * h, p);
    tm_ld16(tc, TM_S + 16 * h, s);
    tm_ld16(tc, TM_TMP + 16 * h, P);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      p[j] = 2.0f * P[j] - p[j];
      const int e = h * 16 + j;
      if (elem_valid(m, tc, e)) { float* q = q_elem(sm, tc, e); *q = *q + eps_signed * s[j] * p[j]; }
    }
    tm_st16(tc, TM_P + 16 * h, p);
  }
  ptx::tmem_wait_st();
}

// One U-turn merge (Stan's three criteria) between the completed left subtree L = {b,e,r} and the
// implicit right subtree R = {b: rb (or P if null), r: P + S, e: P}; S (TM_G) += L.r afterwards.
// `first` : S is implicitly zero.   `single`: L.b = L.e = L.r (Left_0).
__device__ __forceinline__ bool merge_check(const Smem& sm, const TC& tc, const float* Lb, const float* Le, const float* Lr,
                                            const float* Rb, bool first, bool single) {
  float c1a = 0.f, c1b = 0.f, c2a = 0.f, c2b = 0.f, c3a = 0.f, c3b = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float P[16], S[16];
    tm_ld16(tc, TM_TMP + 16 * h, P);
    if (first) {
#pragma unroll
      for (int j = 0; j < 16; ++j) S[j] = 0.f;
    } else {
      tm_ld16(tc, TM_G + 16 * h, S);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int gi = (h * 16 + j) * NT + tc.tid;
      const float lr = Lr[gi];
      const float lb = single ? lr : Lb[gi];
      const float le = single ? lr : Le[gi];
      const float rb = (Rb == nullptr) ? P[j] : Rb[gi];
      const float x = lr + S[j] + P[j];
      c1a += lb * x; c1b += P[j] * x;
      const float y = lr + rb;
      c2a += lb * y; c2b += rb * y;
      const float z = S[j] + P[j] + le;
      c3a += le * z; c3b += P[j] * z;
      S[j] += lr;
    }
    tm_st16(tc, TM_G + 16 * h, S);
  }
  ptx::tmem_wait_st();
  float v[6] = {c1a, c1b, c2a, c2b, c3a, c3b};
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
  }
  __syncthreads();
  if (tc.l == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) sm.ctl->mred[tc.w][i] = v[i];
  }
  __syncthreads();
  bool pred = true;
  if (tc.tid < 6) {
    float s = 0.f;
    for (int w2 = 0; w2 < NWARP; ++w2) s += sm.ctl->mred[w2][tc.tid];
    pred = s > 0.f;
  }
  return __syncthreads_and(pred) != 0;
}

__device__ __forceinline__ double logaddexp_d(double a, double b) {
  if (a == -CUDART_INF) return b;
  if (b == -CUDART_INF) return a;
  return a > b ? a + log1p(exp(b - a)) : b + log1p(exp(a - b));
}

struct TransStats { float lp, accept, eps, depth, nleap, divergent, energy; };

// ================================================================================================
// one NUTS transition (Stan base_nuts::transition, iterative tree).  On entry q(smem) = current point,
// TM_S = sqrt(inverse metric).  On exit q(smem) = new sample.  `em`: outputs for the ENTRY point
// (the previous iteration's draw), emitted during the initial gradient evaluation.
// ================================================================================================
__device__ __noinline__ void transition(const RunArgs& a, const Smem& sm, TC& tc, float* ws, uint32_t chain_gid, uint32_t iter,
                                        float eps, const Emit em, TransStats& st) {
  const ModelDev& m = a.m;
  Ctl* ctl = sm.ctl;
  // fresh whitened momentum into TM_P, gradient at the current point
  float ksq = draw_momentum(a, tc, chain_gid, iter, 1, 0, TM_P);
  eval_point(m, sm, tc, em);
  ksq = block_sum_f(sm, tc, ksq);  // (contains the barrier that publishes ctl->U)
  const double U0 = ctl->U;
  const double H0 = U0 + 0.5 * (double)ksq;
  // the two trajectory ends, stored mid-leapfrog: (q +- eps s p_half, p_half), and the tree summary
  {
    const float hs = 0.5f * eps;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float P[16], s[16], g[16];
      tm_ld16(tc, TM_P + 16 * h, P);
      tm_ld16(tc, TM_S + 16 * h, s);
      tm_ld16(tc, TM_G + 16 * h, g);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int e = h * 16 + j, gi = e * NT + tc.tid;
        const bool ok = elem_valid(m, tc, e);
        const float q0 = ok ? *q_elem(sm, tc, e) : 0.f;
        const float pf = P[j] - hs * s[j] * g[j], pb = P[j] + hs * s[j] * g[j];
        slot_ptr(ws, SLOT_ENDF_P)[gi] = pf;
        slot_ptr(ws, SLOT_ENDB_P)[gi] = pb;
        slot_ptr(ws, SLOT_ENDF_Q)[gi] = q0 + eps * s[j] * pf;
        slot_ptr(ws, SLOT_ENDB_Q)[gi] = q0 - eps * s[j] * pb;
        slot_ptr(ws, SLOT_TOP_BB)[gi] = P[j];
        slot_ptr(ws, SLOT_TOP_FF)[gi] = P[j];
        slot_ptr(ws, SLOT_TOP_RHO)[gi] = P[j];
        slot_ptr(ws, SLOT_CAND_A)[gi] = q0;
      }
    }
  }
  int samp = SLOT_CAND_A, prop = SLOT_CAND_B;
  double U_samp = U0, H_samp = H0, U_prop = U0, H_prop = H0;
  double lsw = 0.0, sum_metro = 0.0;
  int n_leap = 0, depth = 0, loaded = 0;
  bool divergent = false;
  __syncthreads();  // global slot writes by this block are ordered for later reads by the same threads anyway

  while (depth < a.max_depth) {
    uint32_t rw[4];
    rng_words(a.seed, chain_gid, (uint32_t)depth, iter, 2, 0, rw);
    const int dir = (rw[0] >> 31) ? 1 : -1;  // u > 0.5
    if (loaded != dir) {
      if (loaded != 0) {
        q_to_global(m, sm, tc, slot_ptr(ws, loaded > 0 ? SLOT_ENDF_Q : SLOT_ENDB_Q));
        tm_to_global(tc, TM_P, slot_ptr(ws, loaded > 0 ? SLOT_ENDF_P : SLOT_ENDB_P));
      }
      __syncthreads();
      global_to_q(m, sm, tc, slot_ptr(ws, dir > 0 ? SLOT_ENDF_Q : SLOT_ENDB_Q));
      global_to_tm(tc, TM_P, slot_ptr(ws, dir > 0 ? SLOT_ENDF_P : SLOT_ENDB_P));
      loaded = dir;
      __syncthreads();
    }
    const float eps_s = dir > 0 ? eps : -eps;
    const float hs = 0.5f * eps_s;
    double lsw_sub = -CUDART_INF;
    bool ok = true, persist = true;
    const int nleaf = 1 << depth;
    for (int n = 0; n < nleaf; ++n) {
      Emit none{nullptr, nullptr, nullptr};
      eval_point(m, sm, tc, none);
      float kk = full_step_momentum(tc, hs);
      kk = block_sum_f(sm, tc, kk);
      double h = ctl->U + 0.5 * (double)kk;
      if (!(h == h)) h = CUDART_INF;
      ++n_leap;
      if (h - H0 > 1000.0) divergent = true;
      lsw_sub = logaddexp_d(lsw_sub, H0 - h);
      sum_metro += (H0 - h > 0) ? 1.0 : exp(H0 - h);
      if (divergent) { ok = false; break; }
      // multinomial selection inside the new subtree (reservoir form of Stan's pairwise rule)
      {
        uint32_t sw[4];
        rng_words(a.seed, chain_gid, (uint32_t)n_leap, iter, 3, 0, sw);
        if (n == 0 || (double)u01(sw[0]) < exp((H0 - h) - lsw_sub)) {
          q_to_global(m, sm, tc, slot_ptr(ws, prop));
          U_prop = ctl->U; H_prop = h;
        }
      }
      // U-turn checks for every subtree this leaf completes
      int t = 0;
      while ((n >> t) & 1) ++t;
      for (int k = 0; k < t && ok; ++k) {
        if (k == 0) {
          const float* L0p = slot_ptr(ws, SLOT_LEFT0);
          ok = merge_check(sm, tc, L0p, L0p, L0p, nullptr, true, true);
        } else {
          const float* Lk = slot_ptr(ws, SLOT_LEFT + 3 * (k - 1));
          const float* Rb = (k == 1) ? slot_ptr(ws, SLOT_LEFT0) : slot_ptr(ws, SLOT_LEFT + 3 * (k - 2));
          ok = merge_check(sm, tc, Lk, Lk + VEC, Lk + 2 * VEC, Rb, false, false);
        }
      }
      if (!ok) break;
      if (n < nleaf - 1) {
        // this subtree becomes the stored left half at level t: {b, e, r}
        if (t == 0) {
          tm_to_global(tc, TM_TMP, slot_ptr(ws, SLOT_LEFT0));
        } else {
          float* Lt = slot_ptr(ws, SLOT_LEFT + 3 * (t - 1));
          const float* Bsrc = (t == 1) ? slot_ptr(ws, SLOT_LEFT0) : slot_ptr(ws, SLOT_LEFT + 3 * (t - 2));
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            float P[16], S[16];
            tm_ld16(tc, TM_TMP + 16 * hh, P);
            tm_ld16(tc, TM_G + 16 * hh, S);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int gi = (hh * 16 + j) * NT + tc.tid;
              Lt[gi] = Bsrc[gi];
              Lt[VEC + gi] = P[j];
              Lt[2 * VEC + gi] = P[j] + S[j];
            }
          }
        }
      } else {
        // last leaf: merge the finished subtree with the existing trajectory (top level of
        // base_nuts::transition) and fold it into the tree summary
        const float* F = slot_ptr(ws, dir > 0 ? SLOT_TOP_BB : SLOT_TOP_FF);
        const float* A = slot_ptr(ws, dir > 0 ? SLOT_TOP_FF : SLOT_TOP_BB);
        const float* Rb = (depth == 0) ? nullptr : (depth == 1 ? slot_ptr(ws, SLOT_LEFT0) : slot_ptr(ws, SLOT_LEFT + 3 * (depth - 2)));
        persist = merge_check(sm, tc, F, A, slot_ptr(ws, SLOT_TOP_RHO), Rb, depth == 0, false);
        float* rho = slot_ptr(ws, SLOT_TOP_RHO);
        float* endv = slot_ptr(ws, dir > 0 ? SLOT_TOP_FF : SLOT_TOP_BB);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float P[16], S[16];
          tm_ld16(tc, TM_TMP + 16 * hh, P);
          tm_ld16(tc, TM_G + 16 * hh, S);  // = old rho_top + sum of lower lefts
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int gi = (hh * 16 + j) * NT + tc.tid;
            rho[gi] = S[j] + P[j];
            endv[gi] = P[j];
          }
        }
      }
      advance(m, sm, tc, eps_s);
      ptx::tc_fence_before();
      __syncthreads();
    }
    if (!ok) break;
    ++depth;
    // biased progressive sampling at the top level
    if (lsw_sub > lsw || (double)u01(rw[1]) < exp(lsw_sub - lsw)) {
      const int tmp = samp; samp = prop; prop = tmp;
      U_samp = U_prop; H_samp = H_prop;
    }
    lsw = logaddexp_d(lsw, lsw_sub);
    if (!persist) break;
  }
  __syncthreads();
  global_to_q(m, sm, tc, slot_ptr(ws, samp));
  ptx::tc_fence_before();
  __syncthreads();
  st.lp = (float)(-U_samp);  // centred; the host adds lp_const in fp64 (fp32 cannot hold -1.2e6 to 1e-2)
  st.accept = (float)(sum_metro / (double)(n_leap > 0 ? n_leap : 1));
  st.eps = eps; st.depth = (float)depth; st.nleap = (float)n_leap; st.divergent = divergent ? 1.f : 0.f;
  st.energy = (float)H_samp;
  if (tc.tid == 0) { ctl->cs.n_leapfrog += n_leap; ctl->cs.U = (float)U_samp; }
}

// Stan base_hmc::init_stepsize: double / halve eps until the one-step acceptance crosses 0.8
__device__ __noinline__ float find_stepsize(const RunArgs& a, const Smem& sm, TC& tc, float* ws, uint32_t chain_gid, uint32_t iter_tag,
                                            float eps) {
  const ModelDev& m = a.m;
  Ctl* ctl = sm.ctl;
  if (!(eps > 0.f) || eps > 1e7f) return eps;
  q_to_global(m, sm, tc, slot_ptr(ws, SLOT_TMPQ));
  int direction = 0;
  const Emit none{nullptr, nullptr, nullptr};
  for (uint32_t attempt = 0; attempt < 200; ++attempt) {
    __syncthreads();
    global_to_q(m, sm, tc, slot_ptr(ws, SLOT_TMPQ));
    __syncthreads();
    float k0 = draw_momentum(a, tc, chain_gid, iter_tag, 5, attempt, TM_P);
    eval_point(m, sm, tc, none);
    k0 = block_sum_f(sm, tc, k0);
    const double H0 = ctl->U + 0.5 * (double)k0;
    // first half step + position step
    full_step_momentum(tc, 0.5f * eps);  // TM_TMP = p - eps/2 s g  (= p_half)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float p[16], s[16];
      tm_ld16(tc, TM_TMP + 16 * h, p);
      tm_ld16(tc, TM_S + 16 * h, s);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int e = h * 16 + j;
        if (elem_valid(m, tc, e)) { float* q = q_elem(sm, tc, e); *q = *q + eps * s[j] * p[j]; }
      }
      tm_st16(tc, TM_P + 16 * h, p);
    }
    ptx::tmem_wait_st();
    ptx::tc_fence_before();
    __syncthreads();
    eval_point(m, sm, tc, none);
    float k1 = full_step_momentum(tc, 0.5f * eps);
    k1 = block_sum_f(sm, tc, k1);
    double h = ctl->U + 0.5 * (double)k1;
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
  global_to_q(m, sm, tc, slot_ptr(ws, SLOT_TMPQ));
  ptx::tc_fence_before();
  __syncthreads();
  return eps;
}

// ================================================================================================
// kernels
// ================================================================================================
__device__ __forceinline__ void cta_setup(const ModelDev& m, const Smem& sm, TC& tc, unsigned char* base) {
  tc.tid = threadIdx.x; tc.w = tc.tid >> 5; tc.l = tc.tid & 31;
  tc.zlane = tc.l < ZLANES; tc.nzlane = !tc.zlane;
  tc.nz0 = (tc.w * NZ_LANES + (tc.l - ZLANES)) * EPT;
  tc.ph = 0;
  Ctl* ctl = sm.ctl;
  if (tc.w == 0) { ptx::tmem_alloc(&ctl->tmem_base, 512); ptx::tmem_relinquish(); }
  if (tc.tid == 0) {
    ptx::mbar_init(&ctl->bar_mma[0], 1);
    ptx::mbar_init(&ctl->bar_mma[1], 1);
    ptx::mbar_init(&ctl->bar_load, 1);
    ptx::fence_mbar_init();
  }
  // zero all of shared memory that holds padded vectors
  for (int i = tc.tid; i < (int)((SM_CTL - SM_QZ) / 4); i += NT) reinterpret_cast<float*>(base + SM_QZ)[i] = 0.f;
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  // constants: TMA bulk copies (global -> shared), one mbarrier
  if (tc.tid == 0) {
    const uint32_t bytes = 2 * B_PLANE + 5 * NPOLL_CAP * 4 + 64 * 4;
    ptx::mbar_expect_tx(&ctl->bar_load, bytes);
    ptx::bulk_g2s(sm.b, m.btiles, 2 * B_PLANE, &ctl->bar_load);
    ptx::bulk_g2s(sm.pk_idx, m.pk, 5 * NPOLL_CAP * 4, &ctl->bar_load);
    ptx::bulk_g2s(sm.prior, m.prior, 64 * 4, &ctl->bar_load);
  }
  ptx::mbar_wait(&ctl->bar_load, 0);
  tc.tpriv = ctl->tmem_base + ((uint32_t)((tc.w & 3) * 32) << 16) + (uint32_t)(tc.w >> 2) * 32;
  __syncthreads();
}
__device__ __forceinline__ void cta_teardown(const Smem& sm, const TC& tc) {
  ptx::tc_fence_before();
  __syncthreads();
  if (tc.w == 0) ptx::tmem_dealloc(sm.ctl->tmem_base, 512);
}

// test hook: log density + gradient for n positions (potus_logp_grad)
extern "C" __global__ void __launch_bounds__(NT, 1) potus_eval_kernel(const __grid_constant__ EvalArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  Smem sm = carve(base);
  TC tc;
  cta_setup(a.m, sm, tc, base);
  for (int i = blockIdx.x; i < a.n; i += gridDim.x) {
    global_to_q(a.m, sm, tc, a.q_in + (size_t)i * VEC);
    __syncthreads();
    Emit em{nullptr, nullptr, a.mu_out ? a.mu_out + (size_t)i * a.m.S * a.m.T : nullptr};
    eval_point(a.m, sm, tc, em);
    __syncthreads();
    tm_to_global(tc, TM_G, a.g_out + (size_t)i * VEC);
    if (tc.tid == 0) a.u_out[i] = sm.ctl->U;
    ptx::tc_fence_before();
    __syncthreads();
  }
  cta_teardown(sm, tc);
}

// Welford update of the per-chain posterior variance estimate (Stan var_adaptation), owner layout
__device__ __forceinline__ void welford_add(const ModelDev& m, const Smem& sm, const TC& tc, float* mean, float* m2, int nsamp) {
  const float inv = 1.0f / (float)nsamp;
#pragma unroll 8
  for (int e = 0; e < EPT; ++e) {
    if (!elem_valid(m, tc, e)) continue;
    const int gi = e * NT + tc.tid;
    const float q = *q_elem(sm, tc, e);
    const float mu = mean[gi], d = q - mu, mu2 = mu + d * inv;
    mean[gi] = mu2;
    m2[gi] += (q - mu2) * d;
  }
}

extern "C" __global__ void __launch_bounds__(NT, 1) potus_nuts_kernel(const __grid_constant__ RunArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
  Smem sm = carve(base);
  TC tc;
  const ModelDev& m = a.m;
  cta_setup(m, sm, tc, base);
  Ctl* ctl = sm.ctl;
  float* ws = a.workspace + (size_t)blockIdx.x * NSLOT * VEC;
  const int n_iter_total = a.iter_warmup + a.iter_sampling;
  const Emit none{nullptr, nullptr, nullptr};

  for (;;) {
    __syncthreads();
    if (tc.tid == 0) ctl->chain = atomicAdd(a.queue, 1);
    __syncthreads();
    const int chain = ctl->chain;
    if (chain >= a.n_chains) break;
    const uint32_t gid = (uint32_t)(a.chain_id_offset + chain);
    float* qg = a.q + (size_t)chain * VEC;
    float* sg = a.sqrt_m + (size_t)chain * VEC;
    float* wmean = a.wf_mean + (size_t)chain * VEC;
    float* wm2 = a.wf_m2 + (size_t)chain * VEC;
    if (tc.tid == 0) ctl->cs = a.cs[chain];
    __syncthreads();

    if (a.do_init) {
      // ---- random inits U(-r, r) on the unconstrained scale; retry while U / gradient are not finite
      bool good = false;
      for (uint32_t attempt = 0; attempt < 100 && !good; ++attempt) {
#pragma unroll 4
        for (int e = 0; e < EPT; ++e) {
          const int si = m.map_i2s[e * NT + tc.tid];
          if (si >= 0) {
            uint32_t rw[4];
            rng_words(a.seed, gid, (uint32_t)si, 0, 0, attempt, rw);
            *q_elem(sm, tc, e) = a.init_radius * (2.0f * u01(rw[0]) - 1.0f);
          }
        }
        {
          float one[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) one[j] = 1.0f;
          tm_st16(tc, TM_S, one); tm_st16(tc, TM_S + 16, one);
          ptx::tmem_wait_st();
        }
        ptx::tc_fence_before();
        __syncthreads();
        eval_point(m, sm, tc, none);
        int bad = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float g[16];
          tm_ld16(tc, TM_G + 16 * h, g);
#pragma unroll
          for (int j = 0; j < 16; ++j) bad |= !isfinite(g[j]);
        }
        __syncthreads();
        if (!isfinite(ctl->U)) bad = 1;
        good = __syncthreads_or(bad) == 0;
      }
      if (tc.tid == 0) {
        ChainState& cs = ctl->cs;
        cs.status = good ? 0 : -1;
        cs.eps = 1.0f; cs.da_counter = 0; cs.da_sbar = 0; cs.da_xbar = 0;
        cs.w_counter = 0; cs.w_size = a.w_base_window; cs.w_next = a.w_init_buffer + a.w_base_window - 1; cs.w_nsamp = 0;
        cs.iter = 0; cs.n_leapfrog = 0;
      }
      // Welford accumulators start at zero
#pragma unroll 8
      for (int e = 0; e < EPT; ++e) { wmean[e * NT + tc.tid] = 0.f; wm2[e * NT + tc.tid] = 0.f; }
      __syncthreads();
      const float e0 = find_stepsize(a, sm, tc, ws, gid, 0xFFFFFFFFu, 1.0f);
      if (tc.tid == 0) { ctl->cs.eps = e0; ctl->cs.da_mu = log(10.0 * (double)e0); }
      __syncthreads();
    } else {
      global_to_q(m, sm, tc, qg);
      global_to_tm(tc, TM_S, sg);
      ptx::tc_fence_before();
      __syncthreads();
    }

    // ---- iterations
    for (int it = a.iter_begin; it < a.iter_end; ++it) {
      const float eps = ctl->cs.eps;
      // outputs of the ENTRY point = draw of iteration it-1 (if it was a sampling iteration)
      Emit em = none;
      const int kprev = it - 1 - a.iter_warmup;
      if (kprev >= 0) {
        em.monitor = a.monitor + ((size_t)chain * a.iter_sampling + kprev) * (m.S + 1);
        if (a.keep_per_chain > 0 && (kprev % a.keep_every) == a.keep_every - 1) {
          const int slot = kprev / a.keep_every;
          if (slot < a.keep_per_chain) em.draw = a.draws + ((size_t)chain * a.keep_per_chain + slot) * a.draw_len;
        }
      }
      TransStats st;
      transition(a, sm, tc, ws, gid, (uint32_t)it, eps, em, st);
      if (tc.tid == 0) {
        float* sp = a.sampler_params + ((size_t)chain * n_iter_total + it) * 8;
        sp[0] = st.lp; sp[1] = st.accept; sp[2] = st.eps; sp[3] = st.depth; sp[4] = st.nleap; sp[5] = st.divergent; sp[6] = st.energy; sp[7] = 0.f;
      }
      if (it < a.iter_warmup) {
        // ---- Stan stepsize_adaptation::learn_stepsize (dual averaging)
        if (tc.tid == 0) {
          ChainState& cs = ctl->cs;
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
        const int wc = ctl->cs.w_counter;
        const bool in_window = wc >= a.w_init_buffer && wc < a.iter_warmup - a.w_term_buffer && wc != a.iter_warmup;
        const bool end_window = wc == ctl->cs.w_next && wc != a.iter_warmup;
        int nsamp = ctl->cs.w_nsamp;
        if (in_window) { ++nsamp; welford_add(m, sm, tc, wmean, wm2, nsamp); }
        __syncthreads();
        if (end_window) {
          const float n = (float)nsamp;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float s[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int e = h * 16 + j, gi = e * NT + tc.tid;
              float v = 1.0f;
              if (elem_valid(m, tc, e)) {
                const float var = wm2[gi] / (n - 1.0f);
                v = sqrtf((n / (n + 5.0f)) * var + 1e-3f * (5.0f / (n + 5.0f)));
              }
              s[j] = v;
              wmean[gi] = 0.f; wm2[gi] = 0.f;
            }
            tm_st16(tc, TM_S + 16 * h, s);
          }
          ptx::tmem_wait_st();
          if (tc.tid == 0) {
            ChainState& cs = ctl->cs;
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
          const float e1 = find_stepsize(a, sm, tc, ws, gid, (uint32_t)it, ctl->cs.eps);
          if (tc.tid == 0) {
            ChainState& cs = ctl->cs;
            cs.eps = e1; cs.da_mu = log(10.0 * (double)e1); cs.da_counter = 0; cs.da_sbar = 0; cs.da_xbar = 0;
          }
        }
        if (tc.tid == 0) {
          ctl->cs.w_nsamp = nsamp;
          ctl->cs.w_counter = wc + 1;
          if (it == a.iter_warmup - 1) ctl->cs.eps = (float)exp(ctl->cs.da_xbar);
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
        if (a.keep_per_chain > 0 && (kprev % a.keep_every) == a.keep_every - 1) {
          const int slot = kprev / a.keep_every;
          if (slot < a.keep_per_chain) em.draw = a.draws + ((size_t)chain * a.keep_per_chain + slot) * a.draw_len;
        }
        eval_point(m, sm, tc, em);
        __syncthreads();
      }
    }
    // ---- persist the chain
    q_to_global(m, sm, tc, qg);
    tm_to_global(tc, TM_S, sg);
    if (tc.tid == 0) { ctl->cs.iter = a.iter_end; a.cs[chain] = ctl->cs; }
    ptx::tc_fence_before();
    __syncthreads();
  }
  cta_teardown(sm, tc);
}

}  // namespace potus
