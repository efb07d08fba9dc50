// Streaming NUTS kernel for poll_model_2020 at sizes the resident kernel cannot hold (sm_100a):
// BASELINE config 5 (S=256, T=365, N=50k) and any S<=256, T<=512.  One CTA = one chain; the chain's vectors
// live in global memory in the stream layout (potus_stream_layout.h) and every leapfrog is ONE forward sweep
// over the days in 128-day tiles:
//   W tile (suffix sums of the walk innovations, via the carried column sums)  -> fp16 hi/lo operand planes
//   -> tcgen05.mma  mu_b^T = W^T X^T   (M=128 days, N<=256 states, K-steps of 16 streamed through a 3-stage
//      ring of bulk-async copies; X = 256 L0 is lower triangular, so K-step kc only multiplies N - 16 kc columns)
//   -> TMEM epilogue: mu_b tile in shared memory (+ prior), national average per day
//   -> polls of these days (centred binomial-logit), residuals, measurement-noise parameters updated in place
//   -> G tile (residual sums per (state, day) cell) -> fp16 hi/lo planes -> tcgen05.mma  H^T = G^T X
//   -> TMEM epilogue (+ rank-1 national term) -> prefix sums over days -> gradient of the walk block ->
//      momentum and position update written straight back (q, p_half read once, q', p_half' written once).
// Everything that is not day-local (pollster sums, AR(1) adjoint, raw_mu_b_T / raw_polling_bias gradients)
// is finished after the last tile from per-sweep accumulators.
//
// What it restates: exactly what potus_kernel.cu does (same reference lines, same Philox streams, same
// iterative NUTS, same adaptation); the oracle is oracle/potus_oracle.c (tree_mode 1).
#include <cuda_runtime.h>
#include <math_constants.h>
#include "ptx_sm100.cuh"
#include "potus_stream_layout.h"

namespace potus {

struct SCtl {
  uint64_t bar_full[ST_NSTAGE], bar_empty[ST_NSTAGE], bar_done;
  uint32_t tmem_base;
  uint32_t prod_it, cons_it, gemm_cnt;
  int chain;
  double U;        // potential of the last evaluated point (centred)
  double u_extra;
  float rho, mu_e, sig_rho, rho_term, nat_pb, pad0;
  const float* czn_tag;   // vector whose walk-block column sums SS_CZN holds (nullptr: none)
  alignas(16) float mred[16][8];
  ChainState cs;
  double U_samp, H_samp, U_prop, H_prop;
  float sum_metro;
  int pad1;
#ifdef POTUS_PROF
  unsigned long long prof[32];
#endif
};
static_assert(sizeof(SCtl) <= 1024, "control block");
// optional phase clocks (development builds, -DPOTUS_PROF): thread 0 accumulates cycles per phase of the sweep
#ifdef POTUS_PROF
#define SPROF_DECL long long sprof_t_ = clock64()
#define SPROF(i) do { if (threadIdx.x == 0) { long long n_ = clock64(); SCTL().prof[i] += (unsigned long long)(n_ - sprof_t_); sprof_t_ = n_; } } while (0)
#else
#define SPROF_DECL
#define SPROF(i)
#endif

__device__ __forceinline__ const ModelS& SMD() { return *SMP(const ModelS, SS_MODEL); }
__device__ __forceinline__ SCtl& SCTL() { return *SMP(SCtl, SS_CTL); }
__device__ __forceinline__ float* sF(uint32_t off) { return SMP(float, off); }

struct SEmit {
  float* draw;     // [draw_len] or null
  float* monitor;  // [S+1] or null
};

// what one sweep reads and writes (all pointers are stream-layout vectors in global memory)
struct SweepIO {
  const float* qin;    // position at which U and its gradient are evaluated
  // leaf mode (the NUTS inner loop)
  float* qout;         // next position q' = q + eps s p_half'
  float* ph;           // half-step whitened momentum, updated in place
  const float* sm;     // sqrt(inverse metric)
  float* Pdst;         // full-step momentum of this leaf
  const float* Lr;     // odd leaf: momentum of the previous leaf (level-0 U-turn partner), else null
  float eps_s;         // signed step
  // plain evaluation
  float* gout;         // gradient of U
  SEmit em;
  // leaf results (block totals, every thread)
  float kk, c1a, c1b;
};

// ------------------------------------------------------------------------------------------------
// B-operand streaming: one K-step (16 k) per ring stage.  Thread 32 produces (bulk async copies),
// thread 0 issues the MMAs; chunk counters persist in the control block across GEMMs.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s_chunk_n(const ModelS& m, int which, int kc) {   // MMA N of K-step kc
  return which == 0 ? (uint32_t)(m.NP - 16 * kc) : (uint32_t)min(m.NP, 16 * (kc + 1));
}
__device__ __forceinline__ void s_produce(int which, int kc_begin, int kc_end) {   // thread 32 only
  const ModelS& m = SMD();
  SCtl& c = SCTL();
  const unsigned char* src = which == 0 ? m.b1 : m.b2;
  const uint32_t* off = which == 0 ? m.b1_off : m.b2_off;
  for (int kc = kc_begin; kc < kc_end; ++kc) {
    const uint32_t it = c.prod_it, st = it % ST_NSTAGE, use = it / ST_NSTAGE;
    if (use > 0) ptx::mbar_wait(&c.bar_empty[st], (use - 1) & 1);
    const uint32_t o0 = __ldg(off + kc), bytes = __ldg(off + kc + 1) - o0;
    ptx::mbar_expect_tx(&c.bar_full[st], bytes);
    ptx::bulk_g2s(smem_raw + SS_B + st * SB_STAGE, src + o0, bytes, &c.bar_full[st]);
    c.prod_it = it + 1;
  }
}
// issue the split-precision product D1 = A_hi B_hi ; D2 = A_hi B_lo + A_lo B_hi  (thread 0 only)
__device__ __forceinline__ void s_issue(int which) {
  const ModelS& m = SMD();
  SCtl& c = SCTL();
  ptx::tc_fence_after();
  const uint32_t a0 = ptx::smem_u32(smem_raw + SS_A), b0 = ptx::smem_u32(smem_raw + SS_B), tb = c.tmem_base;
  for (int kc = 0; kc < m.KS; ++kc) {
    const uint32_t it = c.cons_it, st = it % ST_NSTAGE, use = it / ST_NSTAGE;
    ptx::mbar_wait(&c.bar_full[st], use & 1);
    ptx::tc_fence_after();
    const uint32_t n = s_chunk_n(m, which, kc);
    const uint32_t ncol0 = which == 0 ? 16u * kc : 0u;          // first output column of this K-step
    const uint32_t lbo_b = n * 16;                              // k-group stride inside the packed chunk
    const uint32_t idesc = ptx::make_idesc_f16(128, (int)n, 0, 0);
    const uint32_t bhi = b0 + st * SB_STAGE, blo = bhi + 2 * lbo_b;
    const uint64_t ahi_d = ptx::make_smem_desc(a0 + kc * 2 * SA_LBO, SA_LBO, SA_SBO);
    const uint64_t alo_d = ptx::make_smem_desc(a0 + SA_PLANE + kc * 2 * SA_LBO, SA_LBO, SA_SBO);
    const uint64_t bhi_d = ptx::make_smem_desc(bhi, lbo_b, 128), blo_d = ptx::make_smem_desc(blo, lbo_b, 128);
    // accumulate flag: column c of D has been written before by K-steps kc' < kc iff it lies in their N range
    //   which 0 (N shrinks from the left): all columns [16kc, NP) were written by every earlier step -> accumulate iff kc > 0
    //   which 1 (N grows to the right): columns [16kc, 16kc+16) are new -> two MMAs: old columns accumulate, new ones overwrite
    if (which == 0) {
      ptx::mma_f16_ss(tb + ncol0, ahi_d, bhi_d, idesc, kc > 0);
      ptx::mma_f16_ss(tb + 256 + ncol0, ahi_d, blo_d, idesc, kc > 0);
      ptx::mma_f16_ss(tb + 256 + ncol0, alo_d, bhi_d, idesc, 1u);
    } else {
      const uint32_t nold = 16u * kc;
      if (nold > 0) {
        const uint32_t idold = ptx::make_idesc_f16(128, (int)nold, 0, 0);
        ptx::mma_f16_ss(tb, ahi_d, bhi_d, idold, 1u);
        ptx::mma_f16_ss(tb + 256, ahi_d, blo_d, idold, 1u);
        ptx::mma_f16_ss(tb + 256, alo_d, bhi_d, idold, 1u);
      }
      if (nold < n) {   // the 16 new columns: rows [nold, n) of the chunk
        const uint32_t idnew = ptx::make_idesc_f16(128, (int)(n - nold), 0, 0);
        const uint64_t bhn = ptx::make_smem_desc(bhi + (nold / 8) * 128, lbo_b, 128), bln = ptx::make_smem_desc(blo + (nold / 8) * 128, lbo_b, 128);
        ptx::mma_f16_ss(tb + nold, ahi_d, bhn, idnew, 0u);
        ptx::mma_f16_ss(tb + 256 + nold, ahi_d, bln, idnew, 0u);
        ptx::mma_f16_ss(tb + 256 + nold, alo_d, bhn, idnew, 1u);
      }
    }
    ptx::mma_commit(&c.bar_empty[st]);
    c.cons_it = it + 1;
  }
  ptx::mma_commit(&c.bar_done);
}
// all threads: run GEMM `which` on the operand planes in SS_A (made visible by the caller's fence + barrier)
__device__ __forceinline__ void s_gemm(int which, int prefetched) {
  const int tid = threadIdx.x;
  SCtl& c = SCTL();
  const uint32_t par = c.gemm_cnt & 1;
  if (tid == 32) s_produce(which, prefetched, SMD().KS);
  if (tid == 0) s_issue(which);
  // (everyone but the issuer and the producer has nothing to do until the accumulators are complete: back off instead of
  //  spinning, the spin was 4.6 % of all issued instructions and competes with the two working threads for issue slots)
  while (!ptx::mbar_try_wait(&c.bar_done, par)) {
    if (tid != 0 && tid != 32) __nanosleep(128);
  }
  ptx::tc_fence_after();
}
__device__ __forceinline__ int s_prefetch(int which) {   // called by ALL threads with a uniform result; thread 32 copies
  const int n = min(SMD().KS, ST_NSTAGE);
  if (threadIdx.x == 32) s_produce(which, 0, n);
  return n;
}

// TMEM accumulators -> fp32 tile [128][NP+4] in SS_A:  (D1 + D2/2048) * scale + add_col[col] + add_row * rowvec[col]
// Thread (w, l): row 32 (w%4) + l, columns [cq NP/4, (cq+1) NP/4), cq = w/4.  Returns this thread's partial
// of sum_col wcol[col] * value (the national average of epilogue 1), 0 if wcol is null.
__device__ __forceinline__ float s_epilogue(float scale, const float* add_col, float add_row, const float* rowvec, const float* wcol) {
  const ModelS& m = SMD();
  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  const int row = 32 * (w & 3) + l, cq = w >> 2, ncq = m.NP >> 2;
  const uint32_t taddr = SCTL().tmem_base + ((uint32_t)(32 * (w & 3)) << 16) + cq * ncq;
  float* out = sF(SS_A) + row * (m.NP + ST_PITCH_PAD) + cq * ncq;
  float acc = 0.f;
  for (int c0 = 0; c0 < ncq; c0 += 16) {
    float d1[16], d2[16];
    ptx::tmem_ld16f(taddr + c0, d1);
    ptx::tmem_ld16f(taddr + 256 + c0, d2);
    ptx::tmem_wait_ld();
    const int cb = cq * ncq + c0;
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float x = fmaf(d2[j + i], 1.0f / 2048.0f, d1[j + i]) * scale;
        if (add_col) x += add_col[cb + j + i];
        if (rowvec) x = fmaf(add_row, rowvec[cb + j + i], x);
        if (wcol) acc = fmaf(wcol[cb + j + i], x, acc);
        v[i] = x;
      }
      *reinterpret_cast<float4*>(out + c0 + j) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
  ptx::tc_fence_before();
  return acc;
}

// element-wise tail of a leaf for one parameter: g = dU/dtheta at q, p/s read from global by the caller
struct LeafAcc { float kk, c1a, c1b; };
__device__ __forceinline__ void s_leaf_elem(float q, float g, float p, float s, float lr, bool odd, float hs, float eps_s, float& qn, float& pn,
                                            float& P, LeafAcc& a) {
  P = fmaf(-hs * s, g, p);
  pn = 2.0f * P - p;
  qn = fmaf(eps_s * s, pn, q);
  a.kk = fmaf(P, P, a.kk);
  if (odd) { const float x = lr + P; a.c1a = fmaf(lr, x, a.c1a); a.c1b = fmaf(P, x, a.c1b); }
}

// ================================================================================================
// s_sweep: potential U (-> ctl.U) and gradient at io.qin; in leaf mode the whole leapfrog tail as well.
// ================================================================================================
template <bool LEAF>
__device__ __forceinline__ void s_sweep_body(SweepIO& io, float* rbuf) {
  const ModelS& m = SMD();
  SCtl& ctl = SCTL();
  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  const int S = m.S, T = m.T, SP = m.SP, NP = m.NP;
  const int pitch = NP + ST_PITCH_PAD;
  const float* qin = io.qin;
  const bool odd = LEAF && io.Lr != nullptr;
  const float eps_s = LEAF ? io.eps_s : 0.f, hs = 0.5f * eps_s;
  float* qnz = sF(SS_QNZ);
  float* gnz = sF(SS_GNZ);
  const int oz = m.o_zT;   // start of the small block in the vector; its element i lives at qnz[i]
  float qsq = 0.f, fsum = 0.f, gm[MAX_MODE] = {0.f, 0.f, 0.f, 0.f}, gp[MAX_MODE] = {0.f, 0.f, 0.f, 0.f};
  LeafAcc la{0.f, 0.f, 0.f};
  const int pr = tid & 127, qd = tid >> 7, c0 = 2 * pr;
  const bool colload = c0 < SP, colwrite = c0 < 16 * m.KS;
  SPROF_DECL;

  // ---------------- stage the small parameter block; zero accumulators
  for (int i = tid; i < m.nzs; i += SNT) { qnz[i] = qin[oz + i]; gnz[i] = 0.f; }
  for (int i = tid; i < ST_MAXT; i += SNT) sF(SS_GE)[i] = 0.f;
  if (tid < 256) { sF(SS_CARZ)[tid] = 0.f; sF(SS_CARH)[tid] = 0.f; }
  const bool have_cz = (ctl.czn_tag == qin);
  int npre = s_prefetch(0);
  __syncthreads();
  // ---------------- column sums of the walk innovations (days 0..T-2), unless the sweep that wrote qin left them
  if (!have_cz) {
    float s0 = 0.f, s1 = 0.f;
    if (colload) {
      for (int t = qd; t < T - 1; t += 4) { const float2 z = *reinterpret_cast<const float2*>(qin + (size_t)t * SP + c0); s0 += z.x; s1 += z.y; }
    }
    *reinterpret_cast<float2*>(sF(SS_QTOT) + qd * 256 + c0) = make_float2(s0, s1);
  }
  // AR(1) partisan non-response bias (warp 1), poll_model_2020.stan:91-93
  if (w == 1) {
    if (m.full) {
      const float* ze = qnz + (m.o_ze - oz);
      const float u_rho = qnz[m.o_urho - oz], u_mu = qnz[m.o_umu - oz];
      const float rho = 1.0f / (1.0f + __expf(-u_rho));
      const float mu_e = 0.02f * u_mu;
      const float s2 = sqrtf(fmaxf(1.0f - rho * rho, 0.f));
      const float sig_rho = s2 * m.sig_e;
      const float cst = mu_e * (1.0f - rho);
      const int per = (T + 31) >> 5;   // <= 16 days per lane
      float A = 1.f, B = 0.f;
      for (int j = 0; j < per; ++j) {
        const int t = per * l + j;
        if (t < T) { const float uj = (t == 0) ? m.sig_e * ze[0] : cst + sig_rho * ze[t]; B = rho * B + uj; A = rho * A; }
      }
      affine_scan(A, B, l);
      float ein = __shfl_up_sync(0xffffffffu, B, 1);
      if (l == 0) ein = 0.f;
      for (int j = 0; j < per; ++j) {
        const int t = per * l + j;
        if (t < T) { const float uj = (t == 0) ? m.sig_e * ze[0] : cst + sig_rho * ze[t]; ein = rho * ein + uj; sF(SS_E)[t] = ein; }
      }
      if (l == 0) {
        ctl.rho = rho; ctl.mu_e = mu_e; ctl.sig_rho = sig_rho; ctl.rho_term = m.sig_e * rho / fmaxf(s2, 1e-20f);
        const double r = 1.0 / (1.0 + exp(-(double)u_rho));
        ctl.u_extra = 0.5 * ((r - 0.7) / 0.1) * ((r - 0.7) / 0.1) - log(r) - log1p(-r);
      }
    } else if (l == 0) {
      ctl.u_extra = 0.0;
    }
  }
  // polling_bias = a_b L0 z_b (stan:77): two threads per state split the k range; fixed order
  {
    const int s = tid & 255, half = tid >> 8;
    float acc = 0.f;
    if (s < S) {
      const float* zb = qnz + (m.o_zb - oz);
      const int kmid = (s + 1) >> 1, k0 = half ? kmid : 0, k1 = half ? s + 1 : kmid;
#pragma unroll 8
      for (int k = k0; k < k1; ++k) acc = fmaf(__ldg(m.l0t + (size_t)k * SP + s), zb[k], acc);
    }
    sF(SS_NATP)[tid] = acc;   // scratch [2][256]
  }
  __syncthreads();
  if (tid < 256) {
    float pb = 0.f, cz = 0.f;
    if (tid < S) pb = m.a_b * (sF(SS_NATP)[tid] + sF(SS_NATP)[256 + tid]);
    sF(SS_PB)[tid] = pb;
    if (have_cz) cz = sF(SS_CZN)[tid];
    else cz = (sF(SS_QTOT)[tid] + sF(SS_QTOT)[256 + tid]) + (sF(SS_QTOT)[512 + tid] + sF(SS_QTOT)[768 + tid]);
    float base = 0.f;
    if (tid < S) base = fmaf(m.a_T, qnz[(m.o_zT - oz) + tid], m.a_w * cz);
    sF(SS_BASE)[tid] = base;
  }
  __syncthreads();
  if (w == 0) {   // national_polling_bias_average (stan:79)
    float v = 0.f;
    for (int s = l; s < S; s += 32) v = fmaf(sF(SS_W)[s], sF(SS_PB)[s], v);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (l == 0) ctl.nat_pb = v;
  }
  if (tid < 256) sF(SS_CZN)[tid] = 0.f;   // from here on: column sums of the position being written
  __syncthreads();
  SPROF(0);

  // ================================================================ tiles of 128 days
  for (int tile = 0; tile < m.NTILE; ++tile) {
    const int t0 = tile * ST_ROWS;
    const int p0 = __ldg(m.tile_ptr + tile), p1 = __ldg(m.tile_ptr + tile + 1);
    // pull what the poll phase and the update phase of THIS tile will read into L2 now (the state vectors of 148 chains are
    // far larger than L2, so every sweep streams them from HBM), and the innovations of the NEXT tile for its phase A
    if (tid == 64) {
      const int nrow = min(ST_ROWS, T - t0);
      const uint32_t wb = (uint32_t)nrow * SP * 4, xb = (uint32_t)(((p1 - p0) * 4 + 15) & ~15);
      const size_t w0 = (size_t)t0 * SP, x0 = ((size_t)m.o_x + p0) & ~(size_t)3;
      if (LEAF) {
        ptx::prefetch_l2_bulk(io.ph + w0, wb); ptx::prefetch_l2_bulk(io.sm + w0, wb);
        if (odd) ptx::prefetch_l2_bulk(io.Lr + w0, wb);
        if (xb) { ptx::prefetch_l2_bulk(io.ph + x0, xb); ptx::prefetch_l2_bulk(io.sm + x0, xb); if (odd) ptx::prefetch_l2_bulk(io.Lr + x0, xb); }
      }
      if (xb) ptx::prefetch_l2_bulk(qin + x0, xb);
      if (tile + 1 < m.NTILE) ptx::prefetch_l2_bulk(qin + (size_t)(t0 + ST_ROWS) * SP, (uint32_t)min(ST_ROWS, T - t0 - ST_ROWS) * SP * 4);
    }
    // ---------------- A: W tile -> fp16 hi/lo planes.  W[t] = a_T zT + a_w (colsum - sum_{u<t} Z[u])
    {
      // pass 1: day-quarter totals of the innovations (rows >= T-1 carry none)
      float r0 = 0.f, r1 = 0.f;
      const float* zp = qin + (size_t)(t0 + 32 * qd) * SP + c0;
      const int nrow_z = min(32, max(0, T - 1 - (t0 + 32 * qd)));   // rows of this thread that hold a walk innovation
      if (colload) {
#pragma unroll
        for (int d0 = 0; d0 < 32; d0 += 8) {
          float2 z[8];
#pragma unroll
          for (int d = 0; d < 8; ++d) z[d] = (d0 + d < nrow_z) ? *reinterpret_cast<const float2*>(zp + (size_t)(d0 + d) * SP) : make_float2(0.f, 0.f);
#pragma unroll
          for (int d = 0; d < 8; ++d) { r0 += z[d].x; r1 += z[d].y; }
        }
      }
      *reinterpret_cast<float2*>(sF(SS_QTOT) + qd * 256 + c0) = make_float2(r0, r1);
      __syncthreads();
      float2 off = *reinterpret_cast<const float2*>(sF(SS_CARZ) + c0);
      for (int q2 = 0; q2 < qd; ++q2) { const float2 t2 = *reinterpret_cast<const float2*>(sF(SS_QTOT) + q2 * 256 + c0); off.x += t2.x; off.y += t2.y; }
      const float2 base = *reinterpret_cast<const float2*>(sF(SS_BASE) + c0);
      // pass 2: the same rows again (L2 hits), running prefix, split, store
      if (colwrite) {
        unsigned char* ap = smem_raw + SS_A + (uint32_t)(c0 >> 3) * SA_LBO + (uint32_t)(c0 & 7) * 2 + (uint32_t)(4 * qd) * SA_SBO;
        float e0 = off.x, e1 = off.y;   // sum of the innovations of all earlier days
        const int nrow = min(32, max(0, T - (t0 + 32 * qd)));
#pragma unroll
        for (int d0 = 0; d0 < 32; d0 += 8) {
          float2 z[8];
#pragma unroll
          for (int d = 0; d < 8; ++d) z[d] = (colload && d0 + d < nrow_z) ? *reinterpret_cast<const float2*>(zp + (size_t)(d0 + d) * SP) : make_float2(0.f, 0.f);
#pragma unroll
          for (int d = 0; d < 8; ++d) {
            float v0 = 0.f, v1 = 0.f;
            if (d0 + d < nrow) { v0 = fmaf(-m.a_w, e0, base.x); v1 = fmaf(-m.a_w, e1, base.y); }
            e0 += z[d].x; e1 += z[d].y;
            if (c0 >= S) v0 = 0.f;
            if (c0 + 1 >= S) v1 = 0.f;
            const __half2 hi = __floats2half2_rn(v0, v1);
            const float2 hf = __half22float2(hi);
            const __half2 lo = __floats2half2_rn((v0 - hf.x) * 2048.0f, (v1 - hf.y) * 2048.0f);
            const uint32_t o = (uint32_t)((d0 + d) >> 3) * SA_SBO + (uint32_t)((d0 + d) & 7) * 16;
            *reinterpret_cast<__half2*>(ap + o) = hi;
            *reinterpret_cast<__half2*>(ap + SA_PLANE + o) = lo;
          }
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      __syncthreads();
      if (qd == 3) *reinterpret_cast<float2*>(sF(SS_CARZ) + c0) = make_float2(off.x + r0, off.y + r1);
    }
    SPROF(1);
    // ---------------- B: mu_b^T = W^T X^T
    s_gemm(0, npre);
    SPROF(2);
    // ---------------- C: mu_b tile (+ prior) and the national average of each day (stan:87)
    {
      const float part = s_epilogue(1.0f / 256.0f, sF(SS_PRIOR), 0.f, nullptr, sF(SS_W));
      sF(SS_NATP)[(w >> 2) * 128 + 32 * (w & 3) + l] = part;
      __syncthreads();
      if (tid == 0) ctl.gemm_cnt++;
      if (tid < 128) sF(SS_NAT)[tid] = (sF(SS_NATP)[tid] + sF(SS_NATP)[128 + tid]) + (sF(SS_NATP)[256 + tid] + sF(SS_NATP)[384 + tid]);
      npre = s_prefetch(1);
      __syncthreads();
    }
    SPROF(3);
    // ---------------- optional outputs of this point (transformed parameters of the days of this tile)
    if (!LEAF) {
      const int nrow = min(ST_ROWS, T - t0);
      if (io.em.draw != nullptr) {
        float* o = io.em.draw + (size_t)S * t0;
        for (int i = tid; i < S * nrow; i += SNT) { const int t = i / S, s = i - t * S; o[i] = sF(SS_A)[t * pitch + s]; }
      }
      if (io.em.monitor != nullptr && T - 1 >= t0 && T - 1 < t0 + ST_ROWS) {
        const int rr = T - 1 - t0;
        for (int i = tid; i < S; i += SNT) io.em.monitor[i] = sF(SS_A)[rr * pitch + i];
        if (tid == 0) io.em.monitor[S] = sF(SS_NAT)[rr];
      }
    }
    // ---------------- D: polls of these days: linear predictor (stan:95-112), centred binomial_logit, residuals;
    //                  the measurement-noise parameters are poll-local, so their leapfrog tail happens right here
    {
      const float* mu = sF(SS_A);
      const float nat_pb = ctl.nat_pb;
      float* rpol = rbuf + m.rp_off;
#ifndef POTUS_UD
#define POTUS_UD 2
#endif
      constexpr int UD = POTUS_UD;   // polls per thread in flight (more costs registers, and spills go to L2: there is almost no L1 left)
      static_assert(UD == 2, "the branch-free main loop below is written for two polls per thread");
      const float fullf = m.full ? 1.0f : 0.f;
      int kb = p0 + tid;
      // main part: both polls of the batch exist.  One basic block for the common case (no bounds tests, selects instead of
      // branches, the series branch of the poll term for both polls) so that the two polls' dependency chains can be
      // scheduled into each other; polls outside the series' range are redone by the wide form afterwards (rare once a
      // chain is near its typical set)
      for (; kb + SNT < p1; kb += 2 * SNT) {
        const int k0 = kb, k1 = kb + SNT;
        const uint32_t wa = __ldg(m.pw0 + k0), wb = __ldg(m.pw0 + k1), pma = __ldg(m.perm + k0), pmb = __ldg(m.perm + k1);
        const float4 pca = __ldg(m.pc + k0), pcb = __ldg(m.pc + k1);
        const size_t ea = (size_t)m.o_x + k0, eb = (size_t)m.o_x + k1;
        const float xa = qin[ea], xb = qin[eb];
        float pva = 0.f, pvb = 0.f, sva = 0.f, svb = 0.f, lra = 0.f, lrb = 0.f;
        if (LEAF) { pva = io.ph[ea]; pvb = io.ph[eb]; sva = io.sm[ea]; svb = io.sm[eb]; if (odd) { lra = io.Lr[ea]; lrb = io.Lr[eb]; } }
        const float una = ((wa >> 20) & 1) ? (m.pun ? __ldg(m.pun + k0) : 1.0f) : 0.f, unb = ((wb >> 20) & 1) ? (m.pun ? __ldg(m.pun + k1) : 1.0f) : 0.f;
        const int sa = wa & 511, sb = wb & 511, dla = (wa >> 9) & 127, dlb = (wb >> 9) & 127;
        const int moa = (wa >> 16) & 3, mob = (wb >> 16) & 3, poa = (wa >> 18) & 3, pob = (wb >> 18) & 3;
        const bool nata = (sa == S), natb = (sb == S);
        const float sga = nata ? m.sig_n : m.sig_s, sgb = natb ? m.sig_n : m.sig_s;
        // (national polls: mu[dl][S] and PB[S] are valid shared-memory addresses whose values the select discards)
        float etaa = (nata ? sF(SS_NAT)[dla] + nat_pb : mu[dla * pitch + sa] + sF(SS_PB)[sa]) + m.sig_c * qnz[(m.o_c - oz) + ((wa >> 21) & 511)] + sga * xa;
        float etab = (natb ? sF(SS_NAT)[dlb] + nat_pb : mu[dlb * pitch + sb] + sF(SS_PB)[sb]) + m.sig_c * qnz[(m.o_c - oz) + ((wb >> 21) & 511)] + sgb * xb;
        etaa += fullf * (m.sig_m * qnz[(m.o_m - oz) + moa] + m.sig_pop * qnz[(m.o_pop - oz) + poa]);
        etab += fullf * (m.sig_m * qnz[(m.o_m - oz) + mob] + m.sig_pop * qnz[(m.o_pop - oz) + pob]);
        etaa = fmaf(una, sF(SS_E)[t0 + dla], etaa);
        etab = fmaf(unb, sF(SS_E)[t0 + dlb], etab);
        const float da = etaa - pca.y, db = etab - pcb.y;
        float fa, ra, fb, rb;
        poll_term_series(da, pca.x, pca.z, pca.w, fa, ra);
        poll_term_series(db, pcb.x, pcb.z, pcb.w, fb, rb);
        if (!(fabsf(da) < 0.4f)) poll_term_wide_cold(etaa, pca.x, pca.y, pca.z, pca.w, fa, ra);
        if (!(fabsf(db) < 0.4f)) poll_term_wide_cold(etab, pcb.x, pcb.y, pcb.z, pcb.w, fb, rb);
        fsum += fa; fsum += fb;
        rbuf[k0] = ra; rbuf[k1] = rb;
        rpol[pma] = ra; rpol[pmb] = rb;
        qsq = fmaf(xa, xa, qsq); qsq = fmaf(xb, xb, qsq);
        const float gxa = xa - sga * ra, gxb = xb - sgb * rb;
        if (LEAF) {
          float qn, pn, P;
          s_leaf_elem(xa, gxa, pva, sva, lra, odd, hs, eps_s, qn, pn, P, la);
          io.qout[ea] = qn; io.ph[ea] = pn; io.Pdst[ea] = P;
          s_leaf_elem(xb, gxb, pvb, svb, lrb, odd, hs, eps_s, qn, pn, P, la);
          io.qout[eb] = qn; io.ph[eb] = pn; io.Pdst[eb] = P;
        } else {
          io.gout[ea] = gxa; io.gout[eb] = gxb;
        }
#pragma unroll
        for (int j = 0; j < MAX_MODE - 1; ++j) {
          gm[j] += (moa == j) ? ra : 0.f; gp[j] += (poa == j) ? ra : 0.f;
          gm[j] += (mob == j) ? rb : 0.f; gp[j] += (pob == j) ? rb : 0.f;
        }
        gm[MAX_MODE - 1] += ra; gm[MAX_MODE - 1] += rb;
      }
      // tail: at most one poll of this thread is left
      if (kb < p1) {
        const int k = kb;
        const uint32_t w0 = __ldg(m.pw0 + k), pm = __ldg(m.perm + k);
        const float4 pc = __ldg(m.pc + k);
        const size_t e = (size_t)m.o_x + k;
        const float x = qin[e];
        const int s = w0 & 511, dl = (w0 >> 9) & 127, mo = (w0 >> 16) & 3, po = (w0 >> 18) & 3, pp = (w0 >> 21) & 511;
        const bool nat = (s == S);
        const float sigx = nat ? m.sig_n : m.sig_s;
        float eta = (nat ? sF(SS_NAT)[dl] + nat_pb : mu[dl * pitch + s] + sF(SS_PB)[s]) + m.sig_c * qnz[(m.o_c - oz) + pp] + sigx * x;
        if (m.full) {
          eta += m.sig_m * qnz[(m.o_m - oz) + mo] + m.sig_pop * qnz[(m.o_pop - oz) + po];
          if ((w0 >> 20) & 1) eta = fmaf(m.pun ? __ldg(m.pun + k) : 1.0f, sF(SS_E)[t0 + dl], eta);
        }
        float f, r;
        poll_term(eta, pc.x, pc.y, pc.z, pc.w, f, r);
        fsum += f;
        rbuf[k] = r;
        rpol[pm] = r;
        qsq = fmaf(x, x, qsq);
        const float gx = x - sigx * r;
        if (LEAF) {
          float qn, pn, P;
          s_leaf_elem(x, gx, io.ph[e], io.sm[e], odd ? io.Lr[e] : 0.f, odd, hs, eps_s, qn, pn, P, la);
          io.qout[e] = qn; io.ph[e] = pn; io.Pdst[e] = P;
        } else {
          io.gout[e] = gx;
        }
#pragma unroll
        for (int j = 0; j < MAX_MODE - 1; ++j) { gm[j] += (mo == j) ? r : 0.f; gp[j] += (po == j) ? r : 0.f; }
        gm[MAX_MODE - 1] += r;
      }
    }
    SPROF(4);
    __syncthreads();
    // ---------------- E1: clear the operand planes (the mu_b tile is dead)
    {
      uint4* a4 = SMP(uint4, SS_A);
      const uint4 z = make_uint4(0, 0, 0, 0);
      for (int i = tid; i < (int)(2 * SA_PLANE / 16); i += SNT) a4[i] = z;
    }
    __syncthreads();
    SPROF(5);
    // ---------------- E2: G operand cells = sums of residuals per (state, day); per-day sums for e_bias / national
    {
      {   // one thread per cell descriptor: the run's residuals are contiguous in the sorted residual buffer
        const int c0 = __ldg(m.tile_cptr + tile), c1 = __ldg(m.tile_cptr + tile + 1);
        for (int cb = c0 + tid; cb < c1; cb += 2 * SNT) {
          uint2 cd[2];
          float acc[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int c = cb + u * SNT;
            cd[u] = make_uint2(0u, 0u); acc[u] = 0.f;
            if (c < c1) { cd[u] = __ldg(m.cells + c); }
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) if (cb + u * SNT < c1) acc[u] = rbuf[cd[u].x];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (cb + u * SNT < c1) {
              const int len = cd[u].y >> 16;
              for (int j = 1; j < len; ++j) acc[u] += rbuf[cd[u].x + j];
              const int s = cd[u].y & 511, dl = (cd[u].y >> 9) & 127;
              __half hi, lo;
              ptx::split_f16(acc[u] * m.scale_G, hi, lo);
              const uint32_t o = (uint32_t)(s >> 3) * SA_LBO + (uint32_t)(dl >> 3) * SA_SBO + (uint32_t)(dl & 7) * 16 + (uint32_t)(s & 7) * 2;
              *reinterpret_cast<__half*>(smem_raw + SS_A + o) = hi;
              *reinterpret_cast<__half*>(smem_raw + SS_A + SA_PLANE + o) = lo;
            }
          }
        }
      }
      SPROF(6);
      for (int dl = w; dl < ST_ROWS; dl += 16) {   // per-day sums: one warp per day, lanes stride over its polls (fixed order)
        const int t = t0 + dl;
        float ae = 0.f, an = 0.f;
        if (t < T) {
          const int k0 = __ldg(m.day_ptr + t), k1 = __ldg(m.day_ptr + t + 1);
          for (int kb = k0 + l; kb < k1; kb += 8 * 32) {
            uint32_t w8[8];
            float r8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int k = kb + 32 * u; w8[u] = 0u; r8[u] = 0.f; if (k < k1) { w8[u] = __ldg(m.pw0 + k); r8[u] = rbuf[k]; } }
#pragma unroll
            for (int u = 0; u < 8; ++u) { if ((w8[u] >> 20) & 1) ae = fmaf(m.pun ? __ldg(m.pun + kb + 32 * u) : 1.0f, r8[u], ae); if ((int)(w8[u] & 511) == S) an += r8[u]; }
          }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) { ae += __shfl_xor_sync(0xffffffffu, ae, off); an += __shfl_xor_sync(0xffffffffu, an, off); }
        if (l == 0) { sF(SS_RND)[dl] = an; if (t < T) sF(SS_GE)[t] = ae; }
      }
    }
    SPROF(7);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    // ---------------- F: H^T = G^T X
    s_gemm(1, npre);
    SPROF(8);
    // ---------------- G: H tile; the national polls enter as the rank-1 term  rn_day[t] * (L0^T w)
    {
      const int row = 32 * (w & 3) + l;
      s_epilogue(m.inv_scale_G * (1.0f / 256.0f), nullptr, sF(SS_RND)[row], sF(SS_LW), nullptr);
      __syncthreads();
      if (tid == 0) ctl.gemm_cnt++;
      npre = (tile + 1 < m.NTILE) ? s_prefetch(0) : 0;
    }
    SPROF(9);
    // ---------------- H: prefix of H over days -> gradient of the walk block -> leapfrog tail
    {
      const float* hp = sF(SS_A) + (32 * qd) * pitch + c0;
      float r0 = 0.f, r1 = 0.f;
      if (c0 < NP) {
#pragma unroll 8
        for (int d = 0; d < 32; ++d) { const float2 h = *reinterpret_cast<const float2*>(hp + d * pitch); r0 += h.x; r1 += h.y; }
      }
      *reinterpret_cast<float2*>(sF(SS_QTOT) + qd * 256 + c0) = make_float2(r0, r1);
      __syncthreads();
      float2 pre = *reinterpret_cast<const float2*>(sF(SS_CARH) + c0);
      for (int q2 = 0; q2 < qd; ++q2) { const float2 t2 = *reinterpret_cast<const float2*>(sF(SS_QTOT) + q2 * 256 + c0); pre.x += t2.x; pre.y += t2.y; }
      const float2 ncar = make_float2(pre.x + r0, pre.y + r1);   // (qd == 3: prefix through the last row of this tile)
      float cz0 = 0.f, cz1 = 0.f;
      if (colload) {
        const size_t e0 = (size_t)(t0 + 32 * qd) * SP + c0;
        const int nrow = min(32, max(0, T - (t0 + 32 * qd)));
#ifndef POTUS_UH
#define POTUS_UH 2
#endif
        constexpr int UH = POTUS_UH;   // days per batch in flight (register budget; the rows were prefetched into L2 at the start of the tile)
#pragma unroll 1
        for (int d0 = 0; d0 < nrow; d0 += UH) {
          float2 zz[UH], pp[UH], ss[UH], ll[UH];
#pragma unroll
          for (int u = 0; u < UH; ++u) {
            const size_t e = e0 + (size_t)(d0 + u) * SP;
            zz[u] = make_float2(0.f, 0.f); pp[u] = zz[u]; ss[u] = zz[u]; ll[u] = zz[u];
            if (d0 + u < nrow) {
              zz[u] = *reinterpret_cast<const float2*>(qin + e);
              if (LEAF) {
                pp[u] = *reinterpret_cast<const float2*>(io.ph + e); ss[u] = *reinterpret_cast<const float2*>(io.sm + e);
                if (odd) ll[u] = *reinterpret_cast<const float2*>(io.Lr + e);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < UH; ++u) {
            const int d = d0 + u;
            if (d < nrow) {
              const int t = t0 + 32 * qd + d;
              const size_t e = e0 + (size_t)d * SP;
              const float2 h = *reinterpret_cast<const float2*>(hp + d * pitch);
              pre.x += h.x; pre.y += h.y;
              const float2 z = zz[u];
              qsq = fmaf(z.x, z.x, fmaf(z.y, z.y, qsq));
              const bool walk = t < T - 1;
              const float g0 = walk ? fmaf(-m.a_w, pre.x, z.x) : z.x, g1 = walk ? fmaf(-m.a_w, pre.y, z.y) : z.y;
              if (LEAF) {
                float2 qn, pn, P;
                s_leaf_elem(z.x, g0, pp[u].x, ss[u].x, ll[u].x, odd, hs, eps_s, qn.x, pn.x, P.x, la);
                s_leaf_elem(z.y, g1, pp[u].y, ss[u].y, ll[u].y, odd, hs, eps_s, qn.y, pn.y, P.y, la);
                *reinterpret_cast<float2*>(io.qout + e) = qn;
                *reinterpret_cast<float2*>(io.ph + e) = pn;
                *reinterpret_cast<float2*>(io.Pdst + e) = P;
                if (walk) { cz0 += qn.x; cz1 += qn.y; }
              } else {
                *reinterpret_cast<float2*>(io.gout + e) = make_float2(g0, g1);
              }
            }
          }
        }
      }
      __syncthreads();   // every read of CARH / QTOT / the H tile is done
      if (qd == 3) *reinterpret_cast<float2*>(sF(SS_CARH) + c0) = ncar;
      if (LEAF) *reinterpret_cast<float2*>(sF(SS_QTOT) + qd * 256 + c0) = make_float2(cz0, cz1);
      __syncthreads();
      if (LEAF) {
        if (tid < 256) sF(SS_CZN)[tid] += (sF(SS_QTOT)[tid] + sF(SS_QTOT)[256 + tid]) + (sF(SS_QTOT)[512 + tid] + sF(SS_QTOT)[768 + tid]);
        __syncthreads();   // QTOT is rewritten by the next tile's phase A
      }
    }
  }
  // ================================================================ after the last tile
  SPROF(10);
  // class sums of residuals (mode / population), fixed shuffle trees
  if (m.full) {
#pragma unroll
    for (int j = 0; j < MAX_MODE; ++j) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        gm[j] += __shfl_xor_sync(0xffffffffu, gm[j], off);
        if (j < MAX_MODE - 1) gp[j] += __shfl_xor_sync(0xffffffffu, gp[j], off);
      }
    }
    if (l == 0) {
      double* red = SMP(double, SS_RED) + w * 16;
#pragma unroll
      for (int j = 0; j < MAX_MODE; ++j) { red[j] = (double)gm[j]; red[4 + j] = (double)gp[j]; }
    }
  }
  // pollster sums, level 1: one thread per 16-residual segment of the pollster-grouped copy (pads are zero), fixed order
  {
    const float* rpol = rbuf + m.rp_off;
    float* psum = sF(SS_A);   // the tile region is free after the last tile
    for (int sg = tid; sg < m.n_seg; sg += SNT) {
      const float4* v = reinterpret_cast<const float4*>(rpol + (size_t)sg * ST_SEGL);
      const float4 a = v[0], b = v[1], c = v[2], d = v[3];
      psum[sg] = (((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w))) + (((c.x + c.y) + (c.z + c.w)) + ((d.x + d.y) + (d.z + d.w)));
    }
  }
  if (tid < S) {   // raw_mu_b_T and raw_polling_bias: both driven by sum_t H[:,t] = L0^T g_pb
    const float h = sF(SS_CARH)[tid];
    gnz[(m.o_zT - oz) + tid] = m.a_T * h;
    gnz[(m.o_zb - oz) + tid] = m.a_b * h;
  }
  __syncthreads();
  for (int p = w; p < m.P; p += 16) {   // level 2: one warp per pollster over its segment sums
    const int a = __ldg(m.seg_ptr + p), b = __ldg(m.seg_ptr + p + 1);
    float acc = 0.f;
    for (int j = a + l; j < b; j += 32) acc += sF(SS_A)[j];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (l == 0) gnz[(m.o_c - oz) + p] = m.sig_c * acc;
  }
  if (m.full && w == 2 && l < 2 * MAX_MODE) {
    const int j = l % MAX_MODE;
    const bool ispop = l >= MAX_MODE;
    const int ncls = ispop ? m.Pop : m.M;
    if (j < ncls) {
      const double* red = SMP(double, SS_RED);
      double cls[MAX_MODE] = {0, 0, 0, 0}, tot = 0;
      for (int w2 = 0; w2 < 16; ++w2) {
#pragma unroll
        for (int q2 = 0; q2 < MAX_MODE - 1; ++q2) cls[q2] += red[w2 * 16 + (ispop ? 4 : 0) + q2];
        tot += red[w2 * 16 + (MAX_MODE - 1)];
      }
      double v = (j == MAX_MODE - 1) ? tot - cls[0] - cls[1] - cls[2] : cls[j];
      gnz[((ispop ? m.o_pop : m.o_m) - oz) + j] = (ispop ? m.sig_pop : m.sig_m) * (float)v;
    }
  }
  if (m.full && w == 1) {   // adjoint of the AR(1) recurrence -> gradients of raw_e_bias, mu_e_bias, rho_e_bias
    const float rho = ctl.rho, mu_e = ctl.mu_e, sig_rho = ctl.sig_rho, rterm = ctl.rho_term;
    const float* ze = qnz + (m.o_ze - oz);
    float* gze = gnz + (m.o_ze - oz);
    const int per = (T + 31) >> 5;
    float A = 1.f, B = 0.f;   // reversed order: lane l handles t = T-1-(per*l+j)
    for (int j = 0; j < per; ++j) {
      const int t = T - 1 - (per * l + j);
      if (t >= 0) { B = rho * B + sF(SS_GE)[t]; A = rho * A; }
    }
    affine_scan(A, B, l);
    float ein = __shfl_up_sync(0xffffffffu, B, 1);
    if (l == 0) ein = 0.f;
    float s_mu = 0.f, s_rho = 0.f;
    for (int j = 0; j < per; ++j) {
      const int t = T - 1 - (per * l + j);
      if (t >= 0) {
        ein = rho * ein + sF(SS_GE)[t];
        if (t >= 1) {
          gze[t] = sig_rho * ein;
          s_mu += ein;
          s_rho += ein * ((sF(SS_E)[t - 1] - mu_e) - ze[t] * rterm);
        } else {
          gze[0] = m.sig_e * ein;
        }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { s_mu += __shfl_xor_sync(0xffffffffu, s_mu, off); s_rho += __shfl_xor_sync(0xffffffffu, s_rho, off); }
    if (l == 0) {
      gnz[m.o_umu - oz] = 0.02f * (1.0f - rho) * s_mu;
      const float d_rho = s_rho - (rho - 0.7f) * 100.0f;
      // stored as (d lp/d u_rho + u_rho) so that the generic "theta - gnz" form yields -d lp/d u_rho
      gnz[m.o_urho - oz] = rho * (1.0f - rho) * d_rho + (1.0f - 2.0f * rho) + qnz[m.o_urho - oz];
    }
  }
  __syncthreads();
  // the small block: gradient and leapfrog tail
  for (int i = tid; i < m.nzs; i += SNT) {
    const float th = qnz[i], g = th - gnz[i];
    if (!(m.full && i == m.o_urho - oz)) qsq = fmaf(th, th, qsq);
    const size_t e = (size_t)oz + i;
    if (LEAF) {
      float qn, pn, P;
      s_leaf_elem(th, g, io.ph[e], io.sm[e], odd ? io.Lr[e] : 0.f, odd, hs, eps_s, qn, pn, P, la);
      io.qout[e] = qn; io.ph[e] = pn; io.Pdst[e] = P;
    } else {
      io.gout[e] = g;
    }
  }
  if (!LEAF && io.em.draw != nullptr) {
    float* o = io.em.draw + (size_t)S * T;
    for (int i = tid; i < m.P; i += SNT) o[i] = m.sig_c * qnz[(m.o_c - oz) + i];
    o += m.P;
    for (int i = tid; i < m.M; i += SNT) o[i] = m.sig_m * qnz[(m.o_m - oz) + i];
    o += m.M;
    for (int i = tid; i < m.Pop; i += SNT) o[i] = m.sig_pop * qnz[(m.o_pop - oz) + i];
    o += m.Pop;
    for (int i = tid; i < T; i += SNT) o[i] = m.full ? sF(SS_E)[i] : 0.f;
    o += T;
    for (int i = tid; i < S; i += SNT) o[i] = sF(SS_PB)[i];
    o += S;
    for (int i = tid; i < m.VL; i += SNT) { const int si = __ldg(m.map_i2s + i); if (si >= 0) o[si] = qin[i]; }
  }
  // block totals: U (fp64), and in leaf mode |P|^2 and the two level-0 U-turn sums
  {
    double v0 = 0.5 * (double)qsq - (double)fsum;
    float k3[3] = {la.kk, la.c1a, la.c1b};
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      v0 += __shfl_xor_sync(0xffffffffu, v0, off);
#pragma unroll
      for (int i = 0; i < 3; ++i) k3[i] += __shfl_xor_sync(0xffffffffu, k3[i], off);
    }
    __syncthreads();   // (SS_RED class sums were consumed above)
    if (l == 0) { double* red = SMP(double, SS_RED) + w * 16; red[0] = v0; red[1] = (double)k3[0]; red[2] = (double)k3[1]; red[3] = (double)k3[2]; }
    __syncthreads();
    const double* red = SMP(double, SS_RED);
    double tot[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double a = 0;
#pragma unroll
      for (int w2 = 0; w2 < 16; ++w2) a += red[w2 * 16 + i];
      tot[i] = a;
    }
    if (tid == 0) { ctl.U = tot[0] + ctl.u_extra; ctl.czn_tag = LEAF ? io.qout : nullptr; }
    io.kk = (float)tot[1]; io.c1a = (float)tot[2]; io.c1b = (float)tot[3];
    __syncthreads();
  }
  SPROF(11);
}

__device__ __noinline__ void s_eval(SweepIO& io, float* rbuf) { s_sweep_body<false>(io, rbuf); }
__device__ __noinline__ void s_leaf(SweepIO& io, float* rbuf) { s_sweep_body<true>(io, rbuf); }

// ================================================================================================
// vector helpers over stream-layout vectors (float4 loops, fixed per-thread order => deterministic)
// ================================================================================================
__device__ __forceinline__ int nv4() { return SMD().VL >> 2; }
__device__ __forceinline__ float4 ld4(const float* v, int i) { return reinterpret_cast<const float4*>(v)[i]; }
__device__ __forceinline__ void st4(float* v, int i, float4 x) { reinterpret_cast<float4*>(v)[i] = x; }
__device__ __forceinline__ float* s_slot(float* ws, int slot) { return ws + (size_t)slot * SMD().VL; }

// block sum of up to 6 floats: every thread gets the totals (fixed order)
template <int N>
__device__ __forceinline__ void s_block_sum(float (&v)[N]) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
  }
  SCtl& ctl = SCTL();
  __syncthreads();
  if ((tid & 31) == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) ctl.mred[tid >> 5][i] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float s = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < 16; ++w2) s += ctl.mred[w2][i];
    v[i] = s;
  }
}

// (every vector length is a multiple of 4 * SNT float4s: loops are unrolled so that all loads of a batch are in flight together)
__device__ __forceinline__ void s_copy(float* dst, const float* src) {
  const int n = nv4();
  for (int i = threadIdx.x; i < n; i += 4 * SNT) {
    const float4 a = ld4(src, i), b = ld4(src, i + SNT), c = ld4(src, i + 2 * SNT), d = ld4(src, i + 3 * SNT);
    st4(dst, i, a); st4(dst, i + SNT, b); st4(dst, i + 2 * SNT, c); st4(dst, i + 3 * SNT, d);
  }
}
// dst = a + b
__device__ __forceinline__ void s_add(float* dst, const float* a, const float* b) {
  const int n = nv4();
  for (int i = threadIdx.x; i < n; i += 4 * SNT) {
    float4 x[4], y[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { x[u] = ld4(a, i + u * SNT); y[u] = ld4(b, i + u * SNT); }
#pragma unroll
    for (int u = 0; u < 4; ++u) st4(dst, i + u * SNT, make_float4(x[u].x + y[u].x, x[u].y + y[u].y, x[u].z + y[u].z, x[u].w + y[u].w));
  }
}

// fresh whitened momentum P ~ N(0, I) on the valid slots; returns |P|^2 (block total)
__device__ __forceinline__ float s_draw_momentum(const SRunArgs& a, float* dst, uint32_t chain_gid, uint32_t iter, uint32_t stream, uint32_t sub) {
  const ModelS& m = SMD();
  float ss[1] = {0.f};
  for (int i = threadIdx.x; i < m.VL; i += SNT) {
    const int si = __ldg(m.map_i2s + i);
    const float v = (si >= 0) ? rng_normal(a.seed, chain_gid, (uint32_t)si, iter, stream, sub) : 0.f;
    dst[i] = v;
    ss[0] = fmaf(v, v, ss[0]);
  }
  s_block_sum(ss);
  return ss[0];
}

// One U-turn merge (Stan's three criteria) between the completed left subtree L = {b, e, r} and the implicit right
// subtree R = {b: Rb (or P if null), e: P, r: P + S}.  S_in null means S = 0; L.r null means L.r = L.b + L.e (level 1).
// Writes S_out = S + L.r.  Same arithmetic as merge_check() of the resident kernel.
__device__ __forceinline__ bool s_merge(const float* Lb, const float* Le, const float* Lr, const float* Rb, const float* P, const float* S_in,
                                        float* S_out) {
  const int n = nv4();
  float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i0 = threadIdx.x; i0 < n; i0 += 2 * SNT) {
    float4 lb4[2], le4[2], p4[2], lr4[2], s4[2], rb4[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = i0 + u * SNT;
      lb4[u] = ld4(Lb, i); le4[u] = ld4(Le, i); p4[u] = ld4(P, i);
      if (Lr) lr4[u] = ld4(Lr, i);
      s4[u] = S_in ? ld4(S_in, i) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (Rb) rb4[u] = ld4(Rb, i);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int i = i0 + u * SNT;
      if (!Lr) lr4[u] = make_float4(lb4[u].x + le4[u].x, lb4[u].y + le4[u].y, lb4[u].z + le4[u].z, lb4[u].w + le4[u].w);
      if (!Rb) rb4[u] = p4[u];
      const float lb[4] = {lb4[u].x, lb4[u].y, lb4[u].z, lb4[u].w}, le[4] = {le4[u].x, le4[u].y, le4[u].z, le4[u].w}, p[4] = {p4[u].x, p4[u].y, p4[u].z, p4[u].w};
      const float lr[4] = {lr4[u].x, lr4[u].y, lr4[u].z, lr4[u].w}, rb[4] = {rb4[u].x, rb4[u].y, rb4[u].z, rb4[u].w};
      float sv[4] = {s4[u].x, s4[u].y, s4[u].z, s4[u].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float x = lr[j] + sv[j] + p[j];
        v[0] = fmaf(lb[j], x, v[0]); v[1] = fmaf(p[j], x, v[1]);
        const float y = lr[j] + rb[j];
        v[2] = fmaf(lb[j], y, v[2]); v[3] = fmaf(rb[j], y, v[3]);
        const float z = sv[j] + p[j] + le[j];
        v[4] = fmaf(le[j], z, v[4]); v[5] = fmaf(p[j], z, v[5]);
        sv[j] += lr[j];
      }
      st4(S_out, i, make_float4(sv[0], sv[1], sv[2], sv[3]));
    }
  }
  s_block_sum(v);
  return v[0] > 0.f && v[1] > 0.f && v[2] > 0.f && v[3] > 0.f && v[4] > 0.f && v[5] > 0.f;
}

// ================================================================================================
// one NUTS transition (Stan base_nuts::transition, iterative tree): same control flow as transition() of the resident
// kernel, with every vector in global memory.  On entry `qcur` = current point, `sm` = sqrt(inverse metric).
// On exit qcur holds the new sample.
// ================================================================================================
__device__ __noinline__ void s_transition(const SRunArgs& a, float* ws, float* rbuf, float* qcur, const float* sm, uint32_t chain_gid, uint32_t iter,
                                          float eps, const SEmit em, TransStats& st) {
  const ModelS& m = SMD();
  SCtl& ctl = SCTL();
  const int tid = threadIdx.x;
  const int n4 = nv4();
  float* G = s_slot(ws, SW_G);
  float* P0 = s_slot(ws, SW_PCUR);
  float ksq = s_draw_momentum(a, P0, chain_gid, iter, 1, 0);
  {
    SweepIO io{};
    io.qin = qcur; io.gout = G; io.em = em;
    s_eval(io, rbuf);
  }
  const double U0 = ctl.U;
  const double H0 = U0 + 0.5 * (double)ksq;
  // the two trajectory ends, stored mid-leapfrog: (q +- eps s p_half, p_half), and the tree summary
  float* endq[2][2] = {{s_slot(ws, SW_ENDB_Q0), s_slot(ws, SW_ENDB_Q1)}, {s_slot(ws, SW_ENDF_Q0), s_slot(ws, SW_ENDF_Q1)}};
  float* endp[2] = {s_slot(ws, SW_ENDB_P), s_slot(ws, SW_ENDF_P)};
  int endcur[2] = {0, 0};
  {
    const float hs = 0.5f * eps;
    float *tbb = s_slot(ws, SW_TOP_BB), *tff = s_slot(ws, SW_TOP_FF), *trho = s_slot(ws, SW_TOP_RHO), *ca = s_slot(ws, SW_CAND_A);
    for (int i0 = tid; i0 < n4; i0 += 2 * SNT) {
      float4 P2[2], s2[2], g2[2], q2[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) { const int i = i0 + u * SNT; P2[u] = ld4(P0, i); s2[u] = ld4(sm, i); g2[u] = ld4(G, i); q2[u] = ld4(qcur, i); }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int i = i0 + u * SNT;
        const float4 P = P2[u], q = q2[u];
        const float Pv[4] = {P.x, P.y, P.z, P.w}, sv[4] = {s2[u].x, s2[u].y, s2[u].z, s2[u].w}, gv[4] = {g2[u].x, g2[u].y, g2[u].z, g2[u].w},
                    qv[4] = {q.x, q.y, q.z, q.w};
        float pf[4], pb[4], qf[4], qb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          pf[j] = fmaf(-hs * sv[j], gv[j], Pv[j]);
          pb[j] = fmaf(hs * sv[j], gv[j], Pv[j]);
          qf[j] = fmaf(eps * sv[j], pf[j], qv[j]);
          qb[j] = fmaf(-eps * sv[j], pb[j], qv[j]);
        }
        st4(endp[1], i, make_float4(pf[0], pf[1], pf[2], pf[3]));
        st4(endp[0], i, make_float4(pb[0], pb[1], pb[2], pb[3]));
        st4(endq[1][0], i, make_float4(qf[0], qf[1], qf[2], qf[3]));
        st4(endq[0][0], i, make_float4(qb[0], qb[1], qb[2], qb[3]));
        st4(tbb, i, P); st4(tff, i, P); st4(trho, i, P); st4(ca, i, q);
      }
    }
  }
  int samp = SW_CAND_A, prop = SW_CAND_B;
  if (tid == 0) { ctl.U_samp = U0; ctl.H_samp = H0; ctl.U_prop = U0; ctl.H_prop = H0; ctl.sum_metro = 0.f; ctl.czn_tag = nullptr; }
  __syncthreads();
  float lsw = 0.f;
  int n_leap = 0, depth = 0;
  bool divergent = false;
  float* SR = s_slot(ws, SW_SRUN);

  while (depth < a.max_depth) {
    uint32_t rw[4];
    rng_words(a.seed, chain_gid, (uint32_t)depth, iter, 2, 0, rw);
    const int dir = (rw[0] >> 31) ? 1 : -1;
    const int di = dir > 0 ? 1 : 0;
    if (tid == 0) ctl.czn_tag = nullptr;   // the other end's column sums are gone
    __syncthreads();
    const float eps_s = dir > 0 ? eps : -eps;
    float lsw_sub = -CUDART_INF_F;
    bool ok = true, persist = true;
    const int nleaf = 1 << depth;
    auto first_slot = [&](int mleaf) -> float* {
      const int z = mleaf ? (__ffs(mleaf) - 1) : depth;
      return s_slot(ws, SW_FIRST + z - 1);
    };
    for (int n = 0; n < nleaf; ++n) {
      const int t = __ffs(~n) - 1;  // trailing ones of n = number of subtrees this leaf completes
      const bool last = (n == nleaf - 1);
      float* qi = endq[di][endcur[di]];
      float* qo = endq[di][endcur[di] ^ 1];
      // where this leaf's momentum goes: even leaf -> FIRST slot (it starts subtrees); a leaf closing a left half at level
      // t >= 1 -> that summary's e slot; else scratch
      float* Pd = (!last && t == 0) ? first_slot(n) : ((!last && t >= 1) ? s_slot(ws, SW_LEFT_E + t - 1) : P0);
      SweepIO io{};
      io.qin = qi; io.qout = qo; io.ph = endp[di]; io.sm = sm; io.Pdst = Pd; io.Lr = (t > 0) ? first_slot(n - 1) : nullptr; io.eps_s = eps_s;
      s_leaf(io, rbuf);
      SPROF_DECL;
      double h = ctl.U + 0.5 * (double)io.kk;
      if (!(h == h)) h = CUDART_INF;
      ++n_leap;
      const float dH = (float)(H0 - h);
      if (h - H0 > 1000.0) divergent = true;
      lsw_sub = logaddexp_f(lsw_sub, dH);
      if (tid == 0) ctl.sum_metro += (dH > 0.f) ? 1.0f : __expf(dH);
      endcur[di] ^= 1;   // the end of the trajectory is now (qo, ph)
      if (divergent) { ok = false; break; }
      {
        uint32_t sw[4];
        rng_words(a.seed, chain_gid, (uint32_t)n_leap, iter, 3, 0, sw);
        if (n == 0 || u01(sw[0]) < __expf(dH - lsw_sub)) {
          s_copy(s_slot(ws, prop), qi);
          if (tid == 0) { ctl.U_prop = ctl.U; ctl.H_prop = h; }
        }
      }
      SPROF(12);
      // U-turn checks for every subtree this leaf completes (level 0 came with the sweep)
      if (t > 0) ok = (io.c1a > 0.f) && (io.c1b > 0.f);
      if (t > 1 && ok) ok = s_merge(first_slot(n - 3), s_slot(ws, SW_LEFT_E), nullptr, first_slot(n - 1), Pd, first_slot(n - 1), SR);
      for (int k = 2; k < t && ok; ++k)
        ok = s_merge(first_slot(n - (2 << k) + 1), s_slot(ws, SW_LEFT_E + k - 1), s_slot(ws, SW_LEFT_R + k - 1), first_slot(n - (1 << k) + 1), Pd, SR, SR);
      if (!ok) break;
      SPROF(13);
      if (!last) {
        if (t >= 2) {   // stored left half at level t: e = P (already there), r = S + P
          s_add(s_slot(ws, SW_LEFT_R + t - 1), SR, Pd);
        }
      } else {
        // last leaf: merge the finished subtree with the existing trajectory (top level of base_nuts::transition)
        const float* F = s_slot(ws, dir > 0 ? SW_TOP_BB : SW_TOP_FF);
        const float* A = s_slot(ws, dir > 0 ? SW_TOP_FF : SW_TOP_BB);
        const float* Rb = (depth == 0) ? nullptr : first_slot(0);
        // S of the new subtree so far: t == 0 (depth 0): none; t == 1: the previous leaf; t >= 2: SR
        const float* Sin = (t == 0) ? nullptr : (t == 1 ? first_slot(n - 1) : SR);
        persist = s_merge(F, A, s_slot(ws, SW_TOP_RHO), Rb, Pd, Sin, SR);
        float* rho = s_slot(ws, SW_TOP_RHO);
        float* end = s_slot(ws, dir > 0 ? SW_TOP_FF : SW_TOP_BB);
        s_add(rho, SR, Pd);
        s_copy(end, Pd);
      }
      __syncthreads();
      SPROF(14);
#ifdef POTUS_PROF
      if (tid == 0) ctl.prof[31] += 1;
#endif
    }
    if (!ok) break;
    ++depth;
    if (lsw_sub > lsw || u01(rw[1]) < __expf(lsw_sub - lsw)) {
      const int tmp = samp; samp = prop; prop = tmp;
      if (tid == 0) { ctl.U_samp = ctl.U_prop; ctl.H_samp = ctl.H_prop; }
    }
    lsw = logaddexp_f(lsw, lsw_sub);
    if (!persist) break;
  }
  __syncthreads();
  s_copy(qcur, s_slot(ws, samp));
  if (tid == 0) ctl.czn_tag = nullptr;
  __syncthreads();
  st.lp = (float)(-ctl.U_samp);
  st.accept = ctl.sum_metro / (float)(n_leap > 0 ? n_leap : 1);
  st.eps = eps; st.depth = (float)depth; st.nleap = (float)n_leap; st.divergent = divergent ? 1.f : 0.f;
  st.energy = (float)ctl.H_samp;
  if (tid == 0) { ctl.cs.n_leapfrog += n_leap; ctl.cs.U = (float)ctl.U_samp; }
  __syncthreads();
}

// Stan base_hmc::init_stepsize on the point in qcur (left unchanged)
__device__ __noinline__ float s_find_stepsize(const SRunArgs& a, float* ws, float* rbuf, float* qcur, const float* sm, uint32_t chain_gid, uint32_t iter_tag,
                                              float eps) {
  SCtl& ctl = SCTL();
  const int tid = threadIdx.x, n4 = nv4();
  if (!(eps > 0.f) || eps > 1e7f) return eps;
  float *G = s_slot(ws, SW_G), *Pm = s_slot(ws, SW_TMPP), *Q1 = s_slot(ws, SW_TMPQ);
  int direction = 0;
  for (uint32_t attempt = 0; attempt < 200; ++attempt) {
    float k0 = s_draw_momentum(a, Pm, chain_gid, iter_tag, 5, attempt);
    SweepIO io{};
    io.qin = qcur; io.gout = G;
    s_eval(io, rbuf);
    const double H0 = ctl.U + 0.5 * (double)k0;
    const float hs = 0.5f * eps;
    for (int i = tid; i < n4; i += SNT) {   // p_half = p - eps/2 s g ; q1 = q + eps s p_half
      const float4 p = ld4(Pm, i), s = ld4(sm, i), g = ld4(G, i), q = ld4(qcur, i);
      const float4 ph = make_float4(fmaf(-hs * s.x, g.x, p.x), fmaf(-hs * s.y, g.y, p.y), fmaf(-hs * s.z, g.z, p.z), fmaf(-hs * s.w, g.w, p.w));
      st4(Pm, i, ph);
      st4(Q1, i, make_float4(fmaf(eps * s.x, ph.x, q.x), fmaf(eps * s.y, ph.y, q.y), fmaf(eps * s.z, ph.z, q.z), fmaf(eps * s.w, ph.w, q.w)));
    }
    __syncthreads();
    SweepIO io2{};
    io2.qin = Q1; io2.gout = G;
    s_eval(io2, rbuf);
    float k1[1] = {0.f};
    for (int i = tid; i < n4; i += SNT) {
      const float4 p = ld4(Pm, i), s = ld4(sm, i), g = ld4(G, i);
      const float a0 = fmaf(-hs * s.x, g.x, p.x), a1 = fmaf(-hs * s.y, g.y, p.y), a2 = fmaf(-hs * s.z, g.z, p.z), a3 = fmaf(-hs * s.w, g.w, p.w);
      k1[0] = fmaf(a0, a0, fmaf(a1, a1, fmaf(a2, a2, fmaf(a3, a3, k1[0]))));
    }
    s_block_sum(k1);
    double h = ctl.U + 0.5 * (double)k1[0];
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
  return eps;
}

// ================================================================================================
// kernels
// ================================================================================================
__device__ __forceinline__ void s_cta_setup(const ModelS& mg) {
  const int tid = threadIdx.x, w = tid >> 5;
  SCtl& ctl = SCTL();
  if (w == 0) { ptx::tmem_alloc(&ctl.tmem_base, 512); ptx::tmem_relinquish(); }
  if (tid == 0) {
    for (int i = 0; i < ST_NSTAGE; ++i) { ptx::mbar_init(&ctl.bar_full[i], 1); ptx::mbar_init(&ctl.bar_empty[i], 1); }
    ptx::mbar_init(&ctl.bar_done, 1);
    ptx::fence_mbar_init();
    ctl.prod_it = 0; ctl.cons_it = 0; ctl.gemm_cnt = 0; ctl.czn_tag = nullptr;
  }
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&mg);
    for (int i = tid; i < (int)(sizeof(ModelS) / 4); i += SNT) SMP(uint32_t, SS_MODEL)[i] = src[i];
  }
  for (int i = tid; i < 256; i += SNT) {
    sF(SS_PRIOR)[i] = mg.prior[i]; sF(SS_W)[i] = mg.w[i]; sF(SS_LW)[i] = mg.lw[i];
    sF(SS_PB)[i] = 0.f; sF(SS_BASE)[i] = 0.f; sF(SS_CZN)[i] = 0.f;
  }
  for (int i = tid; i < ST_MAXT; i += SNT) { sF(SS_E)[i] = 0.f; sF(SS_GE)[i] = 0.f; }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
}
__device__ __forceinline__ void s_cta_teardown() {
  ptx::tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) ptx::tmem_dealloc(SCTL().tmem_base, 512);
}

// test hook: log density + gradient for n positions (potus_logp_grad on shapes the resident kernel does not hold)
extern "C" __global__ void __launch_bounds__(SNT, 1) potus_stream_eval_kernel(const __grid_constant__ SEvalArgs a) {
  s_cta_setup(a.m);
  float* rbuf = a.rbuf + (size_t)blockIdx.x * a.m.rb_len;
  for (int i = blockIdx.x; i < a.n; i += gridDim.x) {
    SweepIO io{};
    io.qin = a.q_in + (size_t)i * a.m.VL;
    io.gout = a.g_out + (size_t)i * a.m.VL;
    s_eval(io, rbuf);
    if (threadIdx.x == 0) a.u_out[i] = SCTL().U;
    __syncthreads();
  }
  s_cta_teardown();
}

extern "C" __global__ void __launch_bounds__(SNT, 1) potus_stream_kernel(const __grid_constant__ SRunArgs a) {
  s_cta_setup(a.m);
  const ModelS& m = SMD();
  SCtl& ctl = SCTL();
  const int tid = threadIdx.x;
  const int VL = m.VL, n4 = VL >> 2;
  float* ws = a.workspace + (size_t)blockIdx.x * SW_NSLOT * VL;
  float* rbuf = a.rbuf + (size_t)blockIdx.x * m.rb_len;
  const int n_iter_total = a.iter_warmup + a.iter_sampling;
  const SEmit none{nullptr, nullptr};
#ifdef POTUS_PROF
  if (tid < 32) ctl.prof[tid] = 0ull;
#endif

  for (;;) {
    __syncthreads();
    if (tid == 0) ctl.chain = atomicAdd(a.queue, 1);
    __syncthreads();
    const int chain = ctl.chain;
    if (chain >= a.n_chains) break;
    const uint32_t gid = (uint32_t)(a.chain_id_offset + chain);
    float* qg = a.q + (size_t)chain * VL;       // the chain's current point lives here throughout
    float* sg = a.sqrt_m + (size_t)chain * VL;
    float* wmean = a.wf_mean + (size_t)chain * VL;
    float* wm2 = a.wf_m2 + (size_t)chain * VL;
    const int32_t* map = m.map_i2s;
    if (tid == 0) { ctl.cs = a.cs[chain]; ctl.czn_tag = nullptr; }
    __syncthreads();

    if (a.do_init) {
      for (int i = tid; i < VL; i += SNT) { sg[i] = (__ldg(map + i) >= 0) ? 1.0f : 0.f; wmean[i] = 0.f; wm2[i] = 0.f; }
      bool good = false;
      float* G = s_slot(ws, SW_G);
      for (uint32_t attempt = 0; attempt < 100 && !good; ++attempt) {
        for (int i = tid; i < VL; i += SNT) {
          const int si = __ldg(map + i);
          float v = 0.f;
          if (si >= 0) { uint32_t rw[4]; rng_words(a.seed, gid, (uint32_t)si, 0, 0, attempt, rw); v = a.init_radius * (2.0f * u01(rw[0]) - 1.0f); }
          qg[i] = v;
        }
        __syncthreads();
        SweepIO io{};
        io.qin = qg; io.gout = G;
        s_eval(io, rbuf);
        int bad = 0;
        for (int i = tid; i < VL; i += SNT) bad |= !isfinite(G[i]);
        if (!isfinite(ctl.U)) bad = 1;
        good = __syncthreads_or(bad) == 0;
      }
      if (tid == 0) {
        ChainState& cs = ctl.cs;
        cs.status = good ? 0 : -1;
        cs.eps = 1.0f; cs.da_counter = 0; cs.da_sbar = 0; cs.da_xbar = 0;
        cs.w_counter = 0; cs.w_size = a.w_base_window; cs.w_nsamp = 0;
        cs.w_next = a.w_base_window > 0 ? a.w_init_buffer + a.w_base_window - 1 : -1;
        cs.iter = 0; cs.n_leapfrog = 0;
      }
      __syncthreads();
      const float e0 = s_find_stepsize(a, ws, rbuf, qg, sg, gid, 0xFFFFFFFFu, 1.0f);
      if (tid == 0) { ctl.cs.eps = e0; ctl.cs.da_mu = log(10.0 * (double)e0); }
      __syncthreads();
    }

    for (int it = a.iter_begin; it < a.iter_end; ++it) {
      const float eps = ctl.cs.eps;
      SEmit em = none;
      const int kprev = it - 1 - a.iter_warmup;
      if (kprev >= 0) {
        em.monitor = a.monitor + ((size_t)chain * a.iter_sampling + kprev) * (m.S + 1);
        if (a.keep_per_chain > 0 && (kprev % a.keep_every) == 0) {
          const int slot = kprev / a.keep_every;
          if (slot < a.keep_per_chain) em.draw = a.draws + ((size_t)chain * a.keep_per_chain + slot) * a.draw_len;
        }
      }
      TransStats st;
      s_transition(a, ws, rbuf, qg, sg, gid, (uint32_t)it, eps, em, st);
      if (tid == 0) {
        float* sp = a.sampler_params + ((size_t)chain * n_iter_total + it) * 8;
        sp[0] = st.lp; sp[1] = st.accept; sp[2] = st.eps; sp[3] = st.depth; sp[4] = st.nleap; sp[5] = st.divergent; sp[6] = st.energy; sp[7] = 0.f;
      }
      if (it < a.iter_warmup) {
        if (tid == 0) {   // Stan stepsize_adaptation::learn_stepsize (dual averaging)
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
        const int wc = ctl.cs.w_counter;
        const bool in_window = wc >= a.w_init_buffer && wc < a.iter_warmup - a.w_term_buffer && wc != a.iter_warmup;
        const bool end_window = wc == ctl.cs.w_next && wc != a.iter_warmup;
        int nsamp = ctl.cs.w_nsamp;
        if (in_window) {   // Welford update (padding: q = 0 keeps mean = m2 = 0)
          ++nsamp;
          const float inv = 1.0f / (float)nsamp;
          for (int i = tid; i < n4; i += SNT) {
            const float4 q = ld4(qg, i), mu = ld4(wmean, i), m2 = ld4(wm2, i);
            const float qv[4] = {q.x, q.y, q.z, q.w}, mv[4] = {mu.x, mu.y, mu.z, mu.w}, sv[4] = {m2.x, m2.y, m2.z, m2.w};
            float mo[4], so[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float dd = qv[j] - mv[j]; mo[j] = fmaf(dd, inv, mv[j]); so[j] = fmaf(qv[j] - mo[j], dd, sv[j]); }
            st4(wmean, i, make_float4(mo[0], mo[1], mo[2], mo[3]));
            st4(wm2, i, make_float4(so[0], so[1], so[2], so[3]));
          }
        }
        __syncthreads();
        if (end_window) {
          const float n = (float)nsamp;
          for (int i = tid; i < VL; i += SNT) {
            float v = 0.f;
            if (__ldg(map + i) >= 0) {
              const float var = wm2[i] / (n - 1.0f);
              v = sqrtf((n / (n + 5.0f)) * var + 1e-3f * (5.0f / (n + 5.0f)));
            }
            sg[i] = v; wmean[i] = 0.f; wm2[i] = 0.f;
          }
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
          __syncthreads();
          const float e1 = s_find_stepsize(a, ws, rbuf, qg, sg, gid, (uint32_t)it, ctl.cs.eps);
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
    // the last iteration's draw needs one more evaluation at the final point
    if (a.iter_end == n_iter_total && a.iter_end > a.iter_begin) {
      SEmit em = none;
      const int kprev = n_iter_total - 1 - a.iter_warmup;
      if (kprev >= 0) {
        em.monitor = a.monitor + ((size_t)chain * a.iter_sampling + kprev) * (m.S + 1);
        if (a.keep_per_chain > 0 && (kprev % a.keep_every) == 0) {
          const int slot = kprev / a.keep_every;
          if (slot < a.keep_per_chain) em.draw = a.draws + ((size_t)chain * a.keep_per_chain + slot) * a.draw_len;
        }
        SweepIO io{};
        io.qin = qg; io.gout = s_slot(ws, SW_G); io.em = em;
        s_eval(io, rbuf);
      }
    }
    if (tid == 0) { ctl.cs.iter = a.iter_end; a.cs[chain] = ctl.cs; }
    __syncthreads();
  }
#ifdef POTUS_PROF
  if (a.prof != nullptr && tid < 32) atomicAdd(a.prof + tid, ctl.prof[tid]);
#endif
  s_cta_teardown();
}

}  // namespace potus
