/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY (also the reported CPU baseline of bench.py).
 * Nothing here is on the product path.  fp64 restatement of
 *   /root/reference/scripts/model/poll_model_2020.stan:42-132 (+ _no_mode_adjustment variant)
 * and of the sampler Stan/CmdStan 2.24.1 runs on it (NOT IN THE REFERENCE TREE -- third-party,
 * pinned only by version strings inside scripts/model/poll_model_2020; restated from its
 * published algorithm: stan/mcmc/hmc/nuts/base_nuts.hpp, hamiltonians/diag_e_metric.hpp,
 * integrators/expl_leapfrog.hpp, stepsize_adaptation.hpp, windowed_adaptation.hpp,
 * var_adaptation.hpp, base_hmc.hpp::init_stepsize).
 * PARITY: "unpinned" below the end-to-end posterior tables (see oracle/potus_oracle.py header).
 *
 * Two gradient forms:  orc_logp_grad(..., literal=1) walks the T-1 matvecs of :86 exactly as the
 * Stan program does (the CPU cost model of the reference); literal=0 is the collapsed
 * scan + one-GEMM form the GPU kernel uses.  They agree to ~1e-12 (tests/test_oracle.py).
 * Two NUTS tree builders with identical semantics: mode 0 = Stan's recursion; mode 1 = the
 * iterative leaf-by-leaf form with streaming (reservoir) multinomial selection that the CUDA
 * kernel implements.  Both use the counter-based Philox4x32-10 streams of the CUDA kernel.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "../include/potus_b200.h"

#define ORC_API __attribute__((visibility("default")))

typedef struct OrcModel {
  int S, T, P, M, Pop, Nn, Ns, full, D;
  int o_zT, o_Z, o_c, o_m, o_pop, o_umu, o_urho, o_ze, o_xn, o_xs, o_zb;
  int *st, *ds, *dn, *ps, *pn, *ms, *mn, *os, *on;
  double *ys, *ns, *yn, *nn, *us, *un, *prior, *w, *L0; /* L0 row-major S*S lower */
  double a_b, a_T, a_w, sig_c, sig_m, sig_pop, sig_n, sig_s, sig_e;
} OrcModel;

static int* dup_idx(const int32_t* a, int n) {
  int* r = (int*)malloc(sizeof(int) * (n > 0 ? n : 1));
  for (int i = 0; i < n; ++i) r[i] = a ? a[i] - 1 : 0;
  return r;
}
static double* dup_i2d(const int32_t* a, int n) {
  double* r = (double*)malloc(sizeof(double) * (n > 0 ? n : 1));
  for (int i = 0; i < n; ++i) r[i] = (double)a[i];
  return r;
}
static double* dup_d(const double* a, int n, double fill) {
  double* r = (double*)malloc(sizeof(double) * (n > 0 ? n : 1));
  for (int i = 0; i < n; ++i) r[i] = a ? a[i] : fill;
  return r;
}

ORC_API OrcModel* orc_model_create(const PotusData* d) {
  OrcModel* m = (OrcModel*)calloc(1, sizeof(OrcModel));
  m->S = d->S; m->T = d->T; m->P = d->P; m->M = d->M; m->Pop = d->Pop;
  m->Nn = d->N_national_polls; m->Ns = d->N_state_polls;
  m->full = d->poll_mode_state != NULL;
  int o = 0;
  m->o_zT = o; o += m->S;
  m->o_Z = o; o += m->S * m->T;
  m->o_c = o; o += m->P;
  if (m->full) {
    m->o_m = o; o += m->M;
    m->o_pop = o; o += m->Pop;
    m->o_umu = o; o += 1;
    m->o_urho = o; o += 1;
    m->o_ze = o; o += m->T;
  }
  m->o_xn = o; o += m->Nn;
  m->o_xs = o; o += m->Ns;
  m->o_zb = o; o += m->S;
  m->D = o;
  m->st = dup_idx(d->state, m->Ns); m->ds = dup_idx(d->day_state, m->Ns); m->dn = dup_idx(d->day_national, m->Nn);
  m->ps = dup_idx(d->poll_state, m->Ns); m->pn = dup_idx(d->poll_national, m->Nn);
  m->ms = dup_idx(d->poll_mode_state, m->Ns); m->mn = dup_idx(d->poll_mode_national, m->Nn);
  m->os = dup_idx(d->poll_pop_state, m->Ns); m->on = dup_idx(d->poll_pop_national, m->Nn);
  m->ys = dup_i2d(d->n_democrat_state, m->Ns); m->ns = dup_i2d(d->n_two_share_state, m->Ns);
  m->yn = dup_i2d(d->n_democrat_national, m->Nn); m->nn = dup_i2d(d->n_two_share_national, m->Nn);
  m->us = dup_d(d->unadjusted_state, m->Ns, 0.0); m->un = dup_d(d->unadjusted_national, m->Nn, 0.0);
  m->prior = dup_d(d->mu_b_prior, m->S, 0.0); m->w = dup_d(d->state_weights, m->S, 0.0);
  /* transformed data, poll_model_2020.stan:43-54: nat_sd, Cholesky of Sigma0 (the three factors
   * are scalar multiples of it) */
  int S = m->S;
  double nat = 0;
  for (int i = 0; i < S; ++i)
    for (int j = 0; j < S; ++j) nat += m->w[i] * d->state_covariance_0[i + S * j] * m->w[j];
  nat = sqrt(nat);
  m->L0 = (double*)calloc((size_t)S * S, sizeof(double));
  for (int j = 0; j < S; ++j) {
    double s = d->state_covariance_0[j + S * j];
    for (int k = 0; k < j; ++k) s -= m->L0[j * S + k] * m->L0[j * S + k];
    if (!(s > 0)) { free(m); return NULL; }
    double ljj = sqrt(s);
    m->L0[j * S + j] = ljj;
    for (int i = j + 1; i < S; ++i) {
      double v = d->state_covariance_0[i + S * j];
      for (int k = 0; k < j; ++k) v -= m->L0[i * S + k] * m->L0[j * S + k];
      m->L0[i * S + j] = v / ljj;
    }
  }
  m->a_b = d->polling_bias_scale / nat; m->a_T = d->mu_b_T_scale / nat; m->a_w = d->random_walk_scale / nat;
  m->sig_c = d->sigma_c; m->sig_m = d->sigma_m; m->sig_pop = d->sigma_pop;
  m->sig_n = d->sigma_measure_noise_national; m->sig_s = d->sigma_measure_noise_state; m->sig_e = d->sigma_e_bias;
  return m;
}
ORC_API void orc_model_destroy(OrcModel* m) {
  if (!m) return;
  free(m->st); free(m->ds); free(m->dn); free(m->ps); free(m->pn); free(m->ms); free(m->mn); free(m->os); free(m->on);
  free(m->ys); free(m->ns); free(m->yn); free(m->nn); free(m->us); free(m->un); free(m->prior); free(m->w); free(m->L0);
  free(m);
}
ORC_API int orc_num_params(const OrcModel* m) { return m->D; }

static inline double softplus(double x) { return (x > 0 ? x : 0) + log1p(exp(-fabs(x))); }
static inline double inv_logit(double x) { return x >= 0 ? 1.0 / (1.0 + exp(-x)) : exp(x) / (1.0 + exp(x)); }

typedef struct { double *mu, *G, *W, *e, *ge, *gc, *pb, *gpb, *tmp, *rs, *rn, *nat; } Work;
static Work* work_new(const OrcModel* m) {
  Work* w = (Work*)calloc(1, sizeof(Work));
  size_t ST = (size_t)m->S * m->T;
  w->mu = (double*)malloc(sizeof(double) * ST); w->G = (double*)malloc(sizeof(double) * ST);
  w->W = (double*)malloc(sizeof(double) * ST);
  w->e = (double*)calloc(m->T, sizeof(double)); w->ge = (double*)calloc(m->T, sizeof(double));
  w->gc = (double*)calloc(m->P + 16, sizeof(double)); w->pb = (double*)calloc(m->S, sizeof(double));
  w->gpb = (double*)calloc(m->S, sizeof(double)); w->tmp = (double*)calloc(2 * m->S + m->T, sizeof(double));
  w->rs = (double*)calloc(m->Ns + 1, sizeof(double)); w->rn = (double*)calloc(m->Nn + 1, sizeof(double));
  w->nat = (double*)calloc(m->T, sizeof(double));
  return w;
}
static void work_free(Work* w) {
  free(w->mu); free(w->G); free(w->W); free(w->e); free(w->ge); free(w->gc); free(w->pb); free(w->gpb); free(w->tmp);
  free(w->rs); free(w->rn); free(w->nat); free(w);
}

/* lp and gradient at theta.  mu is stored day-major: mu[t*S + s]. */
static double logp_grad_w(const OrcModel* m, const double* th, double* g, int literal, Work* wk, double* mu_out) {
  const int S = m->S, T = m->T;
  const double* zT = th + m->o_zT; const double* Z = th + m->o_Z; const double* L0 = m->L0;
  double* mu = wk->mu; double* G = wk->G;
  double lp = 0;
  for (int i = 0; i < m->D; ++i) { lp -= 0.5 * th[i] * th[i]; g[i] = -th[i]; }
  /* ---- mu_b, poll_model_2020.stan:85-86 */
  if (literal) {
    for (int s = 0; s < S; ++s) {
      double a = 0;
      for (int k = 0; k <= s; ++k) a += L0[s * S + k] * zT[k];
      mu[(T - 1) * S + s] = m->a_T * a + m->prior[s];
    }
    for (int t = T - 2; t >= 0; --t)
      for (int s = 0; s < S; ++s) {
        double a = 0;
        for (int k = 0; k <= s; ++k) a += L0[s * S + k] * Z[t * S + k];
        mu[t * S + s] = m->a_w * a + mu[(t + 1) * S + s];
      }
  } else {
    double* W = wk->W; /* W[t*S+k] = a_T zT[k] + a_w sum_{u=t}^{T-2} Z[u*S+k] */
    for (int k = 0; k < S; ++k) W[(T - 1) * S + k] = m->a_T * zT[k];
    for (int t = T - 2; t >= 0; --t)
      for (int k = 0; k < S; ++k) W[t * S + k] = W[(t + 1) * S + k] + m->a_w * Z[t * S + k];
    for (int t = 0; t < T; ++t)
      for (int s = 0; s < S; ++s) {
        double a = 0;
        for (int k = 0; k <= s; ++k) a += L0[s * S + k] * W[t * S + k];
        mu[t * S + s] = a + m->prior[s];
      }
  }
  /* ---- :77,79,87 */
  double* pb = wk->pb; double nat_pb = 0;
  for (int s = 0; s < S; ++s) {
    double a = 0;
    for (int k = 0; k <= s; ++k) a += L0[s * S + k] * th[m->o_zb + k];
    pb[s] = m->a_b * a; nat_pb += pb[s] * m->w[s];
  }
  double* nat = wk->nat;
  for (int t = 0; t < T; ++t) { double a = 0; for (int s = 0; s < S; ++s) a += m->w[s] * mu[t * S + s]; nat[t] = a; }
  /* ---- :88-93 */
  double rho = 0, mu_e = 0, sig_rho = 0; double* e = wk->e;
  if (m->full) {
    double u_mu = th[m->o_umu], u_rho = th[m->o_urho];
    mu_e = 0.02 * u_mu; rho = inv_logit(u_rho); sig_rho = sqrt(1 - rho * rho) * m->sig_e;
    lp += 0.5 * u_rho * u_rho; /* rho's unconstrained value carries no N(0,1) term */
    lp += -0.5 * ((rho - 0.7) / 0.1) * ((rho - 0.7) / 0.1) + log(rho) + log1p(-rho);
    const double* ze = th + m->o_ze;
    e[0] = ze[0] * m->sig_e;
    for (int t = 1; t < T; ++t) e[t] = mu_e + rho * (e[t - 1] - mu_e) + ze[t] * sig_rho;
  }
  /* ---- :95-112 linear predictor, :130-131 likelihood, residuals */
  memset(G, 0, sizeof(double) * (size_t)S * T);
  double* ge = wk->ge; memset(ge, 0, sizeof(double) * T);
  double* gc = wk->gc; memset(gc, 0, sizeof(double) * (m->P + 16));
  double* gm = gc + m->P; double* gp = gc + m->P + 8;
  double* gpb = wk->gpb; memset(gpb, 0, sizeof(double) * S);
  double* rnday = wk->tmp + 2 * S; memset(rnday, 0, sizeof(double) * T);
  double sum_rn = 0;
  for (int i = 0; i < m->Ns; ++i) {
    int s = m->st[i], d = m->ds[i];
    double eta = mu[d * S + s] + m->sig_c * th[m->o_c + m->ps[i]] + m->sig_s * th[m->o_xs + i] + pb[s];
    if (m->full) eta += m->sig_m * th[m->o_m + m->ms[i]] + m->sig_pop * th[m->o_pop + m->os[i]] + m->us[i] * e[d];
    lp += m->ys[i] * eta - m->ns[i] * softplus(eta);
    double r = m->ys[i] - m->ns[i] * inv_logit(eta);
    g[m->o_xs + i] += m->sig_s * r;
    G[d * S + s] += r; gpb[s] += r; gc[m->ps[i]] += r;
    if (m->full) { gm[m->ms[i]] += r; gp[m->os[i]] += r; ge[d] += m->us[i] * r; }
  }
  for (int j = 0; j < m->Nn; ++j) {
    int d = m->dn[j];
    double eta = nat[d] + m->sig_c * th[m->o_c + m->pn[j]] + m->sig_n * th[m->o_xn + j] + nat_pb;
    if (m->full) eta += m->sig_m * th[m->o_m + m->mn[j]] + m->sig_pop * th[m->o_pop + m->on[j]] + m->un[j] * e[d];
    lp += m->yn[j] * eta - m->nn[j] * softplus(eta);
    double r = m->yn[j] - m->nn[j] * inv_logit(eta);
    g[m->o_xn + j] += m->sig_n * r;
    rnday[d] += r; sum_rn += r; gc[m->pn[j]] += r;
    if (m->full) { gm[m->mn[j]] += r; gp[m->on[j]] += r; ge[d] += m->un[j] * r; }
  }
  for (int t = 0; t < T; ++t)
    if (rnday[t] != 0) for (int s = 0; s < S; ++s) G[t * S + s] += m->w[s] * rnday[t];
  for (int s = 0; s < S; ++s) gpb[s] += m->w[s] * sum_rn;
  /* ---- adjoint of mu_b: H = L0^T G, forward cumsum over days */
  double* acc = wk->tmp; /* running sum_{t<=u} H[:,t] */
  memset(acc, 0, sizeof(double) * S);
  double* h = wk->tmp + S;
  for (int t = 0; t < T; ++t) {
    for (int k = 0; k < S; ++k) h[k] = 0;
    for (int s = 0; s < S; ++s) {
      double gs = G[t * S + s];
      if (gs != 0) for (int k = 0; k <= s; ++k) h[k] += L0[s * S + k] * gs;
    }
    for (int k = 0; k < S; ++k) acc[k] += h[k];
    if (t < T - 1) for (int k = 0; k < S; ++k) g[m->o_Z + t * S + k] += m->a_w * acc[k];
  }
  for (int k = 0; k < S; ++k) g[m->o_zT + k] += m->a_T * acc[k];
  for (int k = 0; k < S; ++k) {
    double a = 0;
    for (int s = k; s < S; ++s) a += L0[s * S + k] * gpb[s];
    g[m->o_zb + k] += m->a_b * a;
  }
  for (int p = 0; p < m->P; ++p) g[m->o_c + p] += m->sig_c * gc[p];
  if (m->full) {
    for (int k = 0; k < m->M; ++k) g[m->o_m + k] += m->sig_m * gm[k];
    for (int k = 0; k < m->Pop; ++k) g[m->o_pop + k] += m->sig_pop * gp[k];
    /* AR(1) adjoint: ebar[t] = ge[t] + rho ebar[t+1] */
    const double* ze = th + m->o_ze;
    double eb = 0, s_mu = 0, s_rho = 0;
    for (int t = T - 1; t >= 1; --t) {
      eb = ge[t] + rho * eb;
      g[m->o_ze + t] += sig_rho * eb;
      s_mu += eb;
      s_rho += eb * ((e[t - 1] - mu_e) - ze[t] * m->sig_e * rho / sqrt(1 - rho * rho));
    }
    eb = ge[0] + rho * eb;
    g[m->o_ze + 0] += m->sig_e * eb;
    g[m->o_umu] += 0.02 * (1 - rho) * s_mu;
    double d_rho = s_rho - (rho - 0.7) / 0.01;
    g[m->o_urho] = rho * (1 - rho) * d_rho + (1 - 2 * rho);
  }
  if (mu_out) memcpy(mu_out, mu, sizeof(double) * (size_t)S * T);
  return lp;
}

ORC_API double orc_logp_grad(const OrcModel* m, const double* theta, double* grad, int literal) {
  Work* w = work_new(m);
  double lp = logp_grad_w(m, theta, grad, literal, w, NULL);
  work_free(w);
  return lp;
}
/* transformed parameters of one draw: mu_b[S*T] (Stan column-major s + S*t), mu_c[P], mu_m, mu_pop,
 * e_bias[T], polling_bias[S] */
ORC_API void orc_constrain(const OrcModel* m, const double* th, double* mu_b, double* mu_c, double* mu_m, double* mu_pop,
                           double* e_bias, double* polling_bias) {
  Work* w = work_new(m);
  double* g = (double*)malloc(sizeof(double) * m->D);
  logp_grad_w(m, th, g, 0, w, mu_b); /* day-major t*S+s == column-major s + S*t */
  for (int p = 0; p < m->P; ++p) mu_c[p] = m->sig_c * th[m->o_c + p];
  memcpy(polling_bias, w->pb, sizeof(double) * m->S);
  if (m->full) {
    for (int k = 0; k < m->M; ++k) mu_m[k] = m->sig_m * th[m->o_m + k];
    for (int k = 0; k < m->Pop; ++k) mu_pop[k] = m->sig_pop * th[m->o_pop + k];
    memcpy(e_bias, w->e, sizeof(double) * m->T);
  }
  free(g); work_free(w);
}

/* ------------------------------------------------------------------------------------------ RNG */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
/* streams: 0 init | 1 momentum | 2 per-depth scalars (w0 direction, w1 top-level accept) |
 * 3 per-leaf selection | 4 recursive-mode merge selection | 5 init_stepsize momentum */
ORC_API void orc_rng_words(uint64_t seed, uint32_t chain, uint32_t idx, uint32_t iter, uint32_t stream, uint32_t sub,
                           uint32_t out[4]) {
  out[0] = idx; out[1] = iter; out[2] = stream | (sub << 8); out[3] = chain;
  philox4x32_10(out, (uint32_t)seed, (uint32_t)(seed >> 32));
}
/* 23-bit uniforms: ((w>>9)+0.5)*2^-23 is exact in fp32 too, so the CUDA kernel draws the same numbers */
static inline double u01(uint32_t w) { return ((double)(w >> 9) + 0.5) * (1.0 / 8388608.0); }
static inline double rng_normal(uint64_t seed, uint32_t chain, uint32_t idx, uint32_t iter, uint32_t stream, uint32_t sub) {
  uint32_t w[4];
  orc_rng_words(seed, chain, idx, iter, stream, sub, w);
  return sqrt(-2.0 * log(u01(w[0]))) * cos(6.283185307179586476925 * u01(w[1]));
}
static inline double rng_unif(uint64_t seed, uint32_t chain, uint32_t idx, uint32_t iter, uint32_t stream, int word) {
  uint32_t w[4];
  orc_rng_words(seed, chain, idx, iter, stream, 0, w);
  return u01(w[word]);
}

/* ------------------------------------------------------------------------------------------ NUTS */
typedef struct {
  const OrcModel* m; Work* wk; int D; uint64_t seed; uint32_t chain; int literal; int tree_mode; int max_depth;
  double* inv_metric; /* diag of M^-1 */
  double eps;
  /* phase-space point z */
  double *q, *p, *g; double V;
  /* transition state */
  double H0; int n_leapfrog; double sum_metro; int divergent; uint32_t iter; uint32_t merge_ctr;
  int64_t n_grad;
  double* arena; size_t arena_top, arena_cap; /* bump allocator for the tree recursion (no malloc in the hot path) */
} Chain;

static void z_grad(Chain* c) { c->V = -logp_grad_w(c->m, c->q, c->g, c->literal, c->wk, NULL); for (int i = 0; i < c->D; ++i) c->g[i] = -c->g[i]; c->n_grad++; }
static double kinetic(const Chain* c, const double* p) { double k = 0; for (int i = 0; i < c->D; ++i) k += c->inv_metric[i] * p[i] * p[i]; return 0.5 * k; }
static void leapfrog(Chain* c, double eps) {
  for (int i = 0; i < c->D; ++i) c->p[i] -= 0.5 * eps * c->g[i];
  for (int i = 0; i < c->D; ++i) c->q[i] += eps * c->inv_metric[i] * c->p[i];
  z_grad(c);
  for (int i = 0; i < c->D; ++i) c->p[i] -= 0.5 * eps * c->g[i];
}
static double logaddexp(double a, double b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  return a > b ? a + log1p(exp(b - a)) : b + log1p(exp(a - b));
}
/* crit(p_sharp_minus, p_sharp_plus, rho) with p_sharp = M^-1 p given through p */
static int crit(const Chain* c, const double* p_a, const double* p_b, const double* rho) {
  double da = 0, db = 0;
  for (int i = 0; i < c->D; ++i) { da += c->inv_metric[i] * p_a[i] * rho[i]; db += c->inv_metric[i] * p_b[i] * rho[i]; }
  return db > 0 && da > 0;
}
static double* vnew(int D) { return (double*)malloc(sizeof(double) * D); }
static double* arena_get(Chain* c, int D, int zero) {
  if (c->arena_top + (size_t)D > c->arena_cap) { fprintf(stderr, "oracle: arena exhausted\n"); abort(); }
  double* r = c->arena + c->arena_top; c->arena_top += (size_t)D;
  if (zero) memset(r, 0, sizeof(double) * D);
  return r;
}

/* one leaf: leapfrog + energy bookkeeping (base_nuts.hpp build_tree depth==0) */
static int leaf(Chain* c, double sign, double* lsw, double* h_out) {
  leapfrog(c, sign * c->eps);
  c->n_leapfrog++;
  double h = c->V + kinetic(c, c->p);
  if (isnan(h)) h = INFINITY;
  if (h - c->H0 > 1000.0) c->divergent = 1;
  *lsw = logaddexp(*lsw, c->H0 - h);
  c->sum_metro += (c->H0 - h > 0) ? 1.0 : exp(c->H0 - h);
  *h_out = h;
  return !c->divergent;
}

/* ---- mode 0: Stan's recursion */
static int build_rec(Chain* c, int depth, double sign, double* z_prop, double* ps_beg_p, double* ps_end_p, double* rho,
                     double* lsw) {
  const int D = c->D;
  if (depth == 0) {
    double h;
    int ok = leaf(c, sign, lsw, &h);
    memcpy(z_prop, c->q, sizeof(double) * D);
    for (int i = 0; i < D; ++i) rho[i] += c->p[i];
    memcpy(ps_beg_p, c->p, sizeof(double) * D);
    memcpy(ps_end_p, c->p, sizeof(double) * D);
    return ok;
  }
  double lsw_i = -INFINITY, lsw_f = -INFINITY;
  const size_t mark = c->arena_top;
  double *p_ie = arena_get(c, D, 0), *rho_i = arena_get(c, D, 1);
  double *p_fb = arena_get(c, D, 0), *rho_f = arena_get(c, D, 1), *z_prop_f = arena_get(c, D, 0), *ext = arena_get(c, D, 0);
  int ok = build_rec(c, depth - 1, sign, z_prop, ps_beg_p, p_ie, rho_i, &lsw_i);
  if (ok) ok = build_rec(c, depth - 1, sign, z_prop_f, p_fb, ps_end_p, rho_f, &lsw_f);
  if (ok) {
    double lsw_sub = logaddexp(lsw_i, lsw_f);
    *lsw = logaddexp(*lsw, lsw_sub);
    int take;
    if (lsw_f > lsw_sub) take = 1;
    else take = rng_unif(c->seed, c->chain, c->merge_ctr, c->iter, 4, 0) < exp(lsw_f - lsw_sub);
    c->merge_ctr++;
    if (take) memcpy(z_prop, z_prop_f, sizeof(double) * D);
    for (int i = 0; i < D; ++i) ext[i] = rho_i[i] + rho_f[i];
    for (int i = 0; i < D; ++i) rho[i] += ext[i];
    ok = crit(c, ps_beg_p, ps_end_p, ext);
    for (int i = 0; i < D; ++i) ext[i] = rho_i[i] + p_fb[i];
    ok = ok && crit(c, ps_beg_p, p_fb, ext);
    for (int i = 0; i < D; ++i) ext[i] = rho_f[i] + p_ie[i];
    ok = ok && crit(c, p_ie, ps_end_p, ext);
  }
  c->arena_top = mark;
  return ok;
}

/* ---- mode 1: iterative builder (what the CUDA kernel does).  Left[k] = {b,e,r} of the completed
 * left half at level k; the right half is implicit: R.e = p, R.r = p + sum_{j<k} Left[j].r,
 * R.b = Left[k-1].b (or p).  Multinomial selection inside the subtree by reservoir sampling:
 * leaf n replaces the candidate with probability w_n / sum_{j<=n} w_j -- the same multinomial
 * law as Stan's pairwise merges.  Returns first/last/rho of the new subtree. */
typedef struct { double *b, *e, *r; } Triple;
static int build_iter(Chain* c, int depth, double sign, double* z_prop, double* first, double* last, double* rho_new,
                      double* lsw_sub, Triple* left) {
  const int D = c->D;
  const long nleaf = 1L << depth;
  double *S = vnew(D), *ext = vnew(D);
  int ok = 1;
  for (long n = 0; n < nleaf && ok; ++n) {
    double h;
    ok = leaf(c, sign, lsw_sub, &h);
    if (!ok) break;
    double u = rng_unif(c->seed, c->chain, (uint32_t)c->n_leapfrog, c->iter, 3, 0);
    if (n == 0 || u < exp((c->H0 - h) - *lsw_sub)) memcpy(z_prop, c->q, sizeof(double) * D);
    int t = 0;
    while ((n >> t) & 1) ++t;
    memset(S, 0, sizeof(double) * D);
    for (int k = 0; k < t && ok; ++k) {
      const double* Rb = (k == 0) ? c->p : left[k - 1].b;
      for (int i = 0; i < D; ++i) ext[i] = left[k].r[i] + S[i] + c->p[i];
      ok = crit(c, left[k].b, c->p, ext);
      for (int i = 0; i < D; ++i) ext[i] = left[k].r[i] + Rb[i];
      ok = ok && crit(c, left[k].b, Rb, ext);
      for (int i = 0; i < D; ++i) ext[i] = S[i] + c->p[i] + left[k].e[i];
      ok = ok && crit(c, left[k].e, c->p, ext);
      for (int i = 0; i < D; ++i) S[i] += left[k].r[i];
    }
    if (!ok) break;
    if (n < nleaf - 1) {
      memcpy(left[t].b, t == 0 ? c->p : left[t - 1].b, sizeof(double) * D);
      memcpy(left[t].e, c->p, sizeof(double) * D);
      for (int i = 0; i < D; ++i) left[t].r[i] = c->p[i] + S[i];
    } else {
      memcpy(first, depth == 0 ? c->p : left[depth - 1].b, sizeof(double) * D);
      memcpy(last, c->p, sizeof(double) * D);
      for (int i = 0; i < D; ++i) rho_new[i] = c->p[i] + S[i];
    }
  }
  free(S); free(ext);
  return ok;
}

typedef struct { double lp, accept_stat, stepsize, treedepth, n_leapfrog, divergent, energy; } SampStats;

/* base_nuts::transition.  On entry c->q is the current point; on exit c->q is the new sample. */
static void transition(Chain* c, SampStats* st) {
  const int D = c->D;
  double *qf = vnew(D), *pf = vnew(D), *gf = vnew(D), *qb = vnew(D), *pb = vnew(D), *gb = vnew(D);
  double *z_sample = vnew(D), *z_prop = vnew(D), *rho = vnew(D), *p_bb = vnew(D), *p_ff = vnew(D);
  double *first = vnew(D), *last = vnew(D), *rho_new = vnew(D), *ext = vnew(D);
  Triple left[16];
  for (int k = 0; k < c->max_depth; ++k) { left[k].b = vnew(D); left[k].e = vnew(D); left[k].r = vnew(D); }
  for (int i = 0; i < D; ++i) c->p[i] = rng_normal(c->seed, c->chain, (uint32_t)i, c->iter, 1, 0) / sqrt(c->inv_metric[i]);
  z_grad(c);
  c->H0 = c->V + kinetic(c, c->p);
  double Vf = c->V, Vb = c->V, V_sample = c->V;
  memcpy(qf, c->q, sizeof(double) * D); memcpy(pf, c->p, sizeof(double) * D); memcpy(gf, c->g, sizeof(double) * D);
  memcpy(qb, c->q, sizeof(double) * D); memcpy(pb, c->p, sizeof(double) * D); memcpy(gb, c->g, sizeof(double) * D);
  memcpy(z_sample, c->q, sizeof(double) * D);
  memcpy(rho, c->p, sizeof(double) * D); memcpy(p_bb, c->p, sizeof(double) * D); memcpy(p_ff, c->p, sizeof(double) * D);
  double lsw = 0;
  int depth = 0;
  c->n_leapfrog = 0; c->sum_metro = 0; c->divergent = 0; c->merge_ctr = 0;
  while (depth < c->max_depth) {
    uint32_t w[4];
    orc_rng_words(c->seed, c->chain, (uint32_t)depth, c->iter, 2, 0, w);
    int fwd = u01(w[0]) > 0.5;
    double lsw_sub = -INFINITY;
    if (fwd) { memcpy(c->q, qf, sizeof(double) * D); memcpy(c->p, pf, sizeof(double) * D); memcpy(c->g, gf, sizeof(double) * D); c->V = Vf; }
    else { memcpy(c->q, qb, sizeof(double) * D); memcpy(c->p, pb, sizeof(double) * D); memcpy(c->g, gb, sizeof(double) * D); c->V = Vb; }
    int ok;
    if (c->tree_mode == 0) {
      memset(rho_new, 0, sizeof(double) * D);
      ok = build_rec(c, depth, fwd ? 1.0 : -1.0, z_prop, first, last, rho_new, &lsw_sub);
    } else {
      ok = build_iter(c, depth, fwd ? 1.0 : -1.0, z_prop, first, last, rho_new, &lsw_sub, left);
    }
    if (fwd) { memcpy(qf, c->q, sizeof(double) * D); memcpy(pf, c->p, sizeof(double) * D); memcpy(gf, c->g, sizeof(double) * D); Vf = c->V; }
    else { memcpy(qb, c->q, sizeof(double) * D); memcpy(pb, c->p, sizeof(double) * D); memcpy(gb, c->g, sizeof(double) * D); Vb = c->V; }
    if (!ok) break;
    ++depth;
    int take = (lsw_sub > lsw) ? 1 : (u01(w[1]) < exp(lsw_sub - lsw));
    if (take) memcpy(z_sample, z_prop, sizeof(double) * D);
    lsw = logaddexp(lsw, lsw_sub);
    /* top-level U-turn: L = old tree {far end F, near end A, rho}, R = new subtree */
    const double* F = fwd ? p_bb : p_ff; const double* A = fwd ? p_ff : p_bb;
    for (int i = 0; i < D; ++i) ext[i] = rho[i] + rho_new[i];
    int persist = crit(c, F, last, ext);
    for (int i = 0; i < D; ++i) ext[i] = rho[i] + first[i];
    persist = persist && crit(c, F, first, ext);
    for (int i = 0; i < D; ++i) ext[i] = rho_new[i] + A[i];
    persist = persist && crit(c, A, last, ext);
    for (int i = 0; i < D; ++i) rho[i] += rho_new[i];
    memcpy(fwd ? p_ff : p_bb, last, sizeof(double) * D);
    if (!persist) break;
  }
  memcpy(c->q, z_sample, sizeof(double) * D);
  z_grad(c); /* lp and energy of the selected point (Stan carries them inside z_sample) */
  V_sample = c->V;
  st->lp = -V_sample; st->accept_stat = c->sum_metro / (double)(c->n_leapfrog > 0 ? c->n_leapfrog : 1);
  st->stepsize = c->eps; st->treedepth = depth; st->n_leapfrog = c->n_leapfrog; st->divergent = c->divergent;
  st->energy = NAN; /* H(z_sample) needs the sample's momentum, which only mode-specific code has; not compared */
  for (int k = 0; k < c->max_depth; ++k) { free(left[k].b); free(left[k].e); free(left[k].r); }
  free(qf); free(pf); free(gf); free(qb); free(pb); free(gb); free(z_sample); free(z_prop); free(rho); free(p_bb); free(p_ff);
  free(first); free(last); free(rho_new); free(ext);
}

/* base_hmc::init_stepsize */
static void init_stepsize(Chain* c, uint32_t iter_tag) {
  const int D = c->D;
  if (c->eps == 0 || c->eps > 1e7 || isnan(c->eps)) return;
  double *q0 = vnew(D);
  memcpy(q0, c->q, sizeof(double) * D);
  int direction = 0;
  for (uint32_t attempt = 0; attempt < 200; ++attempt) {
    memcpy(c->q, q0, sizeof(double) * D);
    for (int i = 0; i < D; ++i) c->p[i] = rng_normal(c->seed, c->chain, (uint32_t)i, iter_tag, 5, attempt) / sqrt(c->inv_metric[i]);
    z_grad(c);
    double H0 = c->V + kinetic(c, c->p);
    leapfrog(c, c->eps);
    double h = c->V + kinetic(c, c->p);
    if (isnan(h)) h = INFINITY;
    double dH = H0 - h;
    if (attempt == 0) { direction = dH > log(0.8) ? 1 : -1; continue; }
    if (direction == 1 && !(dH > log(0.8))) break;
    if (direction == -1 && !(dH < log(0.8))) break;
    c->eps = direction == 1 ? 2 * c->eps : 0.5 * c->eps;
    if (c->eps > 1e7 || c->eps == 0) break;
  }
  memcpy(c->q, q0, sizeof(double) * D);
  free(q0);
}

typedef struct {
  /* dual averaging */
  double mu, s_bar, x_bar, delta, gamma, kappa, t0; int counter;
  /* windows */
  int num_warmup, init_buffer, term_buffer, base_window, window_size, next_window, wcounter;
  /* welford */
  double *mean, *m2; int nsamp;
} Adapt;

static void adapt_init(Adapt* a, int D, int num_warmup, double delta) {
  memset(a, 0, sizeof(*a));
  a->delta = delta; a->gamma = 0.05; a->kappa = 0.75; a->t0 = 10;
  a->num_warmup = num_warmup; a->init_buffer = 75; a->term_buffer = 50; a->base_window = 25;
  if (num_warmup < 20) { a->init_buffer = num_warmup; a->term_buffer = 0; a->base_window = 0; } /* no metric adaptation */
  else if (a->init_buffer + a->base_window + a->term_buffer > num_warmup) {
    a->init_buffer = (int)(0.15 * num_warmup); a->term_buffer = (int)(0.1 * num_warmup);
    a->base_window = num_warmup - (a->init_buffer + a->term_buffer);
  }
  a->window_size = a->base_window; a->next_window = a->init_buffer + a->window_size - 1;
  if (num_warmup < 20) a->next_window = -1; /* Stan: adapt_next_window_ stays UINT_MAX, no window ever ends */
  a->mean = (double*)calloc(D, sizeof(double)); a->m2 = (double*)calloc(D, sizeof(double));
}
static void learn_stepsize(Adapt* a, double* eps, double adapt_stat) {
  a->counter++;
  if (adapt_stat > 1) adapt_stat = 1;
  double eta = 1.0 / (a->counter + a->t0);
  a->s_bar = (1 - eta) * a->s_bar + eta * (a->delta - adapt_stat);
  double x = a->mu - a->s_bar * sqrt((double)a->counter) / a->gamma;
  double x_eta = pow((double)a->counter, -a->kappa);
  a->x_bar = (1 - x_eta) * a->x_bar + x_eta * x;
  *eps = exp(x);
}
static int learn_variance(Adapt* a, Chain* c) {
  const int D = c->D;
  int in_window = a->wcounter >= a->init_buffer && a->wcounter < a->num_warmup - a->term_buffer && a->wcounter != a->num_warmup;
  if (in_window) {
    a->nsamp++;
    for (int i = 0; i < D; ++i) { double d = c->q[i] - a->mean[i]; a->mean[i] += d / a->nsamp; a->m2[i] += (c->q[i] - a->mean[i]) * d; }
  }
  int end_window = a->wcounter == a->next_window && a->wcounter != a->num_warmup;
  if (end_window) {
    /* compute_next_window */
    if (a->next_window != a->num_warmup - a->term_buffer - 1) {
      a->window_size *= 2;
      a->next_window = a->wcounter + a->window_size;
      if (a->next_window != a->num_warmup - a->term_buffer - 1) {
        int boundary = a->next_window + 2 * a->window_size;
        if (boundary >= a->num_warmup - a->term_buffer) a->next_window = a->num_warmup - a->term_buffer - 1;
      }
    }
    double n = (double)a->nsamp;
    for (int i = 0; i < D; ++i) {
      double var = a->m2[i] / (n - 1.0);
      c->inv_metric[i] = (n / (n + 5.0)) * var + 1e-3 * (5.0 / (n + 5.0));
    }
    a->nsamp = 0; memset(a->mean, 0, sizeof(double) * D); memset(a->m2, 0, sizeof(double) * D);
    a->wcounter++;
    return 1;
  }
  a->wcounter++;
  return 0;
}

typedef struct OrcRun {
  const OrcModel* m; PotusConfig cfg; int literal, tree_mode, n_threads;
  double* theta_out;   /* [chains][iter_sampling][D] or NULL */
  double* monitor;     /* [chains][iter_sampling][S+1] */
  double* stats;       /* [chains][iter_warmup+iter_sampling][7] */
  double* final_eps;   /* [chains] */
  int64_t* n_leapfrog; /* [chains] total */
  int next_chain; pthread_mutex_t mu;
  int max_iters;       /* optional cap on transitions per chain (bounded benchmarks); <=0: none */
  double* inv_metric_out; /* optional [chains][D]: the adapted diagonal of M^-1 at the end of the run */
} OrcRun;

static void run_chain(OrcRun* r, int ci) {
  const OrcModel* m = r->m; const int D = m->D;
  Chain c; memset(&c, 0, sizeof(c));
  c.m = m; c.wk = work_new(m); c.D = D; c.seed = r->cfg.seed; c.chain = (uint32_t)(r->cfg.chain_id_offset + ci);
  c.literal = r->literal; c.tree_mode = r->tree_mode; c.max_depth = r->cfg.max_treedepth;
  c.inv_metric = vnew(D); for (int i = 0; i < D; ++i) c.inv_metric[i] = 1.0;
  c.q = vnew(D); c.p = vnew(D); c.g = vnew(D); c.eps = 1.0;
  c.arena_cap = (size_t)(c.max_depth + 2) * 6 * D; c.arena = (double*)malloc(sizeof(double) * c.arena_cap); c.arena_top = 0;
  /* random inits U(-r, r), retry while lp/grad not finite (<= 100 attempts) */
  for (uint32_t attempt = 0; attempt < 100; ++attempt) {
    for (int i = 0; i < D; ++i) {
      uint32_t w[4]; orc_rng_words(c.seed, c.chain, (uint32_t)i, 0, 0, attempt, w);
      c.q[i] = r->cfg.init_radius * (2.0 * u01(w[0]) - 1.0);
    }
    z_grad(&c);
    int fin = isfinite(c.V);
    for (int i = 0; i < D && fin; ++i) fin = isfinite(c.g[i]);
    if (fin) break;
  }
  Adapt a; adapt_init(&a, D, r->cfg.iter_warmup, r->cfg.adapt_delta);
  init_stepsize(&c, 0xFFFFFFFFu);
  a.mu = log(10 * c.eps);
  const int nw = r->cfg.iter_warmup, ns = r->cfg.iter_sampling;
  int64_t total_lf = 0;
  for (int it = 0; it < nw + ns; ++it) {
    if (r->max_iters > 0 && it >= r->max_iters) break;
    c.iter = (uint32_t)it;
    SampStats st;
    transition(&c, &st);
    total_lf += c.n_leapfrog;
    double* so = r->stats + ((size_t)ci * (nw + ns) + it) * 7;
    so[0] = st.lp; so[1] = st.accept_stat; so[2] = st.stepsize; so[3] = st.treedepth; so[4] = st.n_leapfrog; so[5] = st.divergent; so[6] = st.energy;
    if (it < nw) {
      learn_stepsize(&a, &c.eps, st.accept_stat);
      if (learn_variance(&a, &c)) {
        init_stepsize(&c, (uint32_t)it);
        a.mu = log(10 * c.eps); a.counter = 0; a.s_bar = 0; a.x_bar = 0;
      }
      if (it == nw - 1) c.eps = exp(a.x_bar);
    } else {
      int k = it - nw;
      if (r->theta_out) memcpy(r->theta_out + ((size_t)ci * ns + k) * D, c.q, sizeof(double) * D);
      /* monitored scalars: mu_b[,T] and national_mu_b_average[T]; c.wk->mu holds mu_b at c.q
       * because transition() ends with a gradient evaluation at the selected point */
      double* mo = r->monitor + ((size_t)ci * ns + k) * (m->S + 1);
      for (int s = 0; s < m->S; ++s) mo[s] = c.wk->mu[(m->T - 1) * m->S + s];
      mo[m->S] = c.wk->nat[m->T - 1];
    }
  }
  r->final_eps[ci] = c.eps; r->n_leapfrog[ci] = total_lf;
  if (r->inv_metric_out) memcpy(r->inv_metric_out + (size_t)ci * D, c.inv_metric, sizeof(double) * D);
  free(a.mean); free(a.m2); free(c.inv_metric); free(c.q); free(c.p); free(c.g); free(c.arena); work_free(c.wk);
}
static void* worker(void* arg) {
  OrcRun* r = (OrcRun*)arg;
  for (;;) {
    pthread_mutex_lock(&r->mu);
    int ci = r->next_chain++;
    pthread_mutex_unlock(&r->mu);
    if (ci >= r->cfg.chains) break;
    run_chain(r, ci);
  }
  return NULL;
}

/* Run `cfg->chains` chains of Stan-semantics NUTS on n_threads host threads (one chain per thread
 * at a time).  Outputs are caller-allocated; theta_out may be NULL.  Returns wall seconds. */
ORC_API double orc_sample(const OrcModel* m, const PotusConfig* cfg, int literal, int tree_mode, int n_threads, int max_iters,
                          double* theta_out, double* monitor, double* stats, double* final_eps, int64_t* n_leapfrog,
                          double* inv_metric_out) {
  OrcRun r; memset(&r, 0, sizeof(r));
  r.m = m; r.cfg = *cfg; r.literal = literal; r.tree_mode = tree_mode; r.n_threads = n_threads; r.max_iters = max_iters;
  r.inv_metric_out = inv_metric_out; r.theta_out = theta_out; r.monitor = monitor; r.stats = stats; r.final_eps = final_eps; r.n_leapfrog = n_leapfrog;
  pthread_mutex_init(&r.mu, NULL);
  struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
  if (n_threads < 1) n_threads = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
  for (int i = 0; i < n_threads; ++i) pthread_create(&th[i], NULL, worker, &r);
  for (int i = 0; i < n_threads; ++i) pthread_join(th[i], NULL);
  free(th);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  pthread_mutex_destroy(&r.mu);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* Debug hook for decision-level comparison with the CUDA kernel: run `n_iter` transitions of ONE
 * chain from a given point with fixed step size and inverse metric; returns stats[n_iter][7] and
 * the points after each transition. */
ORC_API void orc_transitions(const OrcModel* m, uint64_t seed, uint32_t chain, int tree_mode, int max_depth, double eps,
                             const double* inv_metric, const double* q0, uint32_t iter0, int n_iter, double* q_out, double* stats,
                             int literal) {
  Chain c; memset(&c, 0, sizeof(c));
  const int D = m->D;
  c.m = m; c.wk = work_new(m); c.D = D; c.seed = seed; c.chain = chain; c.literal = literal; c.tree_mode = tree_mode; c.max_depth = max_depth;
  c.inv_metric = vnew(D); memcpy(c.inv_metric, inv_metric, sizeof(double) * D);
  c.q = vnew(D); c.p = vnew(D); c.g = vnew(D); c.eps = eps;
  c.arena_cap = (size_t)(max_depth + 2) * 6 * D; c.arena = (double*)malloc(sizeof(double) * c.arena_cap); c.arena_top = 0;
  memcpy(c.q, q0, sizeof(double) * D);
  for (int it = 0; it < n_iter; ++it) {
    c.iter = iter0 + (uint32_t)it;
    SampStats st; transition(&c, &st);
    double* so = stats + (size_t)it * 7;
    so[0] = st.lp; so[1] = st.accept_stat; so[2] = st.stepsize; so[3] = st.treedepth; so[4] = st.n_leapfrog; so[5] = st.divergent; so[6] = st.energy;
    memcpy(q_out + (size_t)it * D, c.q, sizeof(double) * D);
  }
  free(c.inv_metric); free(c.q); free(c.p); free(c.g); free(c.arena); work_free(c.wk);
}
