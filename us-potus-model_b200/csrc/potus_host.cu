// Host side of the C-ABI (include/potus_b200.h): data validation (Stan's data-block constraints),
// construction of the device-side model (Cholesky, UMMA operand planes, sorted polls, segment-sum
// task lists, owner-layout maps), launches, and output reshaping to rstan::extract's layout.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <dlfcn.h>
#include <nccl.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>
#include "../../include/potus_b200.h"
#include "potus_layout.h"
#include "potus_stream_layout.h"

// (compiled as one translation unit with potus_kernel.cu through potus_lib.cu, which defines the kernels)
using namespace potus;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_TRY(x)                                                                                         \
  do {                                                                                                      \
    cudaError_t e_ = (x);                                                                                   \
    if (e_ != cudaSuccess)                                                                                  \
      return fail(POTUS_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_));                         \
  } while (0)

namespace {

struct Offsets { int zT, Z, c, m, pop, umu, urho, ze, xn, xs, zb, D; };

struct HostModel {
  ModelDev m{};
  Offsets o{};
  int S, T, P, M, Pop, Nn, Ns, N, full;
  std::vector<int32_t> map_i2s;            // [VEC]
  std::vector<void*> dev_allocs;
  double nat_sd;
};

template <class Tv>
int upload(HostModel& hm, const std::vector<Tv>& v, const void** out) {
  void* d = nullptr;
  size_t bytes = std::max<size_t>(v.size() * sizeof(Tv), 16);
  CUDA_TRY(cudaMalloc(&d, bytes));
  hm.dev_allocs.push_back(d);
  if (!v.empty()) CUDA_TRY(cudaMemcpy(d, v.data(), v.size() * sizeof(Tv), cudaMemcpyHostToDevice));
  *out = d;
  return POTUS_OK;
}

void free_model(HostModel& hm) {
  for (void* p : hm.dev_allocs) cudaFree(p);
  hm.dev_allocs.clear();
}

// Stan data-block constraints, poll_model_2020.stan:2-40
int validate(const PotusData* d) {
  char buf[256];
  auto bad = [&](const char* what, long i, double v, const char* cons) {
    snprintf(buf, sizeof buf, "Exception: poll_model_2020: %s[%ld] is %g, but must be %s", what, i + 1, v, cons);
    return fail(POTUS_ERR_INVALID_DATA, buf);
  };
  if (!d) return fail(POTUS_ERR_INVALID_DATA, "data is NULL");
  if (d->S < 1 || d->T < 2 || d->P < 1 || d->N_state_polls < 0 || d->N_national_polls < 0)
    return fail(POTUS_ERR_INVALID_DATA, "S, T, P, N_state_polls, N_national_polls must be positive sizes");
  const bool full = d->poll_mode_state != nullptr;
  if (full && (!d->poll_mode_national || !d->poll_pop_state || !d->poll_pop_national || !d->unadjusted_state || !d->unadjusted_national))
    return fail(POTUS_ERR_INVALID_DATA, "poll_mode_*/poll_pop_*/unadjusted_* must be given together (full model) or all absent");
  if (!d->state || !d->day_state || !d->day_national || !d->poll_state || !d->poll_national || !d->n_democrat_national ||
      !d->n_two_share_national || !d->n_democrat_state || !d->n_two_share_state || !d->mu_b_prior || !d->state_weights ||
      !d->state_covariance_0)
    return fail(POTUS_ERR_INVALID_DATA, "a required data vector is NULL");
  for (int i = 0; i < d->N_state_polls; ++i) {
    if (d->state[i] < 1 || d->state[i] > d->S + 1) return bad("state", i, d->state[i], "in [1, S+1]");
    if (d->state[i] > d->S) return bad("state", i, d->state[i], "<= S (mu_b has S rows; index S+1 would fault in Stan)");
    if (d->day_state[i] < 1 || d->day_state[i] > d->T) return bad("day_state", i, d->day_state[i], "in [1, T]");
    if (d->poll_state[i] < 1 || d->poll_state[i] > d->P) return bad("poll_state", i, d->poll_state[i], "in [1, P]");
    if (d->n_two_share_state[i] < 0) return bad("n_two_share_state", i, d->n_two_share_state[i], ">= 0");
    if (d->n_democrat_state[i] < 0 || d->n_democrat_state[i] > d->n_two_share_state[i])
      return bad("n_democrat_state", i, d->n_democrat_state[i], "in [0, n_two_share_state]");
    if (full) {
      if (d->poll_mode_state[i] < 1 || d->poll_mode_state[i] > d->M) return bad("poll_mode_state", i, d->poll_mode_state[i], "in [1, M]");
      if (d->poll_pop_state[i] < 1 || d->poll_pop_state[i] > d->Pop) return bad("poll_pop_state", i, d->poll_pop_state[i], "in [1, Pop]");
      if (!(d->unadjusted_state[i] >= 0 && d->unadjusted_state[i] <= 1)) return bad("unadjusted_state", i, d->unadjusted_state[i], "in [0, 1]");
    }
  }
  for (int i = 0; i < d->N_national_polls; ++i) {
    if (d->day_national[i] < 1 || d->day_national[i] > d->T) return bad("day_national", i, d->day_national[i], "in [1, T]");
    if (d->poll_national[i] < 1 || d->poll_national[i] > d->P) return bad("poll_national", i, d->poll_national[i], "in [1, P]");
    if (d->n_two_share_national[i] < 0) return bad("n_two_share_national", i, d->n_two_share_national[i], ">= 0");
    if (d->n_democrat_national[i] < 0 || d->n_democrat_national[i] > d->n_two_share_national[i])
      return bad("n_democrat_national", i, d->n_democrat_national[i], "in [0, n_two_share_national]");
    if (full) {
      if (d->poll_mode_national[i] < 1 || d->poll_mode_national[i] > d->M) return bad("poll_mode_national", i, d->poll_mode_national[i], "in [1, M]");
      if (d->poll_pop_national[i] < 1 || d->poll_pop_national[i] > d->Pop) return bad("poll_pop_national", i, d->poll_pop_national[i], "in [1, Pop]");
      if (!(d->unadjusted_national[i] >= 0 && d->unadjusted_national[i] <= 1)) return bad("unadjusted_national", i, d->unadjusted_national[i], "in [0, 1]");
    }
  }
  const int S = d->S;  // cov_matrix[S] state_covariance_0: symmetric (Stan's 1e-8 relative tolerance), PD checked by Cholesky
  for (int i = 0; i < S; ++i)
    for (int j = 0; j < i; ++j) {
      double a = d->state_covariance_0[i + S * j], b = d->state_covariance_0[j + S * i];
      if (std::fabs(a - b) > 1e-8 * std::max(1.0, std::max(std::fabs(a), std::fabs(b)))) {
        snprintf(buf, sizeof buf, "Exception: poll_model_2020: state_covariance_0 is not symmetric. state_covariance_0[%d,%d] = %g, but [%d,%d] = %g", i + 1, j + 1, a, j + 1, i + 1, b);
        return fail(POTUS_ERR_INVALID_DATA, buf);
      }
    }
  {  // positive definite (Stan checks cov_matrix with an LDLT; a failing Cholesky is the same condition)
    std::vector<double> L((size_t)S * S, 0.0);
    for (int j = 0; j < S; ++j) {
      double s = d->state_covariance_0[j + S * j];
      for (int k = 0; k < j; ++k) s -= L[j * S + k] * L[j * S + k];
      if (!(s > 0)) return fail(POTUS_ERR_INVALID_DATA, "Exception: poll_model_2020: state_covariance_0 is not positive definite.");
      double ljj = std::sqrt(s);
      L[j * S + j] = ljj;
      for (int i = j + 1; i < S; ++i) {
        double v = d->state_covariance_0[i + S * j];
        for (int k = 0; k < j; ++k) v -= L[i * S + k] * L[j * S + k];
        L[i * S + j] = v / ljj;
      }
    }
  }
  return POTUS_OK;
}

int check_supported(const PotusData* d) {
  char buf[256];
  const bool full = d->poll_mode_state != nullptr;
  const int N = d->N_state_polls + d->N_national_polls;
  if (d->S > MAX_S || d->T > MAX_T || N > NPOLL_CAP - SEG || d->P > 1023 || (full && (d->M > MAX_MODE || d->Pop > MAX_MODE))) {
    snprintf(buf, sizeof buf,
             "problem size S=%d T=%d N=%d P=%d M=%d Pop=%d is outside the resident kernel's limits (S<=%d, T<=%d, N<=%d, P<=1023, M,Pop<=%d); "
             "the streaming large-S/T variant (BASELINE config 5) is not built yet",
             d->S, d->T, N, d->P, d->M, d->Pop, MAX_S, MAX_T, NPOLL_CAP - SEG, MAX_MODE);
    return fail(POTUS_ERR_UNSUPPORTED, buf);
  }
  if (full)
    for (int pass = 0; pass < 2; ++pass) {
      const double* u = pass ? d->unadjusted_national : d->unadjusted_state;
      int n = pass ? d->N_national_polls : d->N_state_polls;
      for (int i = 0; i < n; ++i)
        if (u[i] != 0.0 && u[i] != 1.0) return fail(POTUS_ERR_UNSUPPORTED, "unadjusted_* must be 0 or 1 (fractional values are not supported)");
    }
  return POTUS_OK;
}

void split_half(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * 2048.0f);
}

int build_model(const PotusData* d, HostModel& hm) {
  int rc = validate(d);
  if (rc) return rc;
  rc = check_supported(d);
  if (rc) return rc;
  const int S = d->S, T = d->T, P = d->P, Ns = d->N_state_polls, Nn = d->N_national_polls, N = Ns + Nn;
  const bool full = d->poll_mode_state != nullptr;
  const int M = full ? d->M : 0, Pop = full ? d->Pop : 0;
  // (the no-mode model has no mu_m / mu_pop blocks although its data list still carries M and Pop:
  //  the draw record the kernel writes uses M = Pop = 0 there, and so must every host-side offset)
  hm.S = S; hm.T = T; hm.P = P; hm.M = M; hm.Pop = Pop; hm.Ns = Ns; hm.Nn = Nn; hm.N = N; hm.full = full;
  // Stan unconstrained order (poll_model_2020.stan:56-69)
  Offsets& o = hm.o;
  int off = 0;
  o.zT = off; off += S;
  o.Z = off; off += S * T;
  o.c = off; off += P;
  o.m = o.pop = o.umu = o.urho = o.ze = -1;
  if (full) { o.m = off; off += M; o.pop = off; off += Pop; o.umu = off; off += 1; o.urho = off; off += 1; o.ze = off; off += T; }
  o.xn = off; off += Nn;
  o.xs = off; off += Ns;
  o.zb = off; off += S;
  o.D = off;

  // ---- transformed data (poll_model_2020.stan:42-55): nat_sd, L0 = chol(Sigma0), three scalars
  std::vector<double> L0((size_t)S * S, 0.0), w(d->state_weights, d->state_weights + S);
  double nat = 0;
  for (int i = 0; i < S; ++i)
    for (int j = 0; j < S; ++j) nat += w[i] * d->state_covariance_0[i + S * j] * w[j];
  if (!(nat > 0)) return fail(POTUS_ERR_INVALID_DATA, "state_weights' * state_covariance_0 * state_weights must be positive");
  nat = std::sqrt(nat);
  hm.nat_sd = nat;
  for (int j = 0; j < S; ++j) {
    double s = d->state_covariance_0[j + S * j];
    for (int k = 0; k < j; ++k) s -= L0[j * S + k] * L0[j * S + k];
    if (!(s > 0)) return fail(POTUS_ERR_INVALID_DATA, "Exception: poll_model_2020: state_covariance_0 is not positive definite.");
    double ljj = std::sqrt(s);
    L0[j * S + j] = ljj;
    for (int i = j + 1; i < S; ++i) {
      double v = d->state_covariance_0[i + S * j];
      for (int k = 0; k < j; ++k) v -= L0[i * S + k] * L0[j * S + k];
      L0[i * S + j] = v / ljj;
    }
  }
  ModelDev& m = hm.m;
  m.S = S; m.T = T; m.P = P; m.M = M; m.Pop = Pop; m.Nn = Nn; m.Ns = Ns; m.N = N; m.full = full; m.D = o.D;
  m.npair = (S + 1) / 2;
  m.a_b = (float)(d->polling_bias_scale / nat); m.a_T = (float)(d->mu_b_T_scale / nat); m.a_w = (float)(d->random_walk_scale / nat);
  m.sig_c = (float)d->sigma_c; m.sig_m = (float)d->sigma_m; m.sig_pop = (float)d->sigma_pop;
  m.sig_n = (float)d->sigma_measure_noise_national; m.sig_s = (float)d->sigma_measure_noise_state; m.sig_e = (float)d->sigma_e_bias;
  // nz block layout
  int nz = 0;
  m.nz_zT = nz; nz += S;
  m.nz_c = nz; nz += P;
  m.nz_m = m.nz_pop = m.nz_umu = m.nz_urho = m.nz_ze = 0;
  if (full) { m.nz_m = nz; nz += M; m.nz_pop = nz; nz += Pop; m.nz_umu = nz; nz += 1; m.nz_urho = nz; nz += 1; m.nz_ze = nz; nz += T; }
  m.nz_x = nz; nz += N;
  m.nz_zb = nz; nz += S;
  m.NZ = nz;
  if (nz > NZ_CAP) return fail(POTUS_ERR_UNSUPPORTED, "too many non-walk parameters for the resident kernel (NZ > 3072)");

  // ---- polls sorted by (day, state); national polls by day
  std::vector<int> ord_s(Ns), ord_n(Nn);
  std::iota(ord_s.begin(), ord_s.end(), 0);
  std::iota(ord_n.begin(), ord_n.end(), 0);
  std::stable_sort(ord_s.begin(), ord_s.end(), [&](int a, int b) {
    if (d->day_state[a] != d->day_state[b]) return d->day_state[a] < d->day_state[b];
    return d->state[a] < d->state[b];
  });
  std::stable_sort(ord_n.begin(), ord_n.end(), [&](int a, int b) { return d->day_national[a] < d->day_national[b]; });
  struct HP { int s, dd, p, mo, po, un; double n, y; int stan_x; };
  std::vector<HP> hp(N);
  double sum_n = 0;
  for (int k = 0; k < Ns; ++k) {
    int i = ord_s[k];
    hp[k] = HP{d->state[i] - 1, d->day_state[i] - 1, d->poll_state[i] - 1, full ? d->poll_mode_state[i] - 1 : 0,
               full ? d->poll_pop_state[i] - 1 : 0, full ? (int)d->unadjusted_state[i] : 0, (double)d->n_two_share_state[i],
               (double)d->n_democrat_state[i], o.xs + i};
    sum_n += hp[k].n;
  }
  for (int k = 0; k < Nn; ++k) {
    int j = ord_n[k];
    hp[Ns + k] = HP{NAT_COL, d->day_national[j] - 1, d->poll_national[j] - 1, full ? d->poll_mode_national[j] - 1 : 0,
                    full ? d->poll_pop_national[j] - 1 : 0, full ? (int)d->unadjusted_national[j] : 0,
                    (double)d->n_two_share_national[j], (double)d->n_democrat_national[j], o.xn + j};
    sum_n += hp[Ns + k].n;
  }
  std::vector<uint32_t> pk(5 * (size_t)NPOLL_CAP, 0u);
  double lp_const = 0;
  for (int k = 0; k < N; ++k) {
    const HP& q = hp[k];
    double frac = q.n > 0 ? q.y / q.n : 0.5;
    double fc = std::min(std::max(frac, 1e-4), 1.0 - 1e-4);
    float eh = (float)std::log(fc / (1.0 - fc));
    float ph = (float)(1.0 / (1.0 + std::exp(-(double)eh)));
    float rh = (float)(frac - (double)ph);
    float nf = (float)q.n;
    pk[k] = pack_poll(q.s, q.dd, q.p, q.mo, q.po, q.un);
    memcpy(&pk[1 * NPOLL_CAP + k], &nf, 4);
    memcpy(&pk[2 * NPOLL_CAP + k], &eh, 4);
    memcpy(&pk[3 * NPOLL_CAP + k], &ph, 4);
    memcpy(&pk[4 * NPOLL_CAP + k], &rh, 4);
    // centring constant of the function the kernel evaluates: n[(ph+rh)(eta-eh) - log1p(ph expm1(eta-eh))] = ll(eta) - ll(eh)
    double ehd = eh, sp = (ehd > 0 ? ehd : 0) + std::log1p(std::exp(-std::fabs(ehd)));
    lp_const += q.y * ehd - q.n * sp;
  }
  m.lp_const = lp_const;
  {  // G operand scale: |sum of residuals| <= sum n must stay inside fp16 range
    double sc = 1.0;
    while (sum_n * sc > 32768.0) sc *= 0.5;
    m.scale_G = (float)sc; m.inv_scale_G = (float)(1.0 / sc);
  }

  // ---- X planes: X[r][c], r = output state (row 51 = national: w^T L0), c = k; scaled by 256; K-major canonical layout
  std::vector<__half> bt(2 * (size_t)B_PLANE / 2, __float2half_rn(0.f));
  auto boff = [](int r, int c) { return (size_t)((c / 8) * B_LBO + (r / 8) * B_SBO + (r % 8) * 16 + (c % 8) * 2) / 2; };
  for (int r = 0; r < 64; ++r)
    for (int c = 0; c < 64; ++c) {
      double v = 0;
      if (r < S && c < S) v = L0[r * S + c];
      else if (r == NAT_COL && c < S) { for (int s = 0; s < S; ++s) v += w[s] * L0[s * S + c]; }
      __half hi, lo;
      split_half((float)(v * 256.0), hi, lo);
      bt[boff(r, c)] = hi;
      bt[B_PLANE / 2 + boff(r, c)] = lo;
    }
  std::vector<float> prior(64, 0.f);
  double wp = 0;
  for (int s = 0; s < S; ++s) { prior[s] = (float)d->mu_b_prior[s]; wp += w[s] * d->mu_b_prior[s]; }
  prior[NAT_COL] = (float)wp;

  // ---- segment-sum tasks
  std::vector<uint2> t1;   // x = start | cnt<<16 | type<<24, y = psum slot (packed to one word below)
  std::vector<uint32_t> cells;
  std::vector<uint2> t2;
  std::vector<uint16_t> ids;
  auto add_segments = [&](int type, int start, int cnt, int& pstart, int& pcnt) {
    pstart = (int)t1.size(); pcnt = 0;
    while (cnt > 0) {
      int c = std::min(cnt, (int)SEG);
      t1.push_back(make_uint2((uint32_t)start | ((uint32_t)c << 16) | ((uint32_t)type << 24), (uint32_t)t1.size()));
      start += c; cnt -= c; ++pcnt;
    }
  };
  bool seg_overflow = false;
  auto add_final = [&](int kind, int dest, int pstart, int pcnt) {
    if (pcnt > 255 || pstart > 65535) seg_overflow = true;
    t2.push_back(make_uint2((uint32_t)pstart | ((uint32_t)pcnt << 16) | ((uint32_t)kind << 24), (uint32_t)dest));
  };
  // cells (state polls are sorted by day then state, national by day: groups are contiguous)
  for (int k = 0; k < N;) {
    int e = k;
    while (e < N && hp[e].dd == hp[k].dd && hp[e].s == hp[k].s && ((e < Ns) == (k < Ns))) ++e;
    if (e - k > 255) return fail(POTUS_ERR_UNSUPPORTED, "more than 255 polls in one (state, day) cell");
    if (e - k > 63 || k > 4095) return fail(POTUS_ERR_UNSUPPORTED, "poll cell does not fit the packed descriptor");
    cells.push_back((uint32_t)k | ((uint32_t)(e - k) << 12) | ((uint32_t)(hp[k].dd * 64 + hp[k].s) << 18));
    k = e;
  }
  if (full) {  // g_e[t] = sum over polls of day t of unadjusted * r
    std::vector<int> s0(T + 1, 0), n0(T + 1, 0);
    for (int k = 0; k < Ns; ++k) s0[hp[k].dd + 1]++;
    for (int k = 0; k < Nn; ++k) n0[hp[Ns + k].dd + 1]++;
    for (int t = 0; t < T; ++t) { s0[t + 1] += s0[t]; n0[t + 1] += n0[t]; }
    for (int t = 0; t < T; ++t) {
      int ps, pc, ps2, pc2;
      add_segments(1, s0[t], s0[t + 1] - s0[t], ps, pc);
      add_segments(1, Ns + n0[t], n0[t + 1] - n0[t], ps2, pc2);
      add_final(2, t, ps, pc + pc2);  // segments are appended back to back
    }
  }
  {  // pollsters
    std::vector<std::vector<uint16_t>> byp(P);
    for (int k = 0; k < N; ++k) byp[hp[k].p].push_back((uint16_t)k);
    for (int p = 0; p < P; ++p) {
      int st = (int)ids.size(), ps, pc;
      ids.insert(ids.end(), byp[p].begin(), byp[p].end());
      add_segments(2, st, (int)byp[p].size(), ps, pc);
      add_final(1, m.nz_c + p, ps, pc);
    }
    std::vector<std::vector<uint16_t>> bys(S);
    for (int k = 0; k < Ns; ++k) bys[hp[k].s].push_back((uint16_t)k);
    for (int s = 0; s < S; ++s) {
      int st = (int)ids.size(), ps, pc;
      ids.insert(ids.end(), bys[s].begin(), bys[s].end());
      add_segments(2, st, (int)bys[s].size(), ps, pc);
      add_final(0, PB_ROW * 64 + s, ps, pc);
    }
  }
  // balance: longest tasks first, so thread i of every warp gets tasks of similar length
  std::stable_sort(t1.begin(), t1.end(), [](const uint2& a, const uint2& b) { return ((a.x >> 16) & 0xff) > ((b.x >> 16) & 0xff); });
  std::stable_sort(t2.begin(), t2.end(), [](const uint2& a, const uint2& b) { return ((a.x >> 16) & 0xff) > ((b.x >> 16) & 0xff); });
  if ((int)t1.size() > NT1_CAP || (int)t2.size() > NT2_CAP || (int)t2.size() > NT || (int)ids.size() + SEG > NIDS_CAP || (int)cells.size() > NCELL_CAP)
    return fail(POTUS_ERR_UNSUPPORTED, "poll structure needs more segment tasks than the resident kernel holds");
  std::stable_sort(cells.begin(), cells.end(), [](uint32_t a, uint32_t b) { return ((a >> 12) & 63) > ((b >> 12) & 63); });
  std::vector<uint32_t> t1p(NT1_CAP, 0u);
  for (size_t i = 0; i < t1.size(); ++i) {
    const uint32_t start = t1[i].x & 0xffff, cnt = (t1[i].x >> 16) & 0xff, type = t1[i].x >> 24;
    if (start > 8191 || cnt > SEG || t1[i].y > 4095) return fail(POTUS_ERR_UNSUPPORTED, "segment task does not fit the packed descriptor");
    t1p[i] = start | (cnt << 13) | (type << 18) | (t1[i].y << 20);
  }
  cells.resize(NCELL_CAP, 0u);
  m.n_cell = 0; for (uint32_t c : cells) if ((c >> 12) & 63) ++m.n_cell;
  m.n_ids = (int)ids.size();
  ids.resize(NIDS_CAP, 0);
  if (seg_overflow) return fail(POTUS_ERR_UNSUPPORTED, "a poll group needs more than 255 segments");
  m.n_t1 = (int)t1.size(); m.n_t2 = (int)t2.size();

  // ---- owner-layout map: internal slot -> Stan unconstrained index
  hm.map_i2s.assign(VEC, -1);
  m.urho_owner = -1;
  for (int tid = 0; tid < NT; ++tid) {
    int wq = tid >> 5, l = tid & 31;
    for (int e = 0; e < EPT; ++e) {
      int idx = -1;
      if (l < ZLANES) {
        int s = 2 * l + (e & 1), t = DPW * wq + (e >> 1);
        if (s < S && t < T) idx = o.Z + s + S * t;
      } else {
        int k = nz_slot(wq, l - ZLANES, e);
        if (full && k == m.nz_urho) m.urho_owner = tid;
        if (k < m.NZ) {
          if (k < m.nz_c) idx = o.zT + (k - m.nz_zT);
          else if (k < m.nz_c + P) idx = o.c + (k - m.nz_c);
          else if (full && k < m.nz_pop) idx = o.m + (k - m.nz_m);
          else if (full && k < m.nz_umu) idx = o.pop + (k - m.nz_pop);
          else if (full && k == m.nz_umu) idx = o.umu;
          else if (full && k == m.nz_urho) idx = o.urho;
          else if (full && k < m.nz_x) idx = o.ze + (k - m.nz_ze);
          else if (k < m.nz_zb) idx = hp[k - m.nz_x].stan_x;
          else idx = o.zb + (k - m.nz_zb);
        }
      }
      hm.map_i2s[oslot(e, tid)] = idx;
    }
  }
  if (full && m.urho_owner < 0) return fail(POTUS_ERR_STATE, "rho_e_bias has no owner thread");
  {  // every Stan index must be covered exactly once
    std::vector<int> cnt(o.D, 0);
    for (int v : hm.map_i2s) if (v >= 0) { if (v >= o.D) return fail(POTUS_ERR_STATE, "internal map out of range"); cnt[v]++; }
    for (int i = 0; i < o.D; ++i) if (cnt[i] != 1) return fail(POTUS_ERR_STATE, "internal layout map is not a bijection");
  }
  const void* p;
  if ((rc = upload(hm, bt, &p))) return rc; m.btiles = p;
  if ((rc = upload(hm, pk, &p))) return rc; m.pk = (const uint32_t*)p;
  if ((rc = upload(hm, prior, &p))) return rc; m.prior = (const float*)p;
  if ((rc = upload(hm, t1p, &p))) return rc; m.t1 = (const uint32_t*)p;
  if ((rc = upload(hm, cells, &p))) return rc; m.cells = (const uint32_t*)p;
  if ((rc = upload(hm, t2, &p))) return rc; m.t2 = (const uint2*)p;
  if ((rc = upload(hm, ids, &p))) return rc; m.ids = (const uint16_t*)p;
  if ((rc = upload(hm, hm.map_i2s, &p))) return rc; m.map_i2s = (const int32_t*)p;
  return POTUS_OK;
}

int check_device(int device, int* n_sm) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(POTUS_ERR_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
  if (device < 0 || device >= count) return fail(POTUS_ERR_CUDA, "device ordinal out of range");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    char buf[200];
    snprintf(buf, sizeof buf, "device %d (%s) is sm_%d%d; this library contains sm_100a code only", device, prop.name, prop.major, prop.minor);
    return fail(POTUS_ERR_CUDA, buf);
  }
  if (n_sm) *n_sm = prop.multiProcessorCount;
  return POTUS_OK;
}

constexpr int SMEM_BYTES = (int)SM_TOTAL + 128;

}  // namespace

#include "potus_stream_host.cuh"

// ---- NCCL, resolved at run time (dlopen): the library loads and runs single-GPU without it; n_gpus > 1 needs libnccl.so.2
namespace {
struct NcclApi {
  void* h = nullptr;
  decltype(&ncclCommInitAll) CommInitAll = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
};
NcclApi* nccl_api() {
  static NcclApi api;
  if (api.h) return &api;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return nullptr;
  api.CommInitAll = (decltype(api.CommInitAll))dlsym(h, "ncclCommInitAll");
  api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
  api.GroupStart = (decltype(api.GroupStart))dlsym(h, "ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))dlsym(h, "ncclGroupEnd");
  if (!api.CommInitAll || !api.AllGather || !api.CommDestroy || !api.GetErrorString || !api.GroupStart || !api.GroupEnd) return nullptr;
  api.h = h;
  return &api;
}
void nccl_destroy_all(std::vector<ncclComm_t>& comms) {
  NcclApi* n = nccl_api();
  if (n) for (ncclComm_t c : comms) if (c) n->CommDestroy(c);
  comms.clear();
}
thread_local int g_alloc_chains_hint = 0;   // create_multi: every shard's draw buffer is sized for the largest shard
}  // namespace
struct PotusSampler;
static int create_multi(const PotusData* data, const PotusConfig* config, PotusSampler** out);

struct PotusSampler {
  HostModel hm;       // resident kernel (potus_kernel.cu)
  StreamHost sh;      // streaming kernel (potus_stream.cu): shapes the resident kernel does not hold
  bool stream = false;
  // what both paths share on the host: sizes, the vector length and the slot -> Stan index map
  int S = 0, T = 0, P = 0, M = 0, Pop = 0, full = 0, D = 0, VL = 0;
  double lp_const = 0;
  const std::vector<int32_t>* map = nullptr;
  float* rbuf = nullptr;
  // n_gpus > 1: this object is the parent of one sub-sampler per device (single process); `gath` holds, on every device, the
  // all-gathered kept-draw buffer [n_gpus][pad_draws][draw_len]
  std::vector<PotusSampler*> subs;
  std::vector<ncclComm_t> comms;
  std::vector<float*> gath;
  size_t pad_floats = 0;
  std::vector<float> h_w;       // state_weights (national vote of the post-processing)
  bool have_state = false;      // potus_set_state: chains start from given states; potus_run skips inits and warm-up
  int alloc_chains = 0;         // chains the draw buffer is sized for (>= chains; equal shard size for ncclAllGather)
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  int launches = 0;
  PotusConfig cfg;
  int n_sm = 0, grid = 0, draw_len = 0, keep = 0, keep_every = 1;
  float *q = nullptr, *sqrt_m = nullptr, *wf_mean = nullptr, *wf_m2 = nullptr, *workspace = nullptr;
  ChainState* cs = nullptr;
  int* queue = nullptr;
  float *draws = nullptr, *monitor = nullptr, *sparams = nullptr;
  unsigned long long* prof = nullptr;  // phase cycle counters (POTUS_PROF development builds)
  bool ran = false;
  PotusStats stats{};
  std::vector<float> h_draws, h_monitor, h_sparams;
  bool have_host = false;
};

// contiguous shard of `total` chains for part `rank` of `world` (the same rule bench.py's torchrun path uses)
static void shard_of(int total, int world, int rank, int* off, int* n) {
  const int base = total / world, rem = total % world;
  *n = base + (rank < rem ? 1 : 0);
  *off = rank * base + std::min(rank, rem);
}

extern "C" int potus_create(const PotusData* data, const PotusConfig* config, PotusSampler** out);
extern "C" void potus_destroy(PotusSampler* s);

// n_gpus > 1: ONE process, chains sharded contiguously over devices [device, device + n_gpus), the read-only
// model replicated on each, one ncclCommInitAll, and after sampling ONE ncclAllGather of the kept-draw buffers.
static int create_multi(const PotusData* data, const PotusConfig* config, PotusSampler** out) {
  const int G = config->n_gpus;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    return fail(POTUS_ERR_CUDA, "no CUDA device available; this library has no CPU fallback");
  if (config->device < 0 || config->device + G > count) {
    char b[160]; snprintf(b, sizeof b, "n_gpus = %d from device %d needs devices up to %d, but %d are visible", G, config->device, config->device + G - 1, count);
    return fail(POTUS_ERR_CUDA, b);
  }
  if (config->chains < G) return fail(POTUS_ERR_STATE, "config: chains >= n_gpus required");
  NcclApi* nc = nccl_api();
  if (!nc) return fail(POTUS_ERR_CUDA, "n_gpus > 1 needs NCCL (libnccl.so.2 could not be loaded)");
  PotusSampler* s = new PotusSampler();
  s->cfg = *config;
  int maxn = 0;
  for (int d = 0; d < G; ++d) { int off, n; shard_of(config->chains, G, d, &off, &n); maxn = std::max(maxn, n); }
  for (int d = 0; d < G; ++d) {
    int off, n;
    shard_of(config->chains, G, d, &off, &n);
    PotusConfig c = *config;
    c.n_gpus = 1; c.device = config->device + d; c.chains = n; c.chain_id_offset = config->chain_id_offset + off;
    PotusSampler* sub = nullptr;
    g_alloc_chains_hint = maxn;
    int rc = potus_create(data, &c, &sub);
    g_alloc_chains_hint = 0;
    if (rc) { const std::string msg = g_err; potus_destroy(s); return fail(rc, msg); }
    s->subs.push_back(sub);
  }
  const PotusSampler* s0 = s->subs[0];
  s->stream = s0->stream; s->S = s0->S; s->T = s0->T; s->P = s0->P; s->M = s0->M; s->Pop = s0->Pop; s->full = s0->full; s->D = s0->D; s->VL = s0->VL;
  s->lp_const = s0->lp_const; s->keep = s0->keep; s->keep_every = s0->keep_every; s->draw_len = s0->draw_len;
  s->pad_floats = (size_t)maxn * s->keep * s->draw_len;
  s->gath.assign(G, nullptr);
  for (int d = 0; d < G; ++d) {
    cudaSetDevice(config->device + d);
    cudaError_t e = cudaMalloc((void**)&s->gath[d], std::max<size_t>(s->pad_floats * G * sizeof(float), 16));
    if (e != cudaSuccess) { potus_destroy(s); return fail(POTUS_ERR_CUDA, std::string("cudaMalloc of the all-gather buffer: ") + cudaGetErrorString(e)); }
  }
  std::vector<int> devs(G);
  for (int d = 0; d < G; ++d) devs[d] = config->device + d;
  s->comms.assign(G, nullptr);
  ncclResult_t r = nc->CommInitAll(s->comms.data(), G, devs.data());
  if (r != ncclSuccess) { const std::string m = std::string("ncclCommInitAll: ") + nc->GetErrorString(r); potus_destroy(s); return fail(POTUS_ERR_CUDA, m); }
  *out = s;
  return POTUS_OK;
}

extern "C" {

const char* potus_last_error(void) { return g_err.c_str(); }

int potus_num_params(const PotusData* d) {
  if (!d) return 0;
  const bool full = d->poll_mode_state != nullptr;
  int D = d->S + d->S * d->T + d->P + d->N_national_polls + d->N_state_polls + d->S;
  if (full) D += d->M + d->Pop + 2 + d->T;
  return D;
}

void potus_destroy(PotusSampler* s) {
  if (!s) return;
  if (!s->subs.empty()) {
    for (size_t d = 0; d < s->subs.size(); ++d) {
      cudaSetDevice(s->subs[d]->cfg.device);
      if (d < s->gath.size()) cudaFree(s->gath[d]);
    }
    nccl_destroy_all(s->comms);
    for (PotusSampler* q : s->subs) potus_destroy(q);
    delete s;
    return;
  }
  cudaSetDevice(s->cfg.device);
  for (int i = 0; i < 3; ++i) if (s->ev[i]) cudaEventDestroy(s->ev[i]);
  free_model(s->hm);
  free_stream(s->sh);
  cudaFree(s->rbuf);
  cudaFree(s->q); cudaFree(s->sqrt_m); cudaFree(s->wf_mean); cudaFree(s->wf_m2); cudaFree(s->workspace);
  cudaFree(s->cs); cudaFree(s->queue); cudaFree(s->draws); cudaFree(s->monitor); cudaFree(s->sparams); cudaFree(s->prof);
  delete s;
}

int potus_create(const PotusData* data, const PotusConfig* config, PotusSampler** out) {
  if (!out) return fail(POTUS_ERR_STATE, "out is NULL");
  *out = nullptr;
  if (!config) return fail(POTUS_ERR_STATE, "config is NULL");
  if (config->chains < 1 || config->iter_warmup < 0 || config->iter_sampling < 0 || config->max_treedepth < 1 ||
      config->max_treedepth > MAX_DEPTH_CAP)
    return fail(POTUS_ERR_STATE, "config: chains >= 1, iter_* >= 0, 1 <= max_treedepth <= 10 required");
  int rc = validate(data);
  if (rc) return rc;
  if (config->n_gpus > 1) return create_multi(data, config, out);
  // kernel family: the resident kernel when the problem fits it (and config->reserved bit 0 does not force the streaming one)
  const bool want_stream = (config->flags & POTUS_FLAG_FORCE_STREAM) != 0;
  bool use_stream = want_stream;
  if (!want_stream && check_supported(data) != POTUS_OK) use_stream = true;   // (size, or fractional unadjusted_*)
  if (use_stream && (rc = check_stream_supported(data))) return rc;
  int n_sm = 0;
  rc = check_device(config->device, &n_sm);
  if (rc) return rc;
  PotusSampler* s = new PotusSampler();
  s->cfg = *config;
  s->n_sm = n_sm;
  s->stream = use_stream;
  s->h_w.assign(data->state_weights, data->state_weights + data->S);
  if (use_stream) {
    rc = build_stream_model(data, s->sh);
    if (rc) { potus_destroy(s); return rc; }
    const ModelS& ms = s->sh.m;
    s->S = ms.S; s->T = ms.T; s->P = ms.P; s->M = ms.M; s->Pop = ms.Pop; s->full = ms.full; s->D = ms.D; s->VL = ms.VL;
    s->lp_const = ms.lp_const; s->map = &s->sh.map_i2s;
  } else {
    rc = build_model(data, s->hm);
    if (rc == POTUS_ERR_UNSUPPORTED && check_stream_supported(data) == POTUS_OK) {
      // a shape detail the resident kernel cannot hold (e.g. more than 63 polls in one (state, day) cell): streaming family
      free_model(s->hm);
      s->stream = use_stream = true;
      rc = build_stream_model(data, s->sh);
      if (rc) { potus_destroy(s); return rc; }
      const ModelS& ms = s->sh.m;
      s->S = ms.S; s->T = ms.T; s->P = ms.P; s->M = ms.M; s->Pop = ms.Pop; s->full = ms.full; s->D = ms.D; s->VL = ms.VL;
      s->lp_const = ms.lp_const; s->map = &s->sh.map_i2s;
    } else {
      if (rc) { potus_destroy(s); return rc; }
      const ModelDev& mr = s->hm.m;
      s->S = mr.S; s->T = mr.T; s->P = mr.P; s->M = mr.M; s->Pop = mr.Pop; s->full = mr.full; s->D = mr.D; s->VL = VEC;
      s->lp_const = mr.lp_const; s->map = &s->hm.map_i2s;
    }
  }
  const int C = config->chains;
  s->alloc_chains = std::max(C, g_alloc_chains_hint);
  s->keep = config->keep_per_chain <= 0 ? config->iter_sampling : std::min(config->keep_per_chain, config->iter_sampling);
  s->keep_every = s->keep > 0 ? config->iter_sampling / s->keep : 1;
  s->draw_len = s->S * s->T + s->P + s->M + s->Pop + s->T + s->S + s->D;
  s->grid = std::min(C, n_sm);
  const size_t vb = (size_t)C * s->VL * sizeof(float);
  auto alloc = [&](void** p, size_t bytes) -> int {
    cudaError_t e = cudaMalloc(p, std::max<size_t>(bytes, 16));
    if (e != cudaSuccess) { char b[160]; snprintf(b, sizeof b, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e)); return fail(POTUS_ERR_CUDA, b); }
    return cudaMemset(*p, 0, std::max<size_t>(bytes, 16)) == cudaSuccess ? POTUS_OK : fail(POTUS_ERR_CUDA, "cudaMemset failed");
  };
  const int n_it = config->iter_warmup + config->iter_sampling;
  const size_t ws_bytes = (size_t)s->grid * (use_stream ? SW_NSLOT : NSLOT) * s->VL * sizeof(float);
  if ((rc = alloc((void**)&s->q, vb)) || (rc = alloc((void**)&s->sqrt_m, vb)) || (rc = alloc((void**)&s->wf_mean, vb)) ||
      (rc = alloc((void**)&s->wf_m2, vb)) || (rc = alloc((void**)&s->workspace, ws_bytes)) ||
      (rc = alloc((void**)&s->rbuf, use_stream ? (size_t)s->grid * s->sh.m.rb_len * sizeof(float) : 16)) ||
      (rc = alloc((void**)&s->cs, (size_t)C * sizeof(ChainState))) || (rc = alloc((void**)&s->queue, sizeof(int))) ||
      (rc = alloc((void**)&s->draws, (size_t)s->alloc_chains * s->keep * s->draw_len * sizeof(float))) ||
      (rc = alloc((void**)&s->monitor, (size_t)C * config->iter_sampling * (s->S + 1) * sizeof(float))) ||
      (rc = alloc((void**)&s->sparams, (size_t)C * n_it * 8 * sizeof(float))) || (rc = alloc((void**)&s->prof, 64 * sizeof(unsigned long long)))) {
    potus_destroy(s);
    return rc;
  }
  cudaError_t e = use_stream ? cudaFuncSetAttribute(potus_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SSMEM_BYTES)
                             : cudaFuncSetAttribute(potus_nuts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) { potus_destroy(s); return fail(POTUS_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e)); }
  *out = s;
  return POTUS_OK;
}

static RunArgs make_args(PotusSampler* s, int it0, int it1, int do_init) {
  RunArgs a{};
  a.m = s->hm.m;
  a.n_chains = s->cfg.chains; a.chain_id_offset = s->cfg.chain_id_offset;
  a.iter_begin = it0; a.iter_end = it1; a.iter_warmup = s->cfg.iter_warmup; a.iter_sampling = s->cfg.iter_sampling;
  a.max_depth = s->cfg.max_treedepth; a.do_init = do_init;
  a.keep_every = s->keep_every; a.keep_per_chain = s->keep; a.draw_len = s->draw_len;
  // Stan windowed_adaptation defaults (75 / 50 / 25), and its fallback rule for short warm-ups
  int nw = s->cfg.iter_warmup, ib = 75, tb = 50, bw = 25;
  if (nw < 20) { ib = nw; tb = 0; bw = 0; }
  else if (ib + bw + tb > nw) { ib = (int)(0.15 * nw); tb = (int)(0.1 * nw); bw = nw - (ib + tb); }
  a.w_init_buffer = ib; a.w_term_buffer = tb; a.w_base_window = bw;   // bw == 0 (nw < 20): no window ever ends, see the kernel's do_init block
  a.seed = s->cfg.seed; a.adapt_delta = (float)s->cfg.adapt_delta; a.init_radius = (float)s->cfg.init_radius;
  a.q = s->q; a.sqrt_m = s->sqrt_m; a.wf_mean = s->wf_mean; a.wf_m2 = s->wf_m2; a.cs = s->cs; a.workspace = s->workspace;
  a.queue = s->queue; a.draws = s->draws; a.monitor = s->monitor; a.sampler_params = s->sparams;
  a.prof = s->prof;
  return a;
}

static SRunArgs make_sargs(PotusSampler* s, int it0, int it1, int do_init) {
  const RunArgs r = make_args(s, it0, it1, do_init);
  SRunArgs a{};
  a.m = s->sh.m;
  a.n_chains = r.n_chains; a.chain_id_offset = r.chain_id_offset; a.iter_begin = it0; a.iter_end = it1;
  a.iter_warmup = r.iter_warmup; a.iter_sampling = r.iter_sampling; a.max_depth = r.max_depth; a.do_init = do_init;
  a.keep_every = r.keep_every; a.keep_per_chain = r.keep_per_chain; a.draw_len = r.draw_len;
  a.w_init_buffer = r.w_init_buffer; a.w_term_buffer = r.w_term_buffer; a.w_base_window = r.w_base_window;
  a.seed = r.seed; a.adapt_delta = r.adapt_delta; a.init_radius = r.init_radius;
  a.q = s->q; a.sqrt_m = s->sqrt_m; a.wf_mean = s->wf_mean; a.wf_m2 = s->wf_m2; a.cs = s->cs; a.workspace = s->workspace; a.rbuf = s->rbuf;
  a.queue = s->queue; a.draws = s->draws; a.monitor = s->monitor; a.sampler_params = s->sparams;
  a.prof = s->prof;
  return a;
}

// enqueue the warm-up and the sampling launch on the sampler's device (asynchronous)
static int run_launch(PotusSampler* s) {
  CUDA_TRY(cudaSetDevice(s->cfg.device));
  for (int i = 0; i < 3; ++i) if (!s->ev[i]) CUDA_TRY(cudaEventCreate(&s->ev[i]));
  const int nw = s->cfg.iter_warmup, nt = nw + s->cfg.iter_sampling;
  s->launches = 0;
  CUDA_TRY(cudaMemsetAsync(s->prof, 0, 64 * sizeof(unsigned long long)));
  CUDA_TRY(cudaEventRecord(s->ev[0]));
  if (!s->have_state) {
    CUDA_TRY(cudaMemsetAsync(s->queue, 0, sizeof(int)));
    if (s->stream) potus_stream_kernel<<<s->grid, SNT, SSMEM_BYTES>>>(make_sargs(s, 0, nw, 1));
    else potus_nuts_kernel<<<s->grid, NT, SMEM_BYTES>>>(make_args(s, 0, nw, 1));
    CUDA_TRY(cudaGetLastError());
    ++s->launches;
  }
  CUDA_TRY(cudaEventRecord(s->ev[1]));
  if (nt > nw) {
    CUDA_TRY(cudaMemsetAsync(s->queue, 0, sizeof(int)));
    if (s->stream) potus_stream_kernel<<<s->grid, SNT, SSMEM_BYTES>>>(make_sargs(s, nw, nt, 0));
    else potus_nuts_kernel<<<s->grid, NT, SMEM_BYTES>>>(make_args(s, nw, nt, 0));
    CUDA_TRY(cudaGetLastError());
    ++s->launches;
  }
  CUDA_TRY(cudaEventRecord(s->ev[2]));
  return POTUS_OK;
}

// wait for the launches of run_launch and form the summary statistics
static int run_finish(PotusSampler* s) {
  CUDA_TRY(cudaSetDevice(s->cfg.device));
  const int nw = s->cfg.iter_warmup, nt = nw + s->cfg.iter_sampling;
  CUDA_TRY(cudaEventSynchronize(s->ev[2]));
  CUDA_TRY(cudaGetLastError());
  float ms01 = 0, ms12 = 0;
  CUDA_TRY(cudaEventElapsedTime(&ms01, s->ev[0], s->ev[1])); CUDA_TRY(cudaEventElapsedTime(&ms12, s->ev[1], s->ev[2]));
  // summary statistics from the per-iteration sampler diagnostics, reduced on the device (the table itself goes to the host
  // only when potus_get_draws("sampler_params") asks for it)
  const int C = s->cfg.chains;
  s->h_sparams.clear();
  double hst[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  {
    double* dst = nullptr;
    CUDA_TRY(cudaMalloc(&dst, 8 * sizeof(double)));
    potus_post_runstats_kernel<<<1, PNT>>>(s->sparams, C, nt, nw, dst);
    cudaError_t e = cudaMemcpy(hst, dst, 8 * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(dst);
    if (e != cudaSuccess) return fail(POTUS_ERR_CUDA, std::string("potus_post_runstats_kernel: ") + cudaGetErrorString(e));
    ++s->launches;
  }
  std::vector<ChainState> hcs(C);
  CUDA_TRY(cudaMemcpy(hcs.data(), s->cs, (size_t)C * sizeof(ChainState), cudaMemcpyDeviceToHost));
  PotusStats& st = s->stats;
  st = PotusStats{};
  double acc = hst[3], dep = hst[4], eps = 0;
  for (int c = 0; c < C; ++c) {
    if (hcs[c].status != 0) return fail(POTUS_ERR_INIT, "a chain found no finite initial point in 100 attempts (Stan: 'Initialization failed')");
    eps += hcs[c].eps;
  }
  st.n_leapfrog_total = (int64_t)hst[0]; st.n_leapfrog_sampling = (int64_t)hst[1]; st.n_divergent_sampling = (int64_t)hst[2];
  const double ns = (double)C * std::max(1, nt - nw);
  st.mean_accept_stat = acc / ns; st.mean_treedepth = dep / ns; st.mean_stepsize = eps / C;
  st.gpu_launches = s->launches;
  st.seconds_warmup = ms01 * 1e-3; st.seconds_sampling = ms12 * 1e-3; st.seconds_total = (ms01 + ms12) * 1e-3;
  st.n_params = s->D; st.n_draws_kept = C * s->keep;
#ifdef POTUS_PROF
  {
    unsigned long long hp[64];
    if (cudaMemcpy(hp, s->prof, sizeof hp, cudaMemcpyDeviceToHost) == cudaSuccess) {
      const unsigned long long nl = s->stream ? hp[31] : hp[39];
      fprintf(stderr, "[potus_prof] leaves=%llu cycles/leaf per phase (thread 0, summed over CTAs):", nl);
      for (int i = 0; i < 24; ++i) fprintf(stderr, " p%d=%.0f", i, nl ? (double)hp[i] / (double)nl : 0.0);
      fprintf(stderr, "\n");
    }
  }
#endif
  s->ran = true; s->have_host = false;
  return POTUS_OK;
}

static int run_multi(PotusSampler* s) {
  NcclApi* nc = nccl_api();
  const int G = (int)s->subs.size();
  int rc;
  for (PotusSampler* q : s->subs) if ((rc = run_finish(q))) return rc;
  // the path's one exchange: all-gather of the kept draws over NVLink
  cudaEvent_t g0, g1;
  CUDA_TRY(cudaSetDevice(s->subs[0]->cfg.device));
  CUDA_TRY(cudaEventCreate(&g0)); CUDA_TRY(cudaEventCreate(&g1));
  CUDA_TRY(cudaEventRecord(g0));
  float gms = 0.f;
  if (s->pad_floats > 0) {
    ncclResult_t r = nc->GroupStart();
    for (int d = 0; d < G && r == ncclSuccess; ++d) {
      CUDA_TRY(cudaSetDevice(s->subs[d]->cfg.device));
      r = nc->AllGather(s->subs[d]->draws, s->gath[d], s->pad_floats, ncclFloat, s->comms[d], 0);
    }
    if (r == ncclSuccess) r = nc->GroupEnd();
    if (r != ncclSuccess) return fail(POTUS_ERR_CUDA, std::string("ncclAllGather: ") + nc->GetErrorString(r));
  }
  CUDA_TRY(cudaSetDevice(s->subs[0]->cfg.device));
  CUDA_TRY(cudaEventRecord(g1));
  for (int d = 0; d < G; ++d) { CUDA_TRY(cudaSetDevice(s->subs[d]->cfg.device)); CUDA_TRY(cudaDeviceSynchronize()); }
  CUDA_TRY(cudaSetDevice(s->subs[0]->cfg.device));
  CUDA_TRY(cudaEventElapsedTime(&gms, g0, g1));
  cudaEventDestroy(g0); cudaEventDestroy(g1);
  PotusStats& st = s->stats;
  st = PotusStats{};
  double acc = 0, dep = 0, eps = 0, tw = 0, ts = 0, tt = 0;
  int C = 0;
  for (PotusSampler* q : s->subs) {
    const PotusStats& a = q->stats;
    const int c = q->cfg.chains;
    st.n_leapfrog_total += a.n_leapfrog_total; st.n_leapfrog_sampling += a.n_leapfrog_sampling; st.n_divergent_sampling += a.n_divergent_sampling;
    st.gpu_launches += a.gpu_launches;
    acc += a.mean_accept_stat * c; dep += a.mean_treedepth * c; eps += a.mean_stepsize * c; C += c;
    tw = std::max(tw, a.seconds_warmup); ts = std::max(ts, a.seconds_sampling); tt = std::max(tt, a.seconds_total);
  }
  st.mean_accept_stat = acc / C; st.mean_treedepth = dep / C; st.mean_stepsize = eps / C;
  st.seconds_warmup = tw; st.seconds_sampling = ts; st.seconds_total = tt + gms * 1e-3;   // max over devices + the gather
  st.seconds_gather = gms * 1e-3;
  st.n_params = s->D; st.n_draws_kept = C * s->keep;
  s->ran = true; s->have_host = false;
  return POTUS_OK;
}

// potus_run in three steps, so that a host with an event loop (R: R_CheckUserInterrupt) is not blocked for the whole run:
// begin enqueues every launch and returns; poll reports whether the device work has finished; end collects the results
// (statistics, and with n_gpus > 1 the ncclAllGather).
int potus_run_begin(PotusSampler* s) {
  if (!s) return fail(POTUS_ERR_STATE, "sampler is NULL");
  int rc;
  if (!s->subs.empty()) {
    for (PotusSampler* q : s->subs) if ((rc = run_launch(q))) return rc;      // all devices run concurrently
    return POTUS_OK;
  }
  return run_launch(s);
}
int potus_run_poll(PotusSampler* s, int* done) {
  if (!s || !done) return fail(POTUS_ERR_STATE, "NULL argument");
  *done = 1;
  std::vector<PotusSampler*> one{s};
  for (PotusSampler* q : (s->subs.empty() ? one : s->subs)) {
    if (!q->ev[2]) return fail(POTUS_ERR_STATE, "potus_run_begin has not been called");
    CUDA_TRY(cudaSetDevice(q->cfg.device));
    const cudaError_t e = cudaEventQuery(q->ev[2]);
    if (e == cudaErrorNotReady) *done = 0;
    else if (e != cudaSuccess) return fail(POTUS_ERR_CUDA, std::string("cudaEventQuery: ") + cudaGetErrorString(e));
  }
  return POTUS_OK;
}
int potus_run_end(PotusSampler* s) {
  if (!s) return fail(POTUS_ERR_STATE, "sampler is NULL");
  if (!s->subs.empty()) return run_multi(s);
  return run_finish(s);
}
int potus_run(PotusSampler* s) {
  const int rc = potus_run_begin(s);
  return rc ? rc : potus_run_end(s);
}

// Start every chain from a given state instead of random inits + warm-up: theta [chains][D] (Stan unconstrained order),
// stepsize [chains], inv_metric [chains][D].  Requires iter_warmup == 0 (adaptation is what produced the state).
int potus_set_state(PotusSampler* s, const double* theta, const double* stepsize, const double* inv_metric) {
  if (!s || !theta || !stepsize || !inv_metric) return fail(POTUS_ERR_STATE, "NULL argument");
  if (!s->subs.empty()) return fail(POTUS_ERR_STATE, "potus_set_state: not available with n_gpus > 1");
  if (s->cfg.iter_warmup != 0) return fail(POTUS_ERR_STATE, "potus_set_state requires iter_warmup == 0");
  CUDA_TRY(cudaSetDevice(s->cfg.device));
  const int C = s->cfg.chains, VL = s->VL, D = s->D;
  std::vector<float> q((size_t)C * VL, 0.f), sm((size_t)C * VL, 0.f);
  std::vector<ChainState> cs(C);
  for (int c = 0; c < C; ++c) {
    for (int k = 0; k < VL; ++k) {
      const int si = (*s->map)[k];
      if (si >= 0) {
        const double im = inv_metric[(size_t)c * D + si];
        if (!(im > 0) || !std::isfinite(theta[(size_t)c * D + si])) return fail(POTUS_ERR_INVALID_DATA, "potus_set_state: inv_metric must be positive and theta finite");
        q[(size_t)c * VL + k] = (float)theta[(size_t)c * D + si];
        sm[(size_t)c * VL + k] = (float)std::sqrt(im);
      }
    }
    if (!(stepsize[c] > 0)) return fail(POTUS_ERR_INVALID_DATA, "potus_set_state: stepsize must be positive");
    ChainState z{};
    z.eps = (float)stepsize[c]; z.da_mu = std::log(10.0 * stepsize[c]); z.status = 0; z.w_next = -1;
    cs[c] = z;
  }
  CUDA_TRY(cudaMemcpy(s->q, q.data(), q.size() * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(s->sqrt_m, sm.data(), sm.size() * sizeof(float), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(s->cs, cs.data(), cs.size() * sizeof(ChainState), cudaMemcpyHostToDevice));
  s->have_state = true;
  return POTUS_OK;
}

// Stan's multi-chain effective sample size from the chain-averaged autocovariance (stan/analyze/mcmc/compute_effective_sample_size.hpp;
// the same steps as us-potus-model_b200/diagnostics.py: Geyer's initial positive + monotone sequence)
static double ess_from_acov(const std::vector<double>& acov_mean, double mean_var, double var_between, int c, int n) {
  double var_plus = mean_var * (n - 1.0) / n;
  if (c > 1) var_plus += var_between;
  if (!(var_plus > 0) || n < 4) return std::nan("");
  std::vector<double> rho(n + 2, 0.0);
  int t = 1;
  double rho_even = 1.0, rho_odd = 1 - (mean_var - acov_mean[1]) / var_plus;
  rho[0] = rho_even; rho[1] = rho_odd;
  while (t < n - 4 && (rho_even + rho_odd) > 0) {
    rho_even = 1 - (mean_var - acov_mean[t + 1]) / var_plus;
    rho_odd = 1 - (mean_var - acov_mean[t + 2]) / var_plus;
    if (rho_even + rho_odd >= 0) { rho[t + 1] = rho_even; rho[t + 2] = rho_odd; }
    t += 2;
  }
  const int max_t = t;
  if (rho_even > 0) rho[max_t + 1] = rho_even;
  t = 1;
  while (t <= max_t - 3) {
    if (rho[t + 1] + rho[t + 2] > rho[t - 1] + rho[t]) { rho[t + 1] = (rho[t - 1] + rho[t]) / 2; rho[t + 2] = rho[t + 1]; }
    t += 2;
  }
  double tau = -1;
  for (int i = 0; i < max_t; ++i) tau += 2 * rho[i];
  tau += rho[max_t + 1];
  tau = std::max(tau, 1.0 / std::log10((double)c * n));
  return (double)c * n / tau;
}

// On-device post-processing over ALL chains x iter_sampling monitored draws (csrc/potus_post.cu).
//   ev            [S] electoral votes (README.Rmd:271-300) or NULL; ev_threshold e.g. 270
//   state_table   [(S+2)][8] row-major: mean, sd, 2.5%, 5%, 50%, 95%, 97.5%, P(> 0.5)  of inv_logit(mu_b[s,T]) for the S states,
//                 row S = national vote (state_weights-weighted mean of the shares per draw), row S+1 = democratic electoral votes
//                 (its last column: P(ev >= ev_threshold); zeros without ev)
//   ess_table     [(S+1)][3] row-major: Stan ESS, split R-hat, posterior mean of the monitored scalars (logit scale; row S =
//                 national_mu_b_average[T]); may be NULL
int potus_postprocess(PotusSampler* s, const double* ev, double ev_threshold, double* state_table, double* ess_table) {
  if (!s || !state_table) return fail(POTUS_ERR_STATE, "NULL argument");
  if (!s->ran) return fail(POTUS_ERR_STATE, "potus_run has not completed");
  if (!s->subs.empty()) return fail(POTUS_ERR_STATE, "potus_postprocess: not available with n_gpus > 1 (run it per device)");
  const int C = s->cfg.chains, n = s->cfg.iter_sampling, S = s->S, Q = S + 2;
  const long long R = (long long)C * n;
  if (R < 2) return fail(POTUS_ERR_STATE, "potus_postprocess needs at least two monitored draws");
  CUDA_TRY(cudaSetDevice(s->cfg.device));
  std::vector<float> hw(s->h_w), hev(S, 0.f), thr(Q, 0.5f);
  if (ev) for (int i = 0; i < S; ++i) hev[i] = (float)ev[i];
  thr[S + 1] = (float)(ev_threshold - 0.5);
  static const double qs[5] = {0.025, 0.05, 0.5, 0.95, 0.975};
  std::vector<long long> ranks(5);
  std::vector<double> frac(5);
  for (int j = 0; j < 5; ++j) { const double h = (R - 1) * qs[j]; const long long lo = (long long)std::floor(h); ranks[j] = lo + 1; frac[j] = h - lo; }
  const int cpg = 64, G = (C + cpg - 1) / cpg;   // chains per block of the autocovariance kernel (its partial sums are combined on the host)
  float *dsh = nullptr, *dw = nullptr, *dev = nullptr, *dthr = nullptr;
  double *dmom = nullptr, *dsel = nullptr, *dcs = nullptr, *dac = nullptr;
  long long* drank = nullptr;
  auto cleanup = [&]() { cudaFree(dsh); cudaFree(dw); cudaFree(dev); cudaFree(dthr); cudaFree(dmom); cudaFree(dsel); cudaFree(dcs); cudaFree(dac); cudaFree(drank); };
  cudaError_t e;
#define PP_TRY(x) do { e = (x); if (e != cudaSuccess) { cleanup(); return fail(POTUS_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e)); } } while (0)
  PP_TRY(cudaMalloc(&dsh, (size_t)Q * R * sizeof(float)));
  PP_TRY(cudaMalloc(&dw, S * sizeof(float))); PP_TRY(cudaMalloc(&dev, S * sizeof(float))); PP_TRY(cudaMalloc(&dthr, Q * sizeof(float)));
  PP_TRY(cudaMalloc(&dmom, (size_t)Q * 3 * sizeof(double))); PP_TRY(cudaMalloc(&dsel, (size_t)Q * 5 * 2 * sizeof(double)));
  PP_TRY(cudaMalloc(&drank, 5 * sizeof(long long)));
  PP_TRY(cudaMemcpy(dw, hw.data(), S * sizeof(float), cudaMemcpyHostToDevice));
  PP_TRY(cudaMemcpy(dev, hev.data(), S * sizeof(float), cudaMemcpyHostToDevice));
  PP_TRY(cudaMemcpy(dthr, thr.data(), Q * sizeof(float), cudaMemcpyHostToDevice));
  PP_TRY(cudaMemcpy(drank, ranks.data(), 5 * sizeof(long long), cudaMemcpyHostToDevice));
  potus_post_shares_kernel<<<1184, 256>>>(s->monitor, dw, ev ? dev : nullptr, S, R, dsh);
  potus_post_moments_kernel<<<Q, PNT>>>(dsh, R, dthr, dmom);
  potus_post_select_kernel<<<Q * 5, PNT>>>(dsh, R, drank, 5, dsel);
  std::vector<double> mom((size_t)Q * 3), sel((size_t)Q * 10);
  PP_TRY(cudaMemcpy(mom.data(), dmom, mom.size() * sizeof(double), cudaMemcpyDeviceToHost));
  PP_TRY(cudaMemcpy(sel.data(), dsel, sel.size() * sizeof(double), cudaMemcpyDeviceToHost));
  for (int q = 0; q < Q; ++q) {
    double* row = state_table + (size_t)q * 8;
    row[0] = mom[q * 3]; row[1] = mom[q * 3 + 1]; row[7] = mom[q * 3 + 2];
    for (int j = 0; j < 5; ++j) { const double a = sel[(q * 5 + j) * 2], b = sel[(q * 5 + j) * 2 + 1]; row[2 + j] = a + frac[j] * (b - a); }
    if (q == S + 1 && !ev) for (int j = 0; j < 8; ++j) row[j] = 0;
  }
  if (ess_table) {
    if (n < 4 || n > 2048) { cleanup(); return fail(POTUS_ERR_UNSUPPORTED, "ESS on the device needs 4 <= iter_sampling <= 2048"); }
    const int S1 = S + 1;
    PP_TRY(cudaMalloc(&dcs, (size_t)S1 * C * 6 * sizeof(double)));
    PP_TRY(cudaMalloc(&dac, (size_t)S1 * G * n * sizeof(double)));
    potus_post_acov_kernel<<<dim3(S1, G), 512, n * sizeof(float)>>>(s->monitor, C, n, S1, cpg, dcs, dac);
    std::vector<double> cs((size_t)S1 * C * 6), ac((size_t)S1 * G * n);
    PP_TRY(cudaMemcpy(cs.data(), dcs, cs.size() * sizeof(double), cudaMemcpyDeviceToHost));
    PP_TRY(cudaMemcpy(ac.data(), dac, ac.size() * sizeof(double), cudaMemcpyDeviceToHost));
    std::vector<double> am(n);
    for (int q = 0; q < S1; ++q) {
      for (int l = 0; l < n; ++l) { double a = 0; for (int g = 0; g < G; ++g) a += ac[((size_t)q * G + g) * n + l]; am[l] = a / C; }
      double mean_var = 0, mm = 0;
      for (int c = 0; c < C; ++c) { const double* o = &cs[((size_t)q * C + c) * 6]; mean_var += o[1] * n / (n - 1.0); mm += o[0]; }
      mean_var /= C; mm /= C;
      double vb = 0;
      for (int c = 0; c < C; ++c) { const double d = cs[((size_t)q * C + c) * 6] - mm; vb += d * d; }
      vb = C > 1 ? vb / (C - 1) : 0;
      // split R-hat over the 2C half chains (first h and last h iterations)
      const int h = n / 2;
      double w = 0, hm = 0;
      for (int c = 0; c < C; ++c) { const double* o = &cs[((size_t)q * C + c) * 6]; w += (o[3] + o[5]) * h / (h - 1.0); hm += o[2] + o[4]; }
      w /= 2 * C; hm /= 2 * C;
      double b = 0;
      for (int c = 0; c < C; ++c) { const double* o = &cs[((size_t)q * C + c) * 6]; b += (o[2] - hm) * (o[2] - hm) + (o[4] - hm) * (o[4] - hm); }
      b = h * b / (2 * C - 1);
      double* row = ess_table + (size_t)q * 3;
      row[0] = ess_from_acov(am, mean_var, vb, C, n);
      row[1] = std::sqrt(((h - 1.0) / h * w + b / h) / w);
      row[2] = mm;
    }
  }
  cleanup();
#undef PP_TRY
  return POTUS_OK;
}

int potus_get_stats(PotusSampler* s, PotusStats* out) {
  if (!s || !out) return fail(POTUS_ERR_STATE, "NULL argument");
  if (!s->ran) return fail(POTUS_ERR_STATE, "potus_run has not completed");
  *out = s->stats;
  return POTUS_OK;
}

int potus_device_buffer(PotusSampler* s, int which, void** dptr, size_t* n) {
  if (!s || !dptr || !n) return fail(POTUS_ERR_STATE, "NULL argument");
  const int C = s->cfg.chains, nt = s->cfg.iter_warmup + s->cfg.iter_sampling;
  if (!s->subs.empty()) {
    if (which != 0) return fail(POTUS_ERR_STATE, "n_gpus > 1: only the all-gathered draw buffer (which = 0) is exposed");
    *dptr = s->gath[0]; *n = s->pad_floats * s->subs.size();
    return POTUS_OK;
  }
  switch (which) {
    case 0: *dptr = s->draws; *n = (size_t)C * s->keep * s->draw_len; break;
    case 1: *dptr = s->monitor; *n = (size_t)C * s->cfg.iter_sampling * (s->S + 1); break;
    case 2: *dptr = s->sparams; *n = (size_t)C * nt * 8; break;
    default: return fail(POTUS_ERR_STATE, "unknown buffer id");
  }
  return POTUS_OK;
}

struct ParInfo { size_t off, len; bool exists; };
static bool par_info(const PotusSampler* s, const char* par, ParInfo& pi) {
  const size_t ST = (size_t)s->S * s->T;
  size_t o = 0;
  const std::string p = par ? par : "";
  auto hit = [&](const char* nm, size_t len, bool ex) { bool h = (p == nm); if (h) { pi.off = o; pi.len = len; pi.exists = ex; } o += len; return h; };
  if (hit("mu_b", ST, true)) return true;
  if (hit("mu_c", s->P, true)) return true;
  if (hit("mu_m", s->M, s->full)) return true;
  if (hit("mu_pop", s->Pop, s->full)) return true;
  if (hit("e_bias", s->T, s->full)) return true;
  if (hit("polling_bias", s->S, true)) return true;
  if (hit("theta", s->D, true)) return true;
  return false;
}

size_t potus_draws_size(const PotusSampler* s, const char* par) {
  if (!s || !par) return 0;
  const int C = s->cfg.chains, nt = s->cfg.iter_warmup + s->cfg.iter_sampling;
  const std::string p = par;
  if (p == "monitor") return (size_t)C * s->cfg.iter_sampling * (s->S + 1);
  if (p == "sampler_params") return (size_t)C * nt * 7;
  if (p == "inv_metric") return (size_t)C * s->D;
  if (p == "predicted_score") return (size_t)C * s->keep * s->S * s->T;
  ParInfo pi;
  if (!par_info(s, par, pi) || !pi.exists) return 0;
  return (size_t)C * s->keep * pi.len;
}

int potus_get_draws(PotusSampler* s, const char* par, double* out, size_t n) {
  if (!s || !par || !out) return fail(POTUS_ERR_STATE, "NULL argument");
  if (!s->ran) return fail(POTUS_ERR_STATE, "potus_run has not completed");
  const size_t need = potus_draws_size(s, par);
  if (need == 0) return fail(POTUS_ERR_STATE, std::string("unknown quantity '") + par + "'");
  if (n < need) return fail(POTUS_ERR_STATE, "output buffer too small");
  const int C = s->cfg.chains, nt = s->cfg.iter_warmup + s->cfg.iter_sampling, S = s->S, T = s->T;
  const std::string p = par;
  if (!s->subs.empty()) {
    if (p == "sampler_params" || p == "monitor" || p == "inv_metric") {
      // per-chain quantities: every shard answers for its own chains; rows are ordered by global chain id
      const size_t K = p == "sampler_params" ? 7 : (p == "monitor" ? (size_t)S + 1 : (size_t)s->D);
      const size_t rows_per_chain = p == "sampler_params" ? (size_t)nt : (p == "monitor" ? (size_t)s->cfg.iter_sampling : 1);
      const size_t R = (size_t)C * rows_per_chain;
      size_t row0 = 0;
      std::vector<double> tmp;
      for (PotusSampler* q : s->subs) {
        const size_t Rd = (size_t)q->cfg.chains * rows_per_chain;
        tmp.resize(Rd * K);
        int rc = potus_get_draws(q, par, tmp.data(), tmp.size());
        if (rc) return rc;
        for (size_t k = 0; k < K; ++k)
          for (size_t r = 0; r < Rd; ++r) out[row0 + r + R * k] = tmp[r + Rd * k];
        row0 += Rd;
      }
      return POTUS_OK;
    }
    if (!s->have_host) {   // kept draws of ALL chains from the first device's all-gathered buffer
      CUDA_TRY(cudaSetDevice(s->subs[0]->cfg.device));
      s->h_draws.resize((size_t)C * s->keep * s->draw_len);
      size_t o = 0;
      for (size_t d = 0; d < s->subs.size(); ++d) {
        const size_t nd = (size_t)s->subs[d]->cfg.chains * s->keep * s->draw_len;
        if (nd) CUDA_TRY(cudaMemcpy(s->h_draws.data() + o, s->gath[0] + d * s->pad_floats, nd * sizeof(float), cudaMemcpyDeviceToHost));
        o += nd;
      }
      s->have_host = true;
    }
  } else {
    CUDA_TRY(cudaSetDevice(s->cfg.device));
  }
  if (p == "sampler_params") {  // [(iter)*chains, 7], row index = chain*nt + it (draw-fastest within a column)
    const size_t R = (size_t)C * nt;
    if (s->h_sparams.size() != R * 8) {
      s->h_sparams.resize(R * 8);
      CUDA_TRY(cudaMemcpy(s->h_sparams.data(), s->sparams, s->h_sparams.size() * sizeof(float), cudaMemcpyDeviceToHost));
    }
    const double c0 = s->lp_const;  // device values are centred: lp__ = -U + c0, energy__ = H - c0
    for (size_t r = 0; r < R; ++r)
      for (int k = 0; k < 7; ++k) {
        double v = s->h_sparams[r * 8 + k];
        if (k == 0) v += c0;
        if (k == 6) v -= c0;
        out[r + R * k] = v;
      }
    return POTUS_OK;
  }
  if (p == "inv_metric") {  // adapted diagonal of M^-1 per chain, [chains, D] chain-fastest, Stan parameter order
    const int VL = s->VL;
    std::vector<float> h((size_t)C * VL);
    CUDA_TRY(cudaMemcpy(h.data(), s->sqrt_m, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (int c = 0; c < C; ++c)
      for (int k = 0; k < VL; ++k) {
        const int si = (*s->map)[k];
        if (si >= 0) { const double r = h[(size_t)c * VL + k]; out[c + (size_t)C * si] = r * r; }
      }
    return POTUS_OK;
  }
  if (p == "monitor") {
    const size_t R = (size_t)C * s->cfg.iter_sampling;
    s->h_monitor.resize(R * (S + 1));
    CUDA_TRY(cudaMemcpy(s->h_monitor.data(), s->monitor, s->h_monitor.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (size_t r = 0; r < R; ++r)
      for (int k = 0; k <= S; ++k) out[r + R * k] = s->h_monitor[r * (S + 1) + k];
    return POTUS_OK;
  }
  if (p == "predicted_score" && s->subs.empty()) {  // [draws, T, S] = inv_logit(mu_b[s,t]); poll_model_2020.stan:136-139
    // formed on the device from the kept-draw records, already in extract()'s layout; fp32 over the bus, widened here
    const size_t R = (size_t)C * s->keep, ne = R * S * T;
    float* dps = nullptr;
    CUDA_TRY(cudaMalloc(&dps, std::max<size_t>(ne * sizeof(float), 16)));
    potus_post_pscore_kernel<<<1184, 256>>>(s->draws, (long long)R, s->draw_len, S, T, dps);
    std::vector<float> hps(ne);
    cudaError_t e = cudaMemcpy(hps.data(), dps, ne * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(dps);
    if (e != cudaSuccess) return fail(POTUS_ERR_CUDA, std::string("potus_post_pscore_kernel: ") + cudaGetErrorString(e));
    for (size_t i = 0; i < ne; ++i) out[i] = hps[i];
    return POTUS_OK;
  }
  const size_t R = (size_t)C * s->keep;
  if (p == "predicted_score") {  // (n_gpus > 1: from the gathered records on the host)
    for (size_t r = 0; r < R; ++r) {
      const float* dr = &s->h_draws[r * s->draw_len];
      for (int t = 0; t < T; ++t)
        for (int st = 0; st < S; ++st) {
          double x = dr[st + (size_t)S * t];
          out[r + R * ((size_t)t + (size_t)T * st)] = 1.0 / (1.0 + std::exp(-x));
        }
    }
    return POTUS_OK;
  }
  ParInfo pi;
  par_info(s, par, pi);
  if (s->subs.empty()) {   // only the requested block of every draw record crosses the bus (strided device-to-host copy)
    std::vector<float> tmp(R * pi.len);
    if (!tmp.empty())
      CUDA_TRY(cudaMemcpy2D(tmp.data(), pi.len * sizeof(float), s->draws + pi.off, (size_t)s->draw_len * sizeof(float), pi.len * sizeof(float), R,
                            cudaMemcpyDeviceToHost));
    for (size_t r = 0; r < R; ++r)
      for (size_t k = 0; k < pi.len; ++k) out[r + R * k] = tmp[r * pi.len + k];  // mu_b: k = s + S*t == R's [draw, s, t] order
    return POTUS_OK;
  }
  for (size_t r = 0; r < R; ++r) {
    const float* dr = &s->h_draws[r * s->draw_len + pi.off];
    for (size_t k = 0; k < pi.len; ++k) out[r + R * k] = dr[k];
  }
  return POTUS_OK;
}

// potus_logp_grad on the streaming kernel (shapes the resident kernel does not hold, or force_stream)
static int logp_grad_stream(const PotusData* data, const double* theta, int n, double* lp, double* grad) {
  StreamHost sh;
  int rc = build_stream_model(data, sh);
  if (rc) { free_stream(sh); return rc; }
  const int D = sh.m.D, VL = sh.m.VL;
  std::vector<float> qin((size_t)n * VL, 0.f);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < VL; ++k) {
      int si = sh.map_i2s[k];
      if (si >= 0) qin[(size_t)i * VL + k] = (float)theta[(size_t)i * D + si];
    }
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  const int grid = std::min(n, nsm);
  float *dq = nullptr, *dg = nullptr, *dr = nullptr;
  double* du = nullptr;
  auto cleanup = [&]() { cudaFree(dq); cudaFree(dg); cudaFree(du); cudaFree(dr); free_stream(sh); };
  cudaError_t e;
  if ((e = cudaMalloc(&dq, qin.size() * 4)) != cudaSuccess || (e = cudaMalloc(&dg, qin.size() * 4)) != cudaSuccess ||
      (e = cudaMalloc(&du, (size_t)n * 8)) != cudaSuccess || (e = cudaMalloc(&dr, (size_t)grid * sh.m.rb_len * 4)) != cudaSuccess || (e = cudaMemset(dr, 0, (size_t)grid * sh.m.rb_len * 4)) != cudaSuccess ||
      (e = cudaMemcpy(dq, qin.data(), qin.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemset(dg, 0, qin.size() * 4)) != cudaSuccess ||
      (e = cudaFuncSetAttribute(potus_stream_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SSMEM_BYTES)) != cudaSuccess) {
    cleanup();
    return fail(POTUS_ERR_CUDA, cudaGetErrorString(e));
  }
  SEvalArgs a{};
  a.m = sh.m; a.n = n; a.q_in = dq; a.g_out = dg; a.u_out = du; a.rbuf = dr;
  potus_stream_eval_kernel<<<grid, SNT, SSMEM_BYTES>>>(a);
  std::vector<float> g((size_t)n * VL);
  std::vector<double> u(n);
  if ((e = cudaDeviceSynchronize()) != cudaSuccess || (e = cudaMemcpy(g.data(), dg, g.size() * 4, cudaMemcpyDeviceToHost)) != cudaSuccess ||
      (e = cudaMemcpy(u.data(), du, (size_t)n * 8, cudaMemcpyDeviceToHost)) != cudaSuccess) {
    cleanup();
    return fail(POTUS_ERR_CUDA, std::string("potus_stream_eval_kernel: ") + cudaGetErrorString(e));
  }
  for (int i = 0; i < n; ++i) {
    lp[i] = -u[i] + sh.m.lp_const;
    for (int k = 0; k < VL; ++k) {
      int si = sh.map_i2s[k];
      if (si >= 0) grad[(size_t)i * D + si] = -(double)g[(size_t)i * VL + k];
    }
  }
  cleanup();
  return POTUS_OK;
}

int potus_logp_grad_ex(const PotusData* data, const double* theta, int n, double* lp, double* grad, int force_stream) {
  if (!data || !theta || !lp || !grad || n < 1) return fail(POTUS_ERR_STATE, "NULL argument");
  int rc = validate(data);
  if (rc) return rc;
  rc = check_device(0, nullptr);
  if (rc) return rc;
  if (force_stream || check_supported(data) != POTUS_OK) return logp_grad_stream(data, theta, n, lp, grad);
  HostModel hm;
  rc = build_model(data, hm);
  if (rc == POTUS_ERR_UNSUPPORTED && check_stream_supported(data) == POTUS_OK) { free_model(hm); return logp_grad_stream(data, theta, n, lp, grad); }
  if (rc) { free_model(hm); return rc; }
  const int D = hm.m.D;
  std::vector<float> qin((size_t)n * VEC, 0.f);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < VEC; ++k) {
      int si = hm.map_i2s[k];
      if (si >= 0) qin[(size_t)i * VEC + k] = (float)theta[(size_t)i * D + si];
    }
  float *dq = nullptr, *dg = nullptr;
  double* du = nullptr;
  auto cleanup = [&]() { cudaFree(dq); cudaFree(dg); cudaFree(du); free_model(hm); };
  cudaError_t e;
  if ((e = cudaMalloc(&dq, qin.size() * 4)) != cudaSuccess || (e = cudaMalloc(&dg, qin.size() * 4)) != cudaSuccess ||
      (e = cudaMalloc(&du, (size_t)n * 8)) != cudaSuccess ||
      (e = cudaMemcpy(dq, qin.data(), qin.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaFuncSetAttribute(potus_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)) != cudaSuccess) {
    cleanup();
    return fail(POTUS_ERR_CUDA, cudaGetErrorString(e));
  }
  EvalArgs a{};
  a.m = hm.m; a.n = n; a.q_in = dq; a.g_out = dg; a.u_out = du; a.mu_out = nullptr;
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  potus_eval_kernel<<<std::min(n, nsm), NT, SMEM_BYTES>>>(a);
  std::vector<float> g((size_t)n * VEC);
  std::vector<double> u(n);
  if ((e = cudaDeviceSynchronize()) != cudaSuccess || (e = cudaMemcpy(g.data(), dg, g.size() * 4, cudaMemcpyDeviceToHost)) != cudaSuccess ||
      (e = cudaMemcpy(u.data(), du, (size_t)n * 8, cudaMemcpyDeviceToHost)) != cudaSuccess) {
    cleanup();
    return fail(POTUS_ERR_CUDA, std::string("potus_eval_kernel: ") + cudaGetErrorString(e));
  }
  for (int i = 0; i < n; ++i) {
    lp[i] = -u[i] + hm.m.lp_const;
    for (int k = 0; k < VEC; ++k) {
      int si = hm.map_i2s[k];
      if (si >= 0) grad[(size_t)i * D + si] = -(double)g[(size_t)i * VEC + k];
    }
  }
  cleanup();
  return POTUS_OK;
}

int potus_logp_grad(const PotusData* data, const double* theta, int n, double* lp, double* grad) {
  return potus_logp_grad_ex(data, theta, n, lp, grad, 0);
}

}  // extern "C"
