// Host side of the streaming kernel family (included by potus_host.cu): builds the device-side model in the
// stream layout (potus_stream_layout.h) from the same validated PotusData the resident path uses.
#pragma once

namespace {

struct StreamHost {
  ModelS m{};
  Offsets o{};
  int S = 0, T = 0, P = 0, M = 0, Pop = 0, Nn = 0, Ns = 0, N = 0, full = 0;
  std::vector<int32_t> map_i2s;   // [VL]
  std::vector<void*> dev_allocs;
};

template <class Tv>
int upload_to(std::vector<void*>& allocs, const std::vector<Tv>& v, const void** out) {
  void* d = nullptr;
  size_t bytes = std::max<size_t>(v.size() * sizeof(Tv), 16);
  CUDA_TRY(cudaMalloc(&d, bytes));
  allocs.push_back(d);
  if (!v.empty()) CUDA_TRY(cudaMemcpy(d, v.data(), v.size() * sizeof(Tv), cudaMemcpyHostToDevice));
  *out = d;
  return POTUS_OK;
}
void free_stream(StreamHost& sh) {
  for (void* p : sh.dev_allocs) cudaFree(p);
  sh.dev_allocs.clear();
}

// limits of the streaming kernel; everything inside them that the resident kernel refuses is routed here
int check_stream_supported(const PotusData* d) {
  char buf[320];
  const bool full = d->poll_mode_state != nullptr;
  const int nzs = 2 * d->S + d->P + (full ? d->M + d->Pop + 2 + d->T : 0);
  if (d->S > ST_MAXS || d->T > ST_MAXT || d->P > ST_MAXP || nzs > ST_NZS_CAP || (full && (d->M > MAX_MODE || d->Pop > MAX_MODE))) {
    snprintf(buf, sizeof buf,
             "problem size S=%d T=%d P=%d M=%d Pop=%d is outside the streaming kernel's limits (S<=%d, T<=%d, P<=%d, M,Pop<=%d, "
             "2S+P+M+Pop+2+T<=%d)", d->S, d->T, d->P, d->M, d->Pop, ST_MAXS, ST_MAXT, ST_MAXP, MAX_MODE, ST_NZS_CAP);
    return fail(POTUS_ERR_UNSUPPORTED, buf);
  }
  return POTUS_OK;
}

int build_stream_model(const PotusData* d, StreamHost& sh) {
  int rc = check_stream_supported(d);
  if (rc) return rc;
  const int S = d->S, T = d->T, P = d->P, Ns = d->N_state_polls, Nn = d->N_national_polls, N = Ns + Nn;
  const bool full = d->poll_mode_state != nullptr;
  const int M = full ? d->M : 0, Pop = full ? d->Pop : 0;
  sh.S = S; sh.T = T; sh.P = P; sh.M = M; sh.Pop = Pop; sh.Ns = Ns; sh.Nn = Nn; sh.N = N; sh.full = full;
  Offsets& o = sh.o;   // Stan unconstrained order (poll_model_2020.stan:56-69)
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

  // ---- transformed data (poll_model_2020.stan:42-55)
  std::vector<double> L0((size_t)S * S, 0.0), w(d->state_weights, d->state_weights + S);
  double nat = 0;
  for (int i = 0; i < S; ++i)
    for (int j = 0; j < S; ++j) nat += w[i] * d->state_covariance_0[i + S * j] * w[j];
  if (!(nat > 0)) return fail(POTUS_ERR_INVALID_DATA, "state_weights' * state_covariance_0 * state_weights must be positive");
  nat = std::sqrt(nat);
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
  ModelS& m = sh.m;
  m.S = S; m.T = T; m.P = P; m.M = M; m.Pop = Pop; m.Nn = Nn; m.Ns = Ns; m.N = N; m.full = full; m.D = o.D;
  m.SP = (S + 3) & ~3; m.NP = (S + 63) & ~63; m.KS = (S + 15) / 16; m.NTILE = (T + ST_ROWS - 1) / ST_ROWS;
  m.a_b = (float)(d->polling_bias_scale / nat); m.a_T = (float)(d->mu_b_T_scale / nat); m.a_w = (float)(d->random_walk_scale / nat);
  m.sig_c = (float)d->sigma_c; m.sig_m = (float)d->sigma_m; m.sig_pop = (float)d->sigma_pop;
  m.sig_n = (float)d->sigma_measure_noise_national; m.sig_s = (float)d->sigma_measure_noise_state; m.sig_e = (float)d->sigma_e_bias;
  // stream layout
  int v = T * m.SP;
  m.o_zT = v; v += S;
  m.o_zb = v; v += S;
  m.o_c = v; v += P;
  m.o_m = m.o_pop = m.o_umu = m.o_urho = m.o_ze = m.o_zT;
  if (full) { m.o_m = v; v += M; m.o_pop = v; v += Pop; m.o_umu = v; v += 1; m.o_urho = v; v += 1; m.o_ze = v; v += T; }
  m.nzs = v - m.o_zT;
  m.o_x = (v + 3) & ~3;
  m.VL = ((m.o_x + N + ST_VCHUNK - 1) / ST_VCHUNK) * ST_VCHUNK;

  // ---- polls sorted by (day, state); the national polls of a day (s = S) come last within the day
  struct HP { int s, dd, p, mo, po, un; double n, y; int stan_x; double unf; };   // un: unadjusted != 0; unf: its value in [0, 1]
  std::vector<HP> hp(N);
  for (int i = 0; i < Ns; ++i)
    hp[i] = HP{d->state[i] - 1, d->day_state[i] - 1, d->poll_state[i] - 1, full ? d->poll_mode_state[i] - 1 : 0, full ? d->poll_pop_state[i] - 1 : 0,
               full ? (d->unadjusted_state[i] != 0.0 ? 1 : 0) : 0, (double)d->n_two_share_state[i], (double)d->n_democrat_state[i], o.xs + i,
               full ? d->unadjusted_state[i] : 0.0};
  for (int j = 0; j < Nn; ++j)
    hp[Ns + j] = HP{S, d->day_national[j] - 1, d->poll_national[j] - 1, full ? d->poll_mode_national[j] - 1 : 0, full ? d->poll_pop_national[j] - 1 : 0,
                    full ? (d->unadjusted_national[j] != 0.0 ? 1 : 0) : 0, (double)d->n_two_share_national[j], (double)d->n_democrat_national[j], o.xn + j,
                    full ? d->unadjusted_national[j] : 0.0};
  std::stable_sort(hp.begin(), hp.end(), [](const HP& a, const HP& b) { return a.dd != b.dd ? a.dd < b.dd : a.s < b.s; });
  std::vector<uint32_t> pw0(N);
  std::vector<float4> pc(N);
  std::vector<float> pun;   // fractional `unadjusted_*` (poll_model_2020.stan:22-23 allows [0, 1]; the reference's lists hold 0 / 1): one weight per poll
  for (int k = 0; k < N; ++k) if (hp[k].unf != 0.0 && hp[k].unf != 1.0) { pun.assign(N, 0.f); break; }
  if (!pun.empty()) for (int k = 0; k < N; ++k) pun[k] = (float)hp[k].unf;
  std::vector<int32_t> tile_ptr(m.NTILE + 1, 0), day_ptr(T + 1, 0);
  double lp_const = 0, cell_max = 1.0, cell_n = 0;
  for (int k = 0; k < N; ++k) {
    const HP& q = hp[k];
    double frac = q.n > 0 ? q.y / q.n : 0.5;
    double fc = std::min(std::max(frac, 1e-4), 1.0 - 1e-4);
    float eh = (float)std::log(fc / (1.0 - fc));
    float ph = (float)(1.0 / (1.0 + std::exp(-(double)eh)));
    float rh = (float)(frac - (double)ph);
    const bool head = q.s < S && !(k > 0 && hp[k - 1].dd == q.dd && hp[k - 1].s == q.s);
    const bool more = k + 1 < N && hp[k + 1].dd == q.dd && hp[k + 1].s == q.s;
    pw0[k] = spack_poll(q.s, q.dd % ST_ROWS, q.mo, q.po, q.un, q.p, head ? 1 : 0, (head && !more) ? 1 : 0);
    pc[k] = make_float4((float)q.n, eh, ph, rh);
    double ehd = eh, sp = (ehd > 0 ? ehd : 0) + std::log1p(std::exp(-std::fabs(ehd)));
    lp_const += q.y * ehd - q.n * sp;
    day_ptr[q.dd + 1]++;
    tile_ptr[q.dd / ST_ROWS + 1]++;
    if (k > 0 && hp[k - 1].dd == q.dd && hp[k - 1].s == q.s) cell_n += q.n; else cell_n = q.n;
    if (q.s < S) cell_max = std::max(cell_max, cell_n);
  }
  // (state, day) runs of state polls: the G operand cells, one descriptor each
  std::vector<uint2> cells;
  std::vector<int32_t> tile_cptr(m.NTILE + 1, 0);
  for (int k = 0; k < N;) {
    int e = k + 1;
    while (e < N && hp[e].dd == hp[k].dd && hp[e].s == hp[k].s) ++e;
    if (hp[k].s < S) {
      if (e - k > 65535) return fail(POTUS_ERR_UNSUPPORTED, "more than 65535 polls in one (state, day) cell");
      cells.push_back(make_uint2((uint32_t)k, (uint32_t)hp[k].s | ((uint32_t)(hp[k].dd % ST_ROWS) << 9) | ((uint32_t)(e - k) << 16)));
      tile_cptr[hp[k].dd / ST_ROWS + 1]++;
    }
    k = e;
  }
  for (int t = 0; t < m.NTILE; ++t) tile_cptr[t + 1] += tile_cptr[t];
  for (int t = 0; t < T; ++t) day_ptr[t + 1] += day_ptr[t];
  for (int t = 0; t < m.NTILE; ++t) tile_ptr[t + 1] += tile_ptr[t];
  m.lp_const = lp_const;
  {  // G operand scale: |sum of residuals of a cell| <= sum of its n must stay inside fp16 range
    double sc = 1.0;
    while (cell_max * sc > 32768.0) sc *= 0.5;
    m.scale_G = (float)sc; m.inv_scale_G = (float)(1.0 / sc);
  }

  // ---- X = 256 L0 as K-step chunks (16 k each), triangular-packed; B1: rows n = output state >= 16 kc; B2 (X^T): rows n <= 16 kc + 15
  auto pack_chunks = [&](bool transposed, std::vector<unsigned char>& out, std::vector<uint32_t>& offs) {
    offs.assign(m.KS + 1, 0);
    for (int kc = 0; kc < m.KS; ++kc) {
      const int n0 = transposed ? 0 : 16 * kc, n1 = transposed ? std::min(m.NP, 16 * (kc + 1)) : m.NP, nloc = n1 - n0;
      const size_t plane = (size_t)nloc * 32, base = out.size();
      out.resize(base + 2 * plane, 0);
      for (int nl = 0; nl < nloc; ++nl)
        for (int kk = 0; kk < 16; ++kk) {
          const int n = n0 + nl, k = 16 * kc + kk;
          double val = 0;
          if (n < S && k < S) val = transposed ? L0[(size_t)k * S + n] : L0[(size_t)n * S + k];
          __half hi, lo;
          split_half((float)(val * 256.0), hi, lo);
          const size_t e = (size_t)(kk / 8) * ((size_t)nloc * 16) + (size_t)(nl / 8) * 128 + (size_t)(nl % 8) * 16 + (size_t)(kk % 8) * 2;
          memcpy(&out[base + e], &hi, 2);
          memcpy(&out[base + plane + e], &lo, 2);
        }
      offs[kc + 1] = (uint32_t)out.size();
    }
  };
  std::vector<unsigned char> b1, b2;
  std::vector<uint32_t> b1_off, b2_off;
  pack_chunks(false, b1, b1_off);
  pack_chunks(true, b2, b2_off);
  std::vector<float> l0t((size_t)S * m.SP, 0.f), prior(256, 0.f), wv(256, 0.f), lw(256, 0.f);
  for (int s = 0; s < S; ++s) {
    prior[s] = (float)d->mu_b_prior[s]; wv[s] = (float)w[s];
    double a = 0;
    for (int r = s; r < S; ++r) a += w[r] * L0[(size_t)r * S + s];
    lw[s] = (float)a;
    for (int k = 0; k <= s; ++k) l0t[(size_t)k * m.SP + s] = (float)L0[(size_t)s * S + k];
  }

  // ---- pollster sums: a second copy of the residuals grouped by pollster (each pollster padded to ST_SEGL entries, pads stay 0)
  std::vector<std::vector<uint32_t>> byp(P);
  for (int k = 0; k < N; ++k) byp[hp[k].p].push_back((uint32_t)k);
  std::vector<uint32_t> perm(N, 0);
  std::vector<int32_t> seg_ptr(P + 1, 0);
  {
    uint32_t pos = 0;
    for (int p = 0; p < P; ++p) {
      for (size_t a = 0; a < byp[p].size(); ++a) perm[byp[p][a]] = pos + (uint32_t)a;
      const uint32_t nseg = (uint32_t)((byp[p].size() + ST_SEGL - 1) / ST_SEGL);
      pos += nseg * ST_SEGL;
      seg_ptr[p + 1] = seg_ptr[p] + (int32_t)nseg;
    }
    m.n_seg = seg_ptr[P];
    m.rp_off = (N + 3) & ~3;
    m.rb_len = m.rp_off + (int)pos + 4;
    if ((size_t)m.n_seg * 4 > SA_REGION) return fail(POTUS_ERR_UNSUPPORTED, "too many pollster segments for the shared-memory partial-sum table");
  }

  // ---- vector slot -> Stan unconstrained index
  sh.map_i2s.assign(m.VL, -1);
  for (int t = 0; t < T; ++t)
    for (int s = 0; s < S; ++s) sh.map_i2s[(size_t)t * m.SP + s] = o.Z + s + S * t;
  for (int s = 0; s < S; ++s) { sh.map_i2s[m.o_zT + s] = o.zT + s; sh.map_i2s[m.o_zb + s] = o.zb + s; }
  for (int p = 0; p < P; ++p) sh.map_i2s[m.o_c + p] = o.c + p;
  if (full) {
    for (int j = 0; j < M; ++j) sh.map_i2s[m.o_m + j] = o.m + j;
    for (int j = 0; j < Pop; ++j) sh.map_i2s[m.o_pop + j] = o.pop + j;
    sh.map_i2s[m.o_umu] = o.umu; sh.map_i2s[m.o_urho] = o.urho;
    for (int t = 0; t < T; ++t) sh.map_i2s[m.o_ze + t] = o.ze + t;
  }
  for (int k = 0; k < N; ++k) sh.map_i2s[m.o_x + k] = hp[k].stan_x;
  {
    std::vector<int> cnt(o.D, 0);
    for (int vv : sh.map_i2s) if (vv >= 0) { if (vv >= o.D) return fail(POTUS_ERR_STATE, "stream map out of range"); cnt[vv]++; }
    for (int i = 0; i < o.D; ++i) if (cnt[i] != 1) return fail(POTUS_ERR_STATE, "stream layout map is not a bijection");
  }
  const void* p;
  auto& al = sh.dev_allocs;
  if ((rc = upload_to(al, b1, &p))) return rc; m.b1 = (const unsigned char*)p;
  if ((rc = upload_to(al, b2, &p))) return rc; m.b2 = (const unsigned char*)p;
  if ((rc = upload_to(al, b1_off, &p))) return rc; m.b1_off = (const uint32_t*)p;
  if ((rc = upload_to(al, b2_off, &p))) return rc; m.b2_off = (const uint32_t*)p;
  if ((rc = upload_to(al, l0t, &p))) return rc; m.l0t = (const float*)p;
  if ((rc = upload_to(al, prior, &p))) return rc; m.prior = (const float*)p;
  if ((rc = upload_to(al, wv, &p))) return rc; m.w = (const float*)p;
  if ((rc = upload_to(al, lw, &p))) return rc; m.lw = (const float*)p;
  if ((rc = upload_to(al, pw0, &p))) return rc; m.pw0 = (const uint32_t*)p;
  if ((rc = upload_to(al, pc, &p))) return rc; m.pc = (const float4*)p;
  m.pun = nullptr;
  if (!pun.empty()) { if ((rc = upload_to(al, pun, &p))) return rc; m.pun = (const float*)p; }
  if ((rc = upload_to(al, tile_ptr, &p))) return rc; m.tile_ptr = (const int32_t*)p;
  if ((rc = upload_to(al, day_ptr, &p))) return rc; m.day_ptr = (const int32_t*)p;
  if ((rc = upload_to(al, cells, &p))) return rc; m.cells = (const uint2*)p;
  if ((rc = upload_to(al, tile_cptr, &p))) return rc; m.tile_cptr = (const int32_t*)p;
  if ((rc = upload_to(al, perm, &p))) return rc; m.perm = (const uint32_t*)p;
  if ((rc = upload_to(al, seg_ptr, &p))) return rc; m.seg_ptr = (const int32_t*)p;
  if ((rc = upload_to(al, sh.map_i2s, &p))) return rc; m.map_i2s = (const int32_t*)p;
  return POTUS_OK;
}

constexpr int SSMEM_BYTES = (int)SS_TOTAL + 128;

}  // namespace
