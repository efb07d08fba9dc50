// Shared host/device definitions: the internal ("owner-thread") vector layout, shared-memory map,
// task encodings and the parameter block of the resident NUTS kernel.  See DESIGN.md section 3.
#pragma once
#include <cstdint>

namespace potus {

// ---- CTA geometry: one chain per CTA, 16 warps.  Warp w owns days [16w,16w+16); lane l<26 owns the
// state pair (2l,2l+1); lanes 26..31 own the non-walk ("nz") parameters.  (A 32-warp / 16-element variant
// was measured 17% slower: per-phase fixed latencies -- barriers, TMEM round trips -- dominate, not occupancy.)
constexpr int NT = 512;
constexpr int NWARP = 16;
constexpr int DPW = 16;                 // walk days owned by each warp
constexpr int EPT = 2 * DPW;            // vector elements owned by each thread (two tcgen05.ld/st .x16)
constexpr int VEC = NT * EPT;           // floats per internal vector (16384, 64 KiB)
constexpr int ZLANES = 26;              // state pairs -> S <= 52 incl. the national column
constexpr int NZ_LANES = 6;
constexpr int NZ_CAP = NWARP * NZ_LANES * EPT;  // 3072 non-walk parameters max
constexpr int MAX_S = 51, MAX_T = 254;  // rows 254/255 of the 256-row MMA tile: polling-bias row + spare
constexpr int PB_ROW = 254;
constexpr int NAT_COL = 51;             // GEMM column / K-slot that carries the national series
constexpr int ROWS = 256, KPAD = 64;
constexpr int SCR_PITCH = 54;           // fp32 scratch pitch: even, so a lane's state pair is one 8-byte access, and
                                        // 54 = 22 (mod 32) keeps row-per-lane float2 stores conflict-free per half-warp
constexpr int QZ_PITCH = 52;            // q walk block [t][52] (float2 per state pair)
constexpr int NPOLL_CAP = 1664;
constexpr int SEG = 16;                 // polls per level-1 segment (fully unrolled, predicated)
constexpr int NT1_CAP = 1024;           // level-1 segment-sum tasks: pollster / state / day lists
constexpr int NCELL_CAP = 1440;         // direct (state,day) / national-day cells
constexpr int NT2_CAP = 1024;           // level-2 finals over level-1 partial sums (one per thread)
constexpr int NIDS_CAP = 2 * NPOLL_CAP; // id lists (by pollster, by state)
constexpr int MAX_MODE = 4;             // M, Pop <= 4
constexpr int MAX_DEPTH_CAP = 10;

// ---- UMMA operand geometry (SWIZZLE_NONE, K-major; fp16 hi/lo planes)
constexpr uint32_t A_LBO = 4112;        // 256 rows * 16 B + 16 B pad (bank-conflict-free pair stores)
constexpr uint32_t A_SBO = 128;
constexpr uint32_t A_PLANE = 8 * A_LBO; // 32896 B per hi / lo plane
constexpr uint32_t A_REGION = 2 * A_PLANE;  // 65792 B; doubles as fp32 scratch [255][53] (54060 B)
constexpr uint32_t B_LBO = 1024, B_SBO = 128, B_PLANE = 8192;

// ---- TMEM map (512 columns): accumulators / scratch in [0,256), resident vectors in [256,512).
// A thread-private vector is 32 columns per thread: thread (w,l) -> lane 32*(w%4)+l, columns 32*(w/4)..+31.
constexpr uint32_t TM_D1 = 0, TM_D2 = 128;      // GEMM accumulators (main, cross terms)
constexpr uint32_t TM_TMP = 0;                  // thread-private: full-step momentum P (outside GEMM phases)
constexpr uint32_t TM_G = 128;                  // thread-private: gradient of U / merge running sum
constexpr uint32_t TM_P = 256;                  // thread-private: half-step whitened momentum
constexpr uint32_t TM_S = 384;                  // thread-private: sqrt(inverse metric)

// ---- shared-memory map (byte offsets from a 128-aligned base)
constexpr uint32_t SM_A = 0;
constexpr uint32_t SM_B = SM_A + 65920;                     // A_REGION rounded to 128
constexpr uint32_t SM_QZ = SM_B + 2 * B_PLANE;              // float [256][52]
constexpr uint32_t SM_QNZ = SM_QZ + ROWS * QZ_PITCH * 4;    // float [NZ_CAP]   (walk block has 256 rows; 254/255 stay zero)
constexpr uint32_t SM_GNZ = SM_QNZ + NZ_CAP * 4;            // float [NZ_CAP]   d lp / d theta (data part)
constexpr uint32_t SM_PK = SM_GNZ + NZ_CAP * 4;             // poll data: 5 x [NPOLL_CAP] 32-bit
constexpr uint32_t SM_RR = SM_PK + 5 * NPOLL_CAP * 4;       // float [NPOLL_CAP] residuals
constexpr uint32_t SM_PSUM = SM_RR + NPOLL_CAP * 4;         // float [NT1_CAP]
constexpr uint32_t SM_T1 = SM_PSUM + NT1_CAP * 4;           // uint32 [NT1_CAP] packed level-1 tasks
constexpr uint32_t SM_CELL = SM_T1 + NT1_CAP * 4;           // uint32 [NCELL_CAP] packed cells
constexpr uint32_t SM_IDS = SM_CELL + NCELL_CAP * 4;        // uint16 [NIDS_CAP]
constexpr uint32_t SM_E = SM_IDS + NIDS_CAP * 2;            // float e[256], ebar[256]
constexpr uint32_t SM_TOT = SM_E + 2 * 256 * 4;             // float [NWARP][52]
constexpr uint32_t SM_PRIOR = SM_TOT + NWARP * 52 * 4;      // float [64]
constexpr uint32_t SM_RED = SM_PRIOR + 64 * 4;              // double [NWARP][12]
constexpr uint32_t SM_CTL = SM_RED + (NWARP * 12 + 16) * 8; // control block (scalars), 2048 B
constexpr uint32_t SM_MODEL = SM_CTL + 2048;                // ModelDev copy, 512 B
constexpr uint32_t SM_UTAB = SM_MODEL + 512;                // float [256] per-leaf selection uniforms of the current subtree (one Philox per leaf, not per thread)
constexpr uint32_t SM_TOTAL = SM_UTAB + 1024;
static_assert(SM_TOTAL + 128 <= 232448, "shared memory budget (227 KiB) exceeded");
static_assert(255 * SCR_PITCH * 4 <= A_REGION, "scratch must fit in the operand region");

// ---- workspace slots (per CTA, VEC floats each)
constexpr int SLOT_LEFT0 = 0;                    // (unused)
constexpr int SLOT_LEFT = 1;                     // level k = 1..9 at 1 + 3*(k-1): +0 FIRST[k] (first-leaf momentum, see transition()),
                                                 // +1 Left_k.e, +2 Left_k.r
constexpr int SLOT_TOP_BB = 28, SLOT_TOP_FF = 29, SLOT_TOP_RHO = 30;
constexpr int SLOT_ENDF_Q = 31, SLOT_ENDF_P = 32, SLOT_ENDB_Q = 33, SLOT_ENDB_P = 34;
constexpr int SLOT_CAND_A = 35, SLOT_CAND_B = 36; // sample / proposal positions (roles swap)
constexpr int SLOT_TMPQ = 37;
constexpr int NSLOT = 38;

// nz ownership: warp w owns nz slots [192w, 192w+192); inside, lane ln (0..5) owns the float2 pairs (15-d)*6+ln, d = 0..15,
// so that the six nz lanes of a warp touch 12 consecutive words whenever they walk their elements in step
// (a [lane][32] block layout would put all six on the same banks).  The reversal in d places those 12 words on exactly
// the banks that the 26 walk lanes' 52 words (q walk block, pitch 52) use only once: 12(15-d) = 20d+20 (mod 32), so a
// warp-wide float2 access to "pair d" of every lane is two wavefronts.
__host__ __device__ inline int nz_slot(int w, int ln, int e) {
  return w * (NZ_LANES * EPT) + ((15 - (e >> 1)) * NZ_LANES + ln) * 2 + (e & 1);
}

// owner layout of a D-vector in global memory: element e (0..31) of thread tid lives at ((e/4)*NT + tid)*4 + e%4, i.e.
// every thread moves its elements as eight float4s and a warp's accesses are 512 contiguous bytes.
__host__ __device__ inline int oslot(int e, int tid) { return ((e >> 2) * NT + tid) * 4 + (e & 3); }

// ---- packed poll index: s[0:6) d[6:14) p[14:24) m[24:27) o[27:30) unadj[30]
__host__ __device__ inline uint32_t pack_poll(int s, int d, int p, int m, int o, int un) {
  return (uint32_t)s | ((uint32_t)d << 6) | ((uint32_t)p << 14) | ((uint32_t)m << 24) | ((uint32_t)o << 27) |
         ((uint32_t)un << 30);
}
// level-1 task (one packed word, in shared memory): start[0:13) cnt[13:18) type[18:20) psum-slot[20:32)
//   type 1 contiguous*unadjusted, 2 id list; sorted by decreasing length so that a warp's threads get similar work
// cell (one packed word, shared memory): start[0:12) cnt[12:18) dest[18:32) with dest = t*64+s  -> G operand cell
// level-2 final: word0 = pstart[0:16) pcnt[16:24) kind[24:32); word1 = destination index
//   kind 0: A2[t][s] operand cell (dest = t*64+s, scaled by scale_G); kind 1: gnz[dest] (scaled by sigma_c);
//   kind 2: ebar[dest] (g_e, unscaled); kind 4 / 5: gnz[dest] scaled by sigma_m / sigma_pop (mode / population sums)

struct ModelDev {
  int S, T, P, M, Pop, Nn, Ns, N, full, D, NZ, npair;
  int nz_zT, nz_c, nz_m, nz_pop, nz_umu, nz_urho, nz_ze, nz_x, nz_zb;  // offsets inside the nz block
  int n_t1, n_t2, n_cell, n_ids;
  int urho_owner;           // thread that owns the nz slot of rho_e_bias (-1 if the model has none); see eval_body P1
  float a_b, a_T, a_w, sig_c, sig_m, sig_pop, sig_n, sig_s, sig_e;
  float scale_G, inv_scale_G;
  double lp_const;          // sum_i y_i*eta_hat_i - n_i*softplus(eta_hat_i): the centring constant
  // device pointers (constant for the sampler's lifetime)
  const void* btiles;       // 2 x B_PLANE bytes: X hi, X lo  (X[s][k] = 256*L0[s][k]; row 51 = 256*L0^T w)
  const uint32_t* pk;       // 5 x NPOLL_CAP words: idx, n, eta_hat, p_hat, rho_hat
  const float* prior;       // [64] mu_b_prior, [51] = w . prior
  const uint32_t* t1;       // [NT1_CAP] packed level-1 tasks (copied to shared memory)
  const uint32_t* cells;    // [NCELL_CAP] packed cells (copied to shared memory)
  const uint2* t2;          // [n_t2]
  const uint16_t* ids;      // id lists of the type-2 tasks
  const int32_t* map_i2s;   // [VEC] internal slot -> Stan unconstrained index (-1 = padding)
};

// per-chain adaptation / bookkeeping scalars, persisted in global memory between launches
struct ChainState {
  double da_mu, da_sbar, da_xbar;
  float eps;
  int da_counter;
  int w_counter, w_next, w_size, w_nsamp;
  int iter;            // next iteration index
  int status;          // 0 ok, <0 failed init
  float U, pad;        // potential at the current point (valid after the first transition)
  long long n_leapfrog;
};

struct RunArgs {
  ModelDev m;
  int n_chains, chain_id_offset, iter_begin, iter_end, iter_warmup, iter_sampling, max_depth, do_init;
  int keep_every, keep_per_chain, draw_len;
  int w_init_buffer, w_term_buffer, w_base_window;
  unsigned long long seed;
  float adapt_delta, init_radius;
  // per-chain persistent vectors [n_chains][VEC]
  float* q; float* sqrt_m; float* wf_mean; float* wf_m2;
  ChainState* cs;
  float* workspace;      // [gridDim.x][NSLOT][VEC]
  int* queue;            // work-queue counter
  // outputs
  float* draws;          // [n_chains*keep_per_chain][draw_len]
  float* monitor;        // [n_chains][iter_sampling][S+1]
  float* sampler_params; // [n_chains][iter_warmup+iter_sampling][8]
  unsigned long long* prof; // optional [64] phase cycle counters (POTUS_PROF builds only)
};

struct EvalArgs {      // test hook: lp/grad for n vectors
  ModelDev m;
  int n;
  const float* q_in;   // [n][VEC] internal layout
  float* g_out;        // [n][VEC] gradient of U = -lp(centred), internal layout
  double* u_out;       // [n]
  float* mu_out;       // optional [n][S*T] (mu_b, Stan column-major) or null
};

static_assert(sizeof(ModelDev) <= 512, "ModelDev must fit its shared-memory slot");

}  // namespace potus
