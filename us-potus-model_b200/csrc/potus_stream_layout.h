// Streaming ("large-S/T") kernel family: shared host/device definitions.  BASELINE config 5 (S=256, T=365,
// N=50k, D=144 837) and every shape the resident kernel (potus_layout.h) does not hold.  A chain's vectors live
// in global memory (HBM / L2) in the STREAM LAYOUT below; one CTA sweeps a chain's position once per leapfrog in
// 128-day tiles:  W tile -> tcgen05 GEMM (L0 W) -> polls -> G tile -> tcgen05 GEMM (L0^T G) -> prefix over days ->
// momentum / position update, all inside one forward pass over the days.  See DESIGN.md section 8.
#pragma once
#include <cstdint>
#include "potus_layout.h"

namespace potus {

constexpr int SNT = 512;                 // threads per CTA (16 warps)
constexpr int ST_ROWS = 128;             // days per tile = MMA M
constexpr int ST_MAXS = 256;             // states <= 256 (MMA N <= 256)
constexpr int ST_MAXT = 512;             // days <= 512 (AR(1) vectors in shared memory)
constexpr int ST_MAXTILE = ST_MAXT / ST_ROWS;
constexpr int ST_NZS_CAP = 2048;         // "small" non-walk parameters staged in shared memory: zT, zb, c, m, pop, u_mu, u_rho, ze
constexpr int ST_NSTAGE = 3;             // B-operand ring (one MMA K-step of 16 per stage)
constexpr int ST_SEGL = 16;              // residuals per pollster segment (level 1 of the pollster sums)
constexpr int ST_MAXP = 512;              // pollster index: 9 bits of the packed poll word
constexpr int ST_VCHUNK = SNT * 4 * 4;   // vectors are padded to a multiple of this (uniform float4 loops, unrolled by up to 4)

// ---- UMMA operand geometry (K-major, SWIZZLE_NONE, fp16 hi/lo planes).  A: 128 rows x 256 k.
constexpr uint32_t SA_SBO = 128;                    // 8-row groups contiguous
constexpr uint32_t SA_LBO = 16 * 128 + 16;          // 16 row groups per k-group, +16 B so half2 stores of a warp hit distinct banks
constexpr uint32_t SA_PLANE = 32 * SA_LBO;          // 66 048 B
constexpr uint32_t ST_PITCH_PAD = 4;                // fp32 tile pitch = NP + 4 floats (row-per-lane float4 stores conflict-free)
constexpr uint32_t SA_REGION = ST_ROWS * (ST_MAXS + ST_PITCH_PAD) * 4;   // 133 120 B >= 2 planes; fp32 tile [128][NP+4] aliases the planes
static_assert(SA_REGION >= 2 * SA_PLANE, "operand planes must fit the aliased region");
constexpr uint32_t SB_STAGE = ST_MAXS * 16 * 2 * 2; // one K-step: 256 n x 16 k x fp16 x (hi, lo) = 16 384 B

// ---- shared-memory map (byte offsets from a 128-aligned base)
constexpr uint32_t SS_A = 0;
constexpr uint32_t SS_B = SS_A + SA_REGION;
constexpr uint32_t SS_QNZ = SS_B + ST_NSTAGE * SB_STAGE;       // float [ST_NZS_CAP] small parameters of the evaluated point
constexpr uint32_t SS_GNZ = SS_QNZ + ST_NZS_CAP * 4;           // float [ST_NZS_CAP] their data gradients (d lp)
constexpr uint32_t SS_E = SS_GNZ + ST_NZS_CAP * 4;             // float [ST_MAXT] e_bias
constexpr uint32_t SS_GE = SS_E + ST_MAXT * 4;                 // float [ST_MAXT] per-day sums of unadjusted residuals
constexpr uint32_t SS_PRIOR = SS_GE + ST_MAXT * 4;             // float [256] each:
constexpr uint32_t SS_W = SS_PRIOR + 1024;                     //   state weights
constexpr uint32_t SS_LW = SS_W + 1024;                        //   L0^T w
constexpr uint32_t SS_PB = SS_LW + 1024;                       //   polling_bias
constexpr uint32_t SS_BASE = SS_PB + 1024;                     //   a_T zT + a_w colsum(Z)
constexpr uint32_t SS_CARZ = SS_BASE + 1024;                   //   running prefix of Z over days (tiles done)
constexpr uint32_t SS_CARH = SS_CARZ + 1024;                   //   running prefix of H over days
constexpr uint32_t SS_CZN = SS_CARH + 1024;                    //   column sums of the position being written (next sweep's colsum(Z))
constexpr uint32_t SS_QTOT = SS_CZN + 1024;                    // float [4][256] day-quarter totals
constexpr uint32_t SS_NATP = SS_QTOT + 4096;                   // float [4][128] partial national averages
constexpr uint32_t SS_NAT = SS_NATP + 2048;                    // float [128]
constexpr uint32_t SS_RND = SS_NAT + 512;                      // float [128] national residual per day of the tile
constexpr uint32_t SS_RED = SS_RND + 512;                      // double [16][16]
constexpr uint32_t SS_CTL = SS_RED + 16 * 16 * 8;              // control block, 1024 B
constexpr uint32_t SS_MODEL = SS_CTL + 1024;                   // ModelS copy, 1024 B
constexpr uint32_t SS_TOTAL = SS_MODEL + 1024;
static_assert(SS_TOTAL + 128 <= 232448, "shared memory budget (227 KiB) exceeded");

// ---- workspace slots (per CTA, VL floats each)
constexpr int SW_FIRST = 0;                 // FIRST[k], k = 1..10 at SW_FIRST + k - 1
constexpr int SW_LEFT_E = 10;               // Left_k.e, k = 1..9 at SW_LEFT_E + k - 1
constexpr int SW_LEFT_R = 19;               // Left_k.r, k = 2..9 at SW_LEFT_R + k - 1 (k = 1: r = b + e, never stored)
constexpr int SW_TOP_BB = 28, SW_TOP_FF = 29, SW_TOP_RHO = 30;
constexpr int SW_ENDF_Q0 = 31, SW_ENDF_Q1 = 32, SW_ENDF_P = 33, SW_ENDB_Q0 = 34, SW_ENDB_Q1 = 35, SW_ENDB_P = 36;
constexpr int SW_CAND_A = 37, SW_CAND_B = 38;
constexpr int SW_PCUR = 39, SW_SRUN = 40, SW_G = 41, SW_Q = 42, SW_TMPQ = 43, SW_TMPP = 44;
constexpr int SW_NSLOT = 45;

// packed poll word: s[0:9) (s == S: national)  day-in-tile[9:16)  mode[16:18)  pop[18:20)  unadjusted[20]  pollster[21:30)
//   bit 30: first poll of a (state, day) run of state polls (it writes the run's G cell); bit 31: that run is this poll alone
__host__ __device__ inline uint32_t spack_poll(int s, int dloc, int mo, int po, int un, int p, int head, int single) {
  return (uint32_t)s | ((uint32_t)dloc << 9) | ((uint32_t)mo << 16) | ((uint32_t)po << 18) | ((uint32_t)un << 20) | ((uint32_t)p << 21) |
         ((uint32_t)head << 30) | ((uint32_t)single << 31);
}

// STREAM LAYOUT of a D-vector (length VL floats, zero in padding):
//   [0, T*SP)            walk block, day-major: raw_mu_b[s, t] at t*SP + s        (SP = S rounded up to 4)
//   [o_zT ...)           raw_mu_b_T[S] | raw_polling_bias[S] | raw_mu_c[P] | raw_mu_m[M] | raw_mu_pop[Pop] | mu_e_bias | rho_e_bias |
//                        raw_e_bias[T]                                          (the "small" block, nzs floats, staged in shared memory)
//   [o_x, o_x + N)       measurement noise of poll k in SORTED poll order (day, then state, national polls last within a day)
struct ModelS {
  int S, T, P, M, Pop, Nn, Ns, N, full, D;
  int SP, NP, KS, NTILE, VL;
  int o_zT, o_zb, o_c, o_m, o_pop, o_umu, o_urho, o_ze, o_x, nzs;   // o_* are offsets in the vector; small block = [o_zT, o_zT + nzs)
  int n_seg, rp_off, rb_len;   // pollster segments; offset of the pollster-grouped residual copy in a CTA's residual buffer; its length
  float a_b, a_T, a_w, sig_c, sig_m, sig_pop, sig_n, sig_s, sig_e;
  float scale_G, inv_scale_G;
  double lp_const;
  // device pointers (constant for the sampler's lifetime)
  const unsigned char* b1;     // K-step chunks of X = 256 L0 (n = output state, k), triangular-packed, hi then lo plane per chunk
  const unsigned char* b2;     // same for X^T (n = k-state of the backward product)
  const uint32_t* b1_off;      // [KS+1] byte offsets of the chunks
  const uint32_t* b2_off;
  const float* l0t;            // [S][SP] fp32 L0^T (l0t[k][s] = L0[s][k]) for the polling-bias mat-vec
  const float* prior;          // [256]
  const float* w;              // [256]
  const float* lw;             // [256]  L0^T w
  const uint32_t* pw0;         // [N] packed poll word
  const float4* pc;            // [N] (n, eta_hat, p_hat, rho_hat)
  const float* pun;            // [N] unadjusted weight in [0, 1], or null when every value is 0 / 1 (then bit 20 of the poll word says it all)
  const int32_t* tile_ptr;     // [NTILE+1] poll range of each tile
  const int32_t* day_ptr;      // [T+1]
  const uint2* cells;          // (state, day) runs of state polls in poll order: x = first poll, y = s | day-in-tile << 9 | run length << 16
  const int32_t* tile_cptr;    // [NTILE+1] cell range of each tile
  const uint32_t* perm;        // [N] position of sorted poll k in the pollster-grouped residual copy (each pollster padded to ST_SEGL)
  const int32_t* seg_ptr;      // [P+1] segment range of each pollster
  const int32_t* map_i2s;      // [VL] vector slot -> Stan unconstrained index (-1 = padding)
};
static_assert(sizeof(ModelS) <= 1024, "ModelS must fit its shared-memory slot");

struct SRunArgs {
  ModelS m;
  int n_chains, chain_id_offset, iter_begin, iter_end, iter_warmup, iter_sampling, max_depth, do_init;
  int keep_every, keep_per_chain, draw_len;
  int w_init_buffer, w_term_buffer, w_base_window;
  unsigned long long seed;
  float adapt_delta, init_radius;
  float* q; float* sqrt_m; float* wf_mean; float* wf_m2;   // per-chain persistent vectors [n_chains][VL]
  ChainState* cs;
  float* workspace;      // [gridDim.x][SW_NSLOT][VL]
  float* rbuf;           // [gridDim.x][m.rb_len] residuals of the current sweep (sorted order, then pollster-grouped)
  int* queue;
  float* draws;          // [n_chains*keep_per_chain][draw_len]
  float* monitor;        // [n_chains][iter_sampling][S+1]
  float* sampler_params; // [n_chains][iter_warmup+iter_sampling][8]
  unsigned long long* prof; // optional [64] phase cycle counters (POTUS_PROF builds only)
};

struct SEvalArgs {       // test hook: lp/grad for n vectors
  ModelS m;
  int n;
  const float* q_in;     // [n][VL]
  float* g_out;          // [n][VL] gradient of U
  double* u_out;         // [n]
  float* rbuf;           // [gridDim.x][m.rb_len]
};

}  // namespace potus
