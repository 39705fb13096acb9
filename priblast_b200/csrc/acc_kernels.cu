// CUDA kernels (sm_100a) and the C ABI of libpriblast_acc.so.
//
// Kernel set, version 1: every kernel is "one thread = one column of the batch"; the
// per-thread bodies are the PRIB_HD functions of acc_core.h.  The span wavefront is driven from the
// host: one launch per span for the inside pass (ascending) and one per span for the outside pass
// (descending), over ALL sequences of the batch at once, so a launch has (batch nucleotides) threads.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/priblast_acc.h"
#include "acc_core.h"
#include "acc_tables.h"
#include "acc_tile.h"

using namespace prib;
typedef double real;
typedef Core<real> K;
typedef K::Ctx Ctx;
typedef K::SmallTables SmallTables;
typedef Tile<real> TL;

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int kThreads = 128;

__global__ void __launch_bounds__(kThreads) k_inside(Ctx c, int d) {
  const long long g = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (g < c.NC) K::inside_cell(c, g, d);
}

__global__ void __launch_bounds__(kThreads) k_outside(Ctx c, int d) {
  const long long g = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (g < c.NC) K::outside_cell(c, g, d);
}

__global__ void __launch_bounds__(32) k_outer_scans(Ctx c) {
  const int sq = blockIdx.x * 32 + threadIdx.x;
  if (sq >= c.nseq) return;
  double ring[256];
  K::scan_alpha_outer(c, sq, ring);
  K::scan_beta_outer(c, sq, ring);
}

__global__ void __launch_bounds__(kThreads) k_biloop_left(Ctx c) {
  const long long g = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (g < c.NC) K::biloop_left(c, g);
}

__global__ void __launch_bounds__(kThreads) k_biloop_right(Ctx c) {
  const long long g = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (g < c.NC) K::biloop_right(c, g);
}

__global__ void __launch_bounds__(kThreads) k_hairpin_suffix(Ctx c) {
  const long long g = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (g < c.NC) K::hairpin_suffix(c, g);
}

__global__ void __launch_bounds__(kThreads) k_finalize(Ctx c) {
  const long long g = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (g < c.NC) K::finalize_position(c, g);
}



// ---- kernel set v2: tile-persistent span march (acc_tile.h) ------------------------------------
// grid = resident CTAs (one per SM: the rings take ~all shared memory), blockDim = TC, tiles round-robin.
// dynamic smem: kTileRows * TC reals + (TC + 16) base codes.
__global__ void __launch_bounds__(1024, 1) k_inside_tile(Ctx c, int TX, long long ntiles, real *scratch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int TC = blockDim.x, t = threadIdx.x, W = c.W;
  real *base = reinterpret_cast<real *>(smem_raw);
  uint8_t *sS = reinterpret_cast<uint8_t *>(base + (size_t)kTileRows * TC);
  real *scrM1 = scratch + (size_t)blockIdx.x * 2 * (W + 4) * TC;
  real *scrM2 = scrM1 + (size_t)(W + 4) * TC;
  const TL::InSmem sm = TL::carve_in(base, TC, sS);
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    TL::Geo ge;
    ge.g0 = tile * TX;
    ge.TC = TC;
    ge.TX = TX;
    ge.H = W + 1;
    __syncthreads();  // previous tile fully consumed
    for (int r = 0; r < kTileRows; r++) base[(size_t)r * TC + t] = 0;
    for (int k = t; k < TC + 8; k += TC) sS[k] = (ge.g0 + k < c.NC) ? c.S[ge.g0 + k] : 0;
    TL::ColState cs;
    TL::col_state(c, ge.g0 + t, cs);
    __syncthreads();
    for (int d = kTurn; d <= W + 1; d++) {
      TL::inside_span(c, ge, sm, scrM1, scrM2, t, cs, d);
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(1024, 1) k_outside_tile(Ctx c, int TX, long long ntiles, real *scratch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int TC = blockDim.x, t = threadIdx.x, W = c.W;
  real *base = reinterpret_cast<real *>(smem_raw);
  real *scrBif = scratch + (size_t)blockIdx.x * 2 * (W + 4) * TC;
  const TL::OutSmem sm = TL::carve_out(base, TC);
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    TL::Geo ge;
    ge.g0 = tile * TX;
    ge.TC = TC;
    ge.TX = TX;
    ge.H = W + 1;
    __syncthreads();
    for (int r = 0; r < kTileRows; r++) base[(size_t)r * TC + t] = 0;
    TL::ColState cs;
    TL::col_state(c, ge.g0 - ge.H + t, cs);
    __syncthreads();
    int slot = (W + 1) % kRingOut;
    for (int d = W + 1; d >= kTurn; d--) {
      TL::outside_span(c, ge, sm, scrBif, t, cs, d, slot);
      slot = slot == 0 ? kRingOut - 1 : slot - 1;
      __syncthreads();
    }
  }
}

// ---- outer arrays: one warp per sequence (raccess.cpp:230-241, 260-271) ------------------------
// The recurrence is serial in the position but each step is a W-term dot product: lanes split the
// terms, a butterfly adds them, the window of scaled values lives in shared memory.  Values are kept
// linear with an exact power-of-two rescale; logs are taken 32 positions at a time by all lanes.
constexpr int kScanWarps = 4;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <bool ALPHA>
__device__ void warp_scan(const Ctx &c, int sq, double *ring, int lane) {
  const int L = c.seq_len[sq], W = c.W;
  const long long off = c.seq_off[sq];
  const double kBig = 1.3407807929942597e154, kLn2 = 0.6931471805599453094;
  double *dst = ALPHA ? c.lao : c.lbo;
  const real *src = c.arr[ALPHA ? A_STEMDE : A_STEMD];
  long long e2 = 0;
  const int start = ALPHA ? 0 : L;
  if (lane == 0) {
    ring[start & 255] = 1.0;
    dst[off + start] = 0.0;
  }
  __syncwarp();
  double keep_v = 1.0;   // value of the position this lane will take the log of
  long long keep_e = 0;
  int keep_pos = -1;
  for (int step = 1; step <= L; ++step) {
    const int i = ALPHA ? step : L - step;
    const long long col = off + i;
    const int dmax = ALPHA ? imin(W + 1, i) : imin(W + 1, L - i);
    double part = 0;
    for (int d = 5 + lane; d <= dmax; d += 32) {
      const int j = ALPHA ? i - d : i + d;
      part += (double)src[(long long)d * c.NC + col] * ring[j & 255];
    }
    double v = warp_sum(part) + ring[(ALPHA ? i - 1 : i + 1) & 255];
    if (v > kBig) {  // uniform: every lane holds the same v
      const int lo = ALPHA ? imax(0, i - W - 2) : i + 1, hi = ALPHA ? i - 1 : imin(L, i + W + 2);
      for (int k = lo + lane; k <= hi; k += 32) ring[k & 255] *= 1.0 / kBig;
      v *= 1.0 / kBig;
      e2 += 512;
    }
    if (lane == 0) ring[i & 255] = v;
    if ((step & 31) == lane) {
      keep_v = v;
      keep_e = e2;
      keep_pos = i;
    }
    __syncwarp();
    if ((step & 31) == 31 || step == L) {
      if (keep_pos >= 0) dst[off + keep_pos] = log(keep_v) + (double)keep_e * kLn2;
      keep_pos = -1;
    }
  }
}

__global__ void __launch_bounds__(32 * kScanWarps) k_outer_scans_warp(Ctx c) {
  __shared__ double rings[kScanWarps][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sq = blockIdx.x * kScanWarps + warp;
  if (sq >= c.nseq) return;
  warp_scan<true>(c, sq, rings[warp], lane);
  __syncwarp();
  warp_scan<false>(c, sq, rings[warp], lane);
}

// Issue-rate probes for the roofline denominators (MEASURED_PEAKS.json has no SFU / FP32 /
// FP64 figure, so they are measured here): 8 independent dependency chains per thread.
template <int OP>
__global__ void __launch_bounds__(256) k_probe(float *sink_f, double *sink_d, int iters) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (OP == 2) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = 1.0 + 1e-9 * (tid + k);
    const double m = 1.0000001, b = 1e-7;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int k = 0; k < 8; k++) a[k] = fma(a[k], m, b);
    }
    double t = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) t += a[k];
    if (t == 123.456) sink_d[tid] = t;
  } else {
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = 0.5f + 1e-6f * (float)(tid + k);
    const float m = 1.0000001f, b = 1e-7f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
        else a[k] = fmaf(a[k], m, b);
      }
    }
    float t = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) t += a[k];
    if (t == 123.456f) sink_f[tid] = t;
  }
}

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(PRIB_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));            \
  } while (0)

int rows_of(int a, int W) { return (a == X_ML || a == X_MR) ? 32 : W + 4; }

long long state_bytes_per_column(int W) {
  long long r = 0;
  for (int a = 0; a < kNumArr; a++) r += rows_of(a, W);
  return r * (long long)sizeof(real) + 2 * (long long)sizeof(double);
}

struct Batch {
  int n = 0;
  long long NC = 0, nt = 0;
  std::vector<int> ids;         // caller's sequence indices, batch order
  long long out_base = 0;       // float offset of this batch in d_out
  // device copies of the layout
  uint8_t *d_S = nullptr;
  int32_t *d_col_seq = nullptr, *d_seq_len = nullptr;
  long long *d_seq_off = nullptr, *d_acc_off = nullptr, *d_cond_off = nullptr;
};

}  // namespace

struct prib_ctx {
  prib_acc_params prm{};
  int W = 70, delta = 5;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr;
  cudaEvent_t evp[PRIB_NUM_PHASES + 1] = {};
  bool phases_pending = false;
  bool kernel_timed = true;
  // tables
  SmallTables *d_small = nullptr;
  real *d_int11 = nullptr, *d_int21 = nullptr, *d_int22 = nullptr;
  float *d_log = nullptr;
  // tile kernels
  int kernel_set = 2;  // 1 = per-span launches (v1), 2 = tile-persistent (v2)
  int TC = 0, grid_tiles = 0;
  size_t tile_smem = 0;
  real *d_tile_scratch = nullptr;
  // DP scratch
  char *d_state = nullptr;
  long long state_cap_bytes = 0, max_cols = 0;
  // staged work
  std::vector<Batch> batches;
  std::vector<int32_t> lens;
  float *d_out = nullptr;
  long long out_floats = 0;
  bool staged = false, computed = false;
  float *h_stage = nullptr;
  long long h_stage_floats = 0;
  prib_acc_counters cnt{};
};

namespace {

void free_batches(prib_ctx *c) {
  for (auto &b : c->batches) {
    cudaFree(b.d_S);
    cudaFree(b.d_col_seq);
    cudaFree(b.d_seq_len);
    cudaFree(b.d_seq_off);
    cudaFree(b.d_acc_off);
    cudaFree(b.d_cond_off);
  }
  c->batches.clear();
  if (c->d_out) cudaFree(c->d_out);
  c->d_out = nullptr;
  c->out_floats = 0;
  c->staged = c->computed = false;
}

Ctx make_ctx(prib_ctx *c, const Batch &b) {
  Ctx k;
  std::memset(&k, 0, sizeof(k));
  k.NC = b.NC;
  k.W = c->W;
  k.delta = c->delta;
  k.rows = c->W + 4;
  k.nseq = b.n;
  k.S = b.d_S;
  k.col_seq = b.d_col_seq;
  k.seq_len = b.d_seq_len;
  k.seq_off = b.d_seq_off;
  k.T = c->d_small;
  k.e_int11 = c->d_int11;
  k.e_int21 = c->d_int21;
  k.e_int22 = c->d_int22;
  k.log_tbl = c->d_log;
  char *p = c->d_state;
  for (int a = 0; a < kNumArr; a++) {
    k.arr[a] = (real *)p;
    p += (long long)rows_of(a, c->W) * b.NC * (long long)sizeof(real);
  }
  k.lao = (double *)p;
  p += b.NC * (long long)sizeof(double);
  k.lbo = (double *)p;
  k.acc_off = b.d_acc_off;
  k.cond_off = b.d_cond_off;
  k.out = c->d_out + b.out_base;
  return k;
}

int settle_phases(prib_ctx *c) {
  if (!c->phases_pending) return PRIB_OK;
  CU(cudaEventSynchronize(c->evp[PRIB_NUM_PHASES]));
  for (int p = 0; p < PRIB_NUM_PHASES; p++) {
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, c->evp[p], c->evp[p + 1]));
    c->cnt.phase_ms[p] += ms;
  }
  c->phases_pending = false;
  return PRIB_OK;
}

int run_batch(prib_ctx *c, const Batch &b) {
  // phase events of the previous batch must be read before they are re-recorded
  if (settle_phases(c) != PRIB_OK) return PRIB_ECUDA;
  const Ctx k = make_ctx(c, b);
  const long long used = state_bytes_per_column(c->W) * b.NC;
  cudaStream_t st = c->stream;
  if (used > c->cnt.dp_state_bytes_used) c->cnt.dp_state_bytes_used = used;
  CU(cudaEventRecord(c->evp[0], st));
  CU(cudaMemsetAsync(c->d_state, 0, (size_t)used, st));
  const unsigned grid = (unsigned)((b.NC + kThreads - 1) / kThreads);
  CU(cudaEventRecord(c->evp[1], st));
  const int TX = c->TC - (c->W + 1);
  const long long ntiles = (b.NC + TX - 1) / TX;
  const int tgrid = (int)std::min<long long>(ntiles, c->grid_tiles);
  long long launches = 5;
  if (c->kernel_set == 1) {
    for (int d = kTurn; d <= c->W + 1; d++) k_inside<<<grid, kThreads, 0, st>>>(k, d);
    launches += c->W - 1;
  } else {
    k_inside_tile<<<tgrid, c->TC, c->tile_smem, st>>>(k, TX, ntiles, c->d_tile_scratch);
    launches += 1;
  }
  CU(cudaEventRecord(c->evp[2], st));
  if (c->kernel_set == 1) k_outer_scans<<<(b.n + 31) / 32, 32, 0, st>>>(k);
  else k_outer_scans_warp<<<(b.n + kScanWarps - 1) / kScanWarps, 32 * kScanWarps, 0, st>>>(k);
  CU(cudaEventRecord(c->evp[3], st));
  if (c->kernel_set == 1) {
    for (int d = c->W + 1; d >= kTurn; d--) k_outside<<<grid, kThreads, 0, st>>>(k, d);
    launches += c->W - 1;
  } else {
    k_outside_tile<<<tgrid, c->TC, c->tile_smem, st>>>(k, TX, ntiles, c->d_tile_scratch);
    launches += 1;
  }
  CU(cudaEventRecord(c->evp[4], st));
  k_biloop_left<<<grid, kThreads, 0, st>>>(k);
  CU(cudaEventRecord(c->evp[5], st));
  k_biloop_right<<<grid, kThreads, 0, st>>>(k);
  CU(cudaEventRecord(c->evp[6], st));
  k_hairpin_suffix<<<grid, kThreads, 0, st>>>(k);
  k_finalize<<<grid, kThreads, 0, st>>>(k);
  CU(cudaEventRecord(c->evp[7], st));
  CU(cudaGetLastError());
  c->phases_pending = true;
  c->cnt.kernel_launches += launches;
  c->cnt.batches += 1;
  return PRIB_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

const char *prib_last_error(void) { return g_err.c_str(); }
const char *prib_version(void) { return "priblast-b200 0.1 (sm_100a, fp64 linear-domain)"; }

int prib_acc_create(prib_ctx **out, const prib_acc_params *params) {
  if (!out || !params) return fail(PRIB_EINVAL, "null argument");
  *out = nullptr;
  if (params->min_accessible_length <= 1)
    return fail(PRIB_EINVAL, "Error: -d option must be greater than 1");  // raccess.hpp:47-50
  if (params->maximal_span < 1 || params->maximal_span > kMaxSpan)
    return fail(PRIB_EINVAL, "maximal span must be in 1.." + std::to_string((int)kMaxSpan));
  if (params->mode != 0) return fail(PRIB_EINVAL, "only mode 0 (fast) is implemented");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(PRIB_ECUDA, "no CUDA device available (this library has no CPU fallback)");
  if (params->device < 0 || params->device >= ndev) return fail(PRIB_EINVAL, "device ordinal out of range");
  CU(cudaSetDevice(params->device));

  HostTables tab;
  std::string err;
  if (!build_tables(params->maximal_span, tab, err)) return fail(PRIB_EINVAL, err);

  prib_ctx *c = new (std::nothrow) prib_ctx();
  if (!c) return fail(PRIB_ENOMEM, "out of host memory");
  c->prm = *params;
  c->W = params->maximal_span;
  c->delta = params->min_accessible_length;
  auto bail = [&](int code) {
    prib_acc_destroy(c);
    return code;
  };
#define CUB(call)                                                                             \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return bail(fail(PRIB_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)));      \
  } while (0)
  CUB(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  CUB(cudaEventCreate(&c->ev0));
  CUB(cudaEventCreate(&c->ev1));
  CUB(cudaEventCreate(&c->evk0));
  CUB(cudaEventCreate(&c->evk1));
  for (auto &e : c->evp) CUB(cudaEventCreate(&e));
  CUB(cudaMalloc(&c->d_small, sizeof(SmallTables)));
  CUB(cudaMemcpy(c->d_small, &tab.small, sizeof(SmallTables), cudaMemcpyHostToDevice));
  CUB(cudaMalloc(&c->d_int11, tab.e_int11.size() * sizeof(real)));
  CUB(cudaMemcpy(c->d_int11, tab.e_int11.data(), tab.e_int11.size() * sizeof(real), cudaMemcpyHostToDevice));
  CUB(cudaMalloc(&c->d_int21, tab.e_int21.size() * sizeof(real)));
  CUB(cudaMemcpy(c->d_int21, tab.e_int21.data(), tab.e_int21.size() * sizeof(real), cudaMemcpyHostToDevice));
  CUB(cudaMalloc(&c->d_int22, tab.e_int22.size() * sizeof(real)));
  CUB(cudaMemcpy(c->d_int22, tab.e_int22.data(), tab.e_int22.size() * sizeof(real), cudaMemcpyHostToDevice));
  CUB(cudaMalloc(&c->d_log, tab.log_tbl.size() * sizeof(float)));
  CUB(cudaMemcpy(c->d_log, tab.log_tbl.data(), tab.log_tbl.size() * sizeof(float), cudaMemcpyHostToDevice));

  CUB(cudaMemcpyToSymbol(g_conv_d, &tab.small.conv[0][0], sizeof(double) * 32 * 32));
  CUB(cudaMemcpyToSymbol(g_bulge_d, tab.small.e_bulge, sizeof(double) * 32));
  {
    // tile kernels: the widest CTA whose rings fit the opt-in shared memory of this device
    const char *ks = getenv("PRIB_KERNELS");
    c->kernel_set = (ks && ks[0] == '1') ? 1 : 2;
    cudaDeviceProp prop;
    CUB(cudaGetDeviceProperties(&prop, params->device));
    const size_t smem_max = prop.sharedMemPerBlockOptin;
    int TC = (int)((smem_max - 64) / (kTileRows * sizeof(real) + 1)) / 32 * 32;
    if (TC > 1024) TC = 1024;
    const char *tce = getenv("PRIB_TILE_COLS");
    if (tce && atoi(tce) >= c->W + 34 && atoi(tce) <= TC) TC = atoi(tce) / 32 * 32;
    if (TC < c->W + 34) return bail(fail(PRIB_ECUDA, "shared memory too small for the tile kernels"));
    c->TC = TC;
    c->tile_smem = (size_t)kTileRows * TC * sizeof(real) + TC + 16;
    c->grid_tiles = prop.multiProcessorCount;
    CUB(cudaFuncSetAttribute(k_inside_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->tile_smem));
    CUB(cudaFuncSetAttribute(k_outside_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->tile_smem));
    CUB(cudaMalloc(&c->d_tile_scratch, (size_t)c->grid_tiles * 2 * (c->W + 4) * TC * sizeof(real)));
  }
  size_t free_b = 0, total_b = 0;
  CUB(cudaMemGetInfo(&free_b, &total_b));
  long long budget = params->max_batch_bytes > 0 ? params->max_batch_bytes : (long long)(free_b * 0.6);
  if (budget > (long long)(free_b * 0.9)) budget = (long long)(free_b * 0.9);
  const long long per_col = state_bytes_per_column(c->W);
  c->max_cols = budget / per_col / 32 * 32;
  if (c->max_cols < 4096) return bail(fail(PRIB_ECUDA, "device memory budget too small for the DP state"));
  c->state_cap_bytes = c->max_cols * per_col;
  CUB(cudaMalloc(&c->d_state, (size_t)c->state_cap_bytes));
  c->cnt.dp_state_bytes = c->state_cap_bytes;
#undef CUB
  *out = c;
  return PRIB_OK;
}

void prib_acc_destroy(prib_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->prm.device);
  free_batches(c);
  cudaFree(c->d_state);
  cudaFree(c->d_tile_scratch);
  cudaFree(c->d_small);
  cudaFree(c->d_int11);
  cudaFree(c->d_int21);
  cudaFree(c->d_int22);
  cudaFree(c->d_log);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->evk0) cudaEventDestroy(c->evk0);
  if (c->evk1) cudaEventDestroy(c->evk1);
  for (auto &e : c->evp) if (e) cudaEventDestroy(e);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
}

int prib_acc_set_stream(prib_ctx *c, void *cuda_stream) {
  if (!c) return fail(PRIB_EINVAL, "null context");
  c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
  return PRIB_OK;
}

int prib_acc_stage(prib_ctx *c, int32_t n, const char *const *seq, const int32_t *len) {
  if (!c || n < 0 || (n > 0 && (!seq || !len))) return fail(PRIB_EINVAL, "bad argument");
  CU(cudaSetDevice(c->prm.device));
  free_batches(c);
  c->lens.assign(len, len + n);
  for (int k = 0; k < n; k++) {
    if (len[k] < 0) return fail(PRIB_EINVAL, "negative sequence length");
    if (layout_columns(len[k]) + 2 * kPad > c->max_cols)
      return fail(PRIB_ECUDA, "sequence " + std::to_string(k) + " does not fit the device DP budget");
  }
  // longest first (the order of SortSequences, utils.cpp:53-60), then greedy fill of column budgets:
  // neighbours in a batch have similar lengths, which keeps the per-sequence scan warps balanced.
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return len[a] > len[b]; });
  long long out_base = 0;
  size_t pos = 0;
  while (pos < order.size()) {
    Batch b;
    long long cols = 2 * kPad;
    while (pos < order.size() && cols + layout_columns(len[order[pos]]) <= c->max_cols) {
      cols += layout_columns(len[order[pos]]);
      b.ids.push_back(order[pos++]);
    }
    b.n = (int)b.ids.size();
    std::vector<const char *> sp(b.n);
    std::vector<int32_t> sl(b.n);
    std::vector<long long> ao(b.n), co(b.n);
    long long o = 0;
    for (int k = 0; k < b.n; k++) {
      sp[k] = seq[b.ids[k]];
      sl[k] = len[b.ids[k]];
      ao[k] = o;
      co[k] = o + sl[k];
      o += 2LL * sl[k];
      b.nt += sl[k];
    }
    BatchLayout lay;
    build_layout(b.n, sp.data(), sl.data(), lay);
    b.NC = lay.NC;
    b.out_base = out_base;
    out_base += o;
    CU(cudaMalloc(&b.d_S, (size_t)b.NC));
    CU(cudaMalloc(&b.d_col_seq, (size_t)b.NC * sizeof(int32_t)));
    CU(cudaMalloc(&b.d_seq_len, (size_t)std::max(b.n, 1) * sizeof(int32_t)));
    CU(cudaMalloc(&b.d_seq_off, (size_t)std::max(b.n, 1) * sizeof(long long)));
    CU(cudaMalloc(&b.d_acc_off, (size_t)std::max(b.n, 1) * sizeof(long long)));
    CU(cudaMalloc(&b.d_cond_off, (size_t)std::max(b.n, 1) * sizeof(long long)));
    CU(cudaEventRecord(c->ev0, c->stream));
    CU(cudaMemcpyAsync(b.d_S, lay.S.data(), (size_t)b.NC, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(b.d_col_seq, lay.col_seq.data(), (size_t)b.NC * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(b.d_seq_len, sl.data(), (size_t)b.n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(b.d_seq_off, lay.seq_off.data(), (size_t)b.n * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(b.d_acc_off, ao.data(), (size_t)b.n * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(b.d_cond_off, co.data(), (size_t)b.n * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaEventRecord(c->ev1, c->stream));
    CU(cudaStreamSynchronize(c->stream));  // host vectors go out of scope
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->cnt.h2d_ms += ms;
    c->cnt.h2d_bytes += b.NC * 5 + (long long)b.n * 28;
    c->batches.push_back(std::move(b));
  }
  c->out_floats = out_base;
  CU(cudaMalloc(&c->d_out, (size_t)std::max<long long>(out_base, 1) * sizeof(float)));
  c->staged = true;
  return PRIB_OK;
}

int prib_acc_compute(prib_ctx *c) {
  if (!c) return fail(PRIB_EINVAL, "null context");
  if (!c->staged) return fail(PRIB_ESTATE, "prib_acc_compute called before prib_acc_stage");
  CU(cudaSetDevice(c->prm.device));
  // entries the kernels never write (acc tail, cond head) must read 0: raccess.cpp:487-488
  CU(cudaEventRecord(c->evk0, c->stream));
  CU(cudaMemsetAsync(c->d_out, 0, (size_t)std::max<long long>(c->out_floats, 1) * sizeof(float), c->stream));
  for (const Batch &b : c->batches) {
    if (b.n == 0) continue;
    int rc = run_batch(c, b);
    if (rc != PRIB_OK) return rc;
    c->cnt.sequences += b.n;
    c->cnt.nucleotides += b.nt;
  }
  CU(cudaEventRecord(c->evk1, c->stream));
  c->kernel_timed = false;
  c->computed = true;
  return PRIB_OK;
}

static int settle_kernel_time(prib_ctx *c) {
  if (settle_phases(c) != PRIB_OK) return PRIB_ECUDA;
  if (!c->kernel_timed) {
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, c->evk0, c->evk1));
    c->cnt.kernel_ms += ms;
    c->kernel_timed = true;
  }
  return PRIB_OK;
}

int prib_acc_sync(prib_ctx *c) {
  if (!c) return fail(PRIB_EINVAL, "null context");
  CU(cudaSetDevice(c->prm.device));
  CU(cudaStreamSynchronize(c->stream));
  return settle_kernel_time(c);
}

int prib_acc_fetch(prib_ctx *c, float *out, const int64_t *acc_off, const int64_t *cond_off) {
  if (!c || !out || !acc_off || !cond_off) return fail(PRIB_EINVAL, "null argument");
  if (!c->computed) return fail(PRIB_ESTATE, "prib_acc_fetch called before prib_acc_compute");
  CU(cudaSetDevice(c->prm.device));
  if (c->h_stage_floats < c->out_floats) {
    if (c->h_stage) cudaFreeHost(c->h_stage);
    c->h_stage = nullptr;
    CU(cudaMallocHost(&c->h_stage, (size_t)c->out_floats * sizeof(float)));
    c->h_stage_floats = c->out_floats;
  }
  if (c->out_floats > 0) {
    CU(cudaEventRecord(c->ev0, c->stream));
    CU(cudaMemcpyAsync(c->h_stage, c->d_out, (size_t)c->out_floats * sizeof(float), cudaMemcpyDeviceToHost,
                       c->stream));
    CU(cudaEventRecord(c->ev1, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  if (settle_kernel_time(c) != PRIB_OK) return PRIB_ECUDA;
  if (c->out_floats > 0) {
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->cnt.d2h_ms += ms;
    c->cnt.d2h_bytes += c->out_floats * (long long)sizeof(float);
  }
  for (const Batch &b : c->batches) {
    long long o = b.out_base;
    for (int k = 0; k < b.n; k++) {
      const int id = b.ids[k];
      const int L = c->lens[id];
      std::memcpy(out + acc_off[id], c->h_stage + o, sizeof(float) * (size_t)L);
      std::memcpy(out + cond_off[id], c->h_stage + o + L, sizeof(float) * (size_t)L);
      o += 2LL * L;
    }
  }
  return PRIB_OK;
}

int prib_acc_run(prib_ctx *c, int32_t n, const char *const *seq, const int32_t *len, float *out,
                 const int64_t *acc_off, const int64_t *cond_off) {
  int rc = prib_acc_stage(c, n, seq, len);
  if (rc != PRIB_OK) return rc;
  rc = prib_acc_compute(c);
  if (rc != PRIB_OK) return rc;
  return prib_acc_fetch(c, out, acc_off, cond_off);
}

int prib_acc_get_counters(prib_ctx *c, prib_acc_counters *out) {
  if (!c || !out) return fail(PRIB_EINVAL, "null argument");
  *out = c->cnt;
  return PRIB_OK;
}

void *prib_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    g_err = "cudaMallocHost failed";
    return nullptr;
  }
  return p;
}

void prib_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

int prib_peak_probe(int32_t device, double *mufu_gops, double *ffma_gops, double *dfma_gops) {
  if (!mufu_gops || !ffma_gops || !dfma_gops) return fail(PRIB_EINVAL, "null argument");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  float *sf = nullptr;
  double *sd = nullptr;
  CU(cudaMalloc(&sf, (size_t)blocks * threads * sizeof(float)));
  CU(cudaMalloc(&sd, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  double *outs[3] = {mufu_gops, ffma_gops, dfma_gops};
  for (int op = 0; op < 3; op++) {
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
      CU(cudaEventRecord(e0));
      if (op == 0) k_probe<0><<<blocks, threads>>>(sf, sd, iters);
      if (op == 1) k_probe<1><<<blocks, threads>>>(sf, sd, iters);
      if (op == 2) k_probe<2><<<blocks, threads>>>(sf, sd, iters);
      CU(cudaEventRecord(e1));
      CU(cudaEventSynchronize(e1));
      float ms = 0;
      CU(cudaEventElapsedTime(&ms, e0, e1));
      const double gops = (double)blocks * threads * iters * 8.0 / (ms * 1e6);
      if (rep > 0 && gops > best) best = gops;
    }
    *outs[op] = best;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sf);
  cudaFree(sd);
  return PRIB_OK;
}

int64_t prib_acc_record_bytes(int32_t len, int32_t delta) {
  if (len < delta || delta < 1) return PRIB_EINVAL;
  return 8 + 4 * (2LL * len - delta + 1);
}

int64_t prib_acc_write_record(const float *acc, const float *cond, int32_t len, int32_t delta, void *dst) {
  // raccess.cpp:447-481: count, acc[0..count), L, delta zeros, cond[delta..L)
  if (!acc || !cond || !dst || len < delta || delta < 1) return PRIB_EINVAL;
  char *p = (char *)dst;
  const int32_t n1 = len - delta + 1;
  std::memcpy(p, &n1, 4);
  p += 4;
  std::memcpy(p, acc, 4 * (size_t)n1);
  p += 4 * (size_t)n1;
  std::memcpy(p, &len, 4);
  p += 4;
  std::memset(p, 0, 4 * (size_t)delta);
  std::memcpy(p + 4 * (size_t)delta, cond + delta, 4 * (size_t)(len - delta));
  p += 4 * (size_t)len;
  return (int64_t)(p - (char *)dst);
}

}  // extern "C"
