// Tensor-core probe for the generic interior-loop stencil (VERDICT r1 "next" #10; output in profiles/r2/tc_probe.txt).
//
// The stencil of one finished source row x (span-major band row, one value per column) onto its 31 future target
// rows is    D[i][n] = sum_u  W[n][u] * x[i + u],   u = 0..31,  n = target slot 0..31,   W[n][u] = conv[u][n - u].
// As a tcgen05.mma that is  D (128 columns x 32 targets, FP32 in TMEM) = A (128 x 32 Hankel) * B (32 x 32 weights).
// What the probe establishes:
//   1. the Hankel operand needs NO im2col: a K-major, no-swizzle shared-memory descriptor whose 8-row groups are
//      128 B apart and whose two 16-byte K chunks are 16 B apart addresses A[m][k] = x[4m + k] straight out of the
//      band row (overlapping "core matrices": a row of a core matrix is 16 B = 4 columns further on).  Lane m of
//      the accumulator is then column 4m + c; the residue c = 0..3 comes from four copies of the row shifted by one
//      element (a descriptor's start address has 16-byte granularity), each with its own accumulator block, and
//      the shifts 8, 16, 24 from the start address.  (An MN-major operand would need no residue blocks, but the
//      MN-major no-swizzle encoding tried first produced zeros: see profiles/r2/tc_probe.txt.)
//   2. FP32-grade results with the 3-product TF32 split (x_hi w_hi + x_lo w_hi + x_hi w_lo): error against FP64
//   3. the issue rate of the 48 MMAs (M = 128, N = 32, K = 8, kind::tf32) a source row costs per 512 columns
// One CTA of 128 threads per SM; every wait is bounded (a stuck barrier sets an error code, it cannot hang the box).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tc_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CU(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      std::exit(2);                                                                \
    }                                                                              \
  } while (0)

constexpr int kM = 128;      // columns per MMA (TMEM lanes)
constexpr int kU = 32;       // shifts (K of the whole product)
constexpr int kCols = 512;    // columns of one accumulator set: 128 lanes x 4 residues
constexpr int kXLen = 560;    // floats of a staged row copy: 4 * 127 + 31 + 3 read, rounded up

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (sm_100 layout: start >> 4 at [0,14), LBO >> 4 at [16,30), SBO >> 4 at [32,46),
// version 1 at [46,48), layout type 0 = no swizzle at [61,64))
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool bar_wait_bounded(uint32_t bar, uint32_t parity, long long max_clk) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return true;
    if (clock64() - t0 > max_clk) return false;
  }
}

struct Params {
  const float *x;        // kXLen + 4 floats of one source row
  const float *bmat;     // [4 shift blocks k0][2 parts (hi, lo)][N * 8] in the K-major core-matrix order
  float *out;            // [2 variants][512 columns][N]: plain TF32, 3-product split
  long long *cycles;     // per CTA: clocks of the timed MMA stream
  int *err;
  int iters;             // timed groups of 48 MMAs (one source row onto N targets, 512 columns)
  int hankel;            // 1 = operand addressed in place (overlapping core matrices), 0 = im2col'd copy
};

// operands of one MMA (K-major, no swizzle): element (row, k) at (k / 4) * LBO + (row / 8) * SBO + (row % 8) * 16 B
// + (k % 4) * 4 B.   B (N x 8): LBO = N * 16 B, SBO = 128 B.   A im2col'd (128 x 8): LBO = 2048 B, SBO = 128 B.
// A in place: the same with LBO = 16 B -> byte offset 16 m + 4 k = element 4 m + k of the row.
template <int N>
__global__ void __launch_bounds__(128, 1) k_tc_probe(Params p) {
  extern __shared__ __align__(128) uint8_t smraw[];
  float *xs = reinterpret_cast<float *>(smraw);  // [2 parts][4 residues][kXLen]
  float *bs = xs + 2 * 4 * kXLen;                // [4 k0][2 parts][N * 8]
  float *im = bs + 4 * 2 * N * 8;                // [2 parts][4 residues][4 k0][1024]
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5;
  // stage the row: hi = the 19 bits the tensor core reads, lo = the exact remainder; copy c is shifted by c elements
  for (int k = t; k < 4 * kXLen; k += 128) {
    const int c = k / kXLen, q = k % kXLen;
    const float v = p.x[q + c];
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    xs[(0 * 4 + c) * kXLen + q] = hi;
    xs[(1 * 4 + c) * kXLen + q] = v - hi;
  }
  for (int k = t; k < 4 * 2 * N * 8; k += 128) bs[k] = p.bmat[k];
  for (int k = t; k < 2 * 4 * 4 * 1024; k += 128) {
    const int blk = k / 1024, kk = (k / 128) % 8, m = k % 128;  // blk = (part * 4 + c) * 4 + k0 / 8
    const int k0 = (blk % 4) * 8, pc = blk / 4;
    im[blk * 1024 + (kk / 4) * 512 + (m / 8) * 32 + (m % 8) * 4 + (kk % 4)] = xs[pc * kXLen + 4 * m + k0 + kk];
  }
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the MMA unit
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(4 * N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = make_idesc(kM, N);
  const uint32_t barA = smem_u32(&bar);
  auto adesc = [&](int part, int c, int kb) {
    return p.hankel ? make_desc(smem_u32(xs + (part * 4 + c) * kXLen + 8 * kb), 16, 128)
                    : make_desc(smem_u32(im + ((part * 4 + c) * 4 + kb) * 1024), 2048, 128);
  };
  auto bdesc = [&](int part, int kb) { return make_desc(smem_u32(bs + (kb * 2 + part) * N * 8), N * 16, 128); };
  uint32_t parity = 0;
  bool ok = true;
  // ---- numerics: variant 0 = x_hi w_hi only (what plain TF32 computes), variant 1 = 3-product split ----
  for (int variant = 0; variant < 2; ++variant) {
    if (t == 0) {
      for (int c = 0; c < 4; ++c) {
        uint32_t acc = 0;
        for (int kb = 0; kb < 4; ++kb) {
          mma_tf32(tmem + c * N, adesc(0, c, kb), bdesc(0, kb), idesc, acc);
          acc = 1;
          if (variant == 1) {
            mma_tf32(tmem + c * N, adesc(1, c, kb), bdesc(0, kb), idesc, acc);  // x_lo w_hi
            mma_tf32(tmem + c * N, adesc(0, c, kb), bdesc(1, kb), idesc, acc);  // x_hi w_lo
          }
        }
      }
      mma_commit(barA);
    }
    ok = ok && bar_wait_bounded(barA, parity, 200000000LL);
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (ok) {
      for (int n0 = 0; n0 < 4 * N; n0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + n0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (blockIdx.x == 0) {
          const int c = n0 / N, col = 4 * t + c;  // lane t of residue block c is column 4 t + c
          for (int j = 0; j < 32; ++j) p.out[((size_t)variant * kCols + col) * N + (n0 % N) + j] = __uint_as_float(r[j]);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  // ---- issue rate: iters groups of the 48 MMAs of one source row, one commit at the end ----
  long long t0 = 0, t1 = 0;
  if (ok) {
    if (t == 0) {
      t0 = clock64();
      for (int it = 0; it < p.iters; ++it) {
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            const uint64_t ah = adesc(0, c, kb), al = adesc(1, c, kb), bh = bdesc(0, kb), bl = bdesc(1, kb);
            mma_tf32(tmem + c * N, ah, bh, idesc, 1);
            mma_tf32(tmem + c * N, al, bh, idesc, 1);
            mma_tf32(tmem + c * N, ah, bl, idesc, 1);
          }
        }
      }
      mma_commit(barA);
    }
    ok = bar_wait_bounded(barA, parity, 4000000000LL);
    t1 = clock64();
    parity ^= 1;
  }
  if (t == 0) {
    p.cycles[blockIdx.x] = t1 - t0;
    if (!ok) atomicExch(p.err, 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(4 * N));
}

static double clk_ghz() {
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  return khz * 1e-6;
}

template <int N>
static void run(int hankel, int iters, const std::vector<float> &x, const std::vector<double> &W /* [N][32] */) {
  // split the weights like the kernel splits x; order them into the descriptor's layout
  std::vector<float> bmat((size_t)4 * 2 * N * 8, 0.f);
  for (int kb = 0; kb < 4; ++kb)
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < 8; ++k) {
        const float w = (float)W[(size_t)n * kU + 8 * kb + k];
        uint32_t bits;
        std::memcpy(&bits, &w, 4);
        bits &= 0xFFFFE000u;
        float hi;
        std::memcpy(&hi, &bits, 4);
        const size_t off = (size_t)(k / 4) * N * 4 + (size_t)(n / 8) * 32 + (n % 8) * 4 + (k % 4);
        bmat[((size_t)kb * 2 + 0) * N * 8 + off] = hi;
        bmat[((size_t)kb * 2 + 1) * N * 8 + off] = w - hi;
      }
  float *d_x, *d_b, *d_out;
  long long *d_cyc;
  int *d_err;
  int nsm = 0;
  CU(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  const size_t nout = (size_t)2 * kCols * N;
  CU(cudaMalloc(&d_x, (kXLen + 4) * sizeof(float)));
  CU(cudaMalloc(&d_b, bmat.size() * sizeof(float)));
  CU(cudaMalloc(&d_out, nout * sizeof(float)));
  CU(cudaMalloc(&d_cyc, nsm * sizeof(long long)));
  CU(cudaMalloc(&d_err, sizeof(int)));
  CU(cudaMemset(d_err, 0, sizeof(int)));
  CU(cudaMemset(d_out, 0, nout * sizeof(float)));
  CU(cudaMemcpy(d_x, x.data(), (kXLen + 4) * sizeof(float), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_b, bmat.data(), bmat.size() * sizeof(float), cudaMemcpyHostToDevice));
  Params p{d_x, d_b, d_out, d_cyc, d_err, iters, hankel};
  const size_t smem = (size_t)(2 * 4 * kXLen + 4 * 2 * N * 8 + 2 * 4 * 4 * 1024) * sizeof(float);
  CU(cudaFuncSetAttribute(k_tc_probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  k_tc_probe<N><<<nsm, 128, smem>>>(p);  // warm-up
  CU(cudaDeviceSynchronize());
  CU(cudaEventRecord(e0));
  k_tc_probe<N><<<nsm, 128, smem>>>(p);
  CU(cudaEventRecord(e1));
  CU(cudaDeviceSynchronize());
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  int err = 0;
  std::vector<float> out(nout);
  std::vector<long long> cyc(nsm);
  CU(cudaMemcpy(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(out.data(), d_out, out.size() * sizeof(float), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(cyc.data(), d_cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost));
  double maxrel[2] = {0, 0}, fp32rel = 0;
  for (int m = 0; m < kCols; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      float f32 = 0;
      for (int u = 0; u < kU; ++u) {
        ref += W[(size_t)n * kU + u] * (double)x[m + u];
        f32 = std::fmaf((float)W[(size_t)n * kU + u], x[m + u], f32);
      }
      if (ref == 0) continue;
      for (int v = 0; v < 2; ++v) maxrel[v] = std::fmax(maxrel[v], std::fabs(out[((size_t)v * kCols + m) * N + n] - ref) / std::fabs(ref));
      fp32rel = std::fmax(fp32rel, std::fabs((double)f32 - ref) / std::fabs(ref));
    }
  if (std::getenv("TC_PROBE_DEBUG")) {
    for (int m : {0, 1, 5, 511})
      for (int n : {0, 2, 30}) {
        double ref = 0;
        for (int u = 0; u < kU; ++u) ref += W[(size_t)n * kU + u] * (double)x[m + u];
        std::printf("  col=%d n=%d ref=%.6e plain=%.6e split=%.6e\n", m, n, ref, out[((size_t)0 * kCols + m) * N + n], out[((size_t)1 * kCols + m) * N + n]);
      }
  }
  long long cmax = 0;
  for (int k = 0; k < nsm; ++k) cmax = cyc[k] > cmax ? cyc[k] : cmax;
  const double clk_per_mma = (double)cmax / ((double)iters * 48);
  const double cells = (double)nsm * iters * kCols;  // (column, source row) pairs fully expanded onto N targets
  std::printf(
      "N=%3d %s err=%d  max rel error vs FP64: plain TF32 %.2e, 3-product split %.2e (FP32 FMA chain %.2e)\n"
      "       %.1f clk per MMA (128 x %d x 8), %.0f clk per source row of 512 columns = %.2f clk per cell per SM;"
      " %.3e cells/s on %d SMs at the %.2f GHz maximum clock\n",
      N, hankel ? "operand in place (overlapping descriptor)" : "im2col'd operand                         ", err, maxrel[0],
      maxrel[1], fp32rel, clk_per_mma, N, clk_per_mma * 48, clk_per_mma * 48 / kCols,
      (double)nsm * kCols / (clk_per_mma * 48) * clk_ghz() * 1e9, nsm, clk_ghz());
  (void)ms, (void)cells;
  cudaFree(d_x), cudaFree(d_b), cudaFree(d_out), cudaFree(d_cyc), cudaFree(d_err);
}

// ---- roles exchanged: the weights are the A operand (lane m = (shift block j, target n): all four shift blocks in
// ONE instruction), the band row in place is the B operand (N = columns 4q + c, K = 8 shifts).  Lane (j, n) of
// accumulator column q then holds the shift-block-j part of target n at column 4 (q - 2j) + c; a finished target
// cell sums its 4 parts when it is read out:  out(n, 4q + c) = sum_j D_c[(j, n)][q + 2j].
// 12 MMAs (4 residues x 3 products) of 128 x N x 8 per source row and 4 N columns.  TS = weights read from TMEM. ----
template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) k_tc_probe_t(Params p) {
  extern __shared__ __align__(128) uint8_t smraw[];
  float *xs = reinterpret_cast<float *>(smraw);  // [2 parts][4 residues][kXLen]
  float *as = xs + 2 * 4 * kXLen;                // [2 parts][128 * 8]
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5;
  for (int k = t; k < 4 * kXLen; k += 128) {
    const int c = k / kXLen, q = k % kXLen;
    const float v = p.x[q + c];
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    xs[(0 * 4 + c) * kXLen + q] = hi;
    xs[(1 * 4 + c) * kXLen + q] = v - hi;
  }
  for (int k = t; k < 2 * 128 * 8; k += 128) as[k] = p.bmat[k];
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_s;
  const uint32_t a_tmem = tmem + 4 * N;  // TS: weights (hi: 8 columns, lo: 8 columns), lane = row
  if (TS) {
    for (int part = 0; part < 2; ++part)
      for (int k = 0; k < 8; ++k) {
        const float w = as[part * 1024 + (k / 4) * 512 + (t / 8) * 32 + (t % 8) * 4 + (k % 4)];
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(a_tmem + ((uint32_t)(warp * 32) << 16) + part * 8 + k), "r"(__float_as_uint(w)));
      }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  const uint32_t idesc = make_idesc(kM, N);
  const uint32_t barA = smem_u32(&bar);
  auto mma = [&](int c, int apart, int bpart, uint32_t acc) {
    const uint64_t b = make_desc(smem_u32(xs + (bpart * 4 + c) * kXLen), 16, 128);  // B[q][k] = x[4 q + c + k], in place
    if (TS) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem + c * N),
          "r"(a_tmem + apart * 8), "l"(b), "r"(idesc), "r"(acc)
          : "memory");
    } else {
      mma_tf32(tmem + c * N, make_desc(smem_u32(as + apart * 1024), 2048, 128), b, idesc, acc);
    }
  };
  uint32_t parity = 0;
  bool ok = true;
  for (int variant = 0; variant < 2; ++variant) {
    if (t == 0) {
      for (int c = 0; c < 4; ++c) {
        mma(c, 0, 0, 0);
        if (variant == 1) mma(c, 0, 1, 1), mma(c, 1, 0, 1);
      }
      mma_commit(barA);
    }
    ok = ok && bar_wait_bounded(barA, parity, 200000000LL);
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (ok && blockIdx.x == 0) {
      for (int n0 = 0; n0 < 4 * N; n0 += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + n0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) p.out[((size_t)variant * 128 + t) * 4 * N + n0 + j] = __uint_as_float(r[j]);  // [variant][lane][c * N + q]
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  long long t0 = 0, t1 = 0;
  if (ok) {
    if (t == 0) {
      t0 = clock64();
      for (int it = 0; it < p.iters; ++it) {
#pragma unroll
        for (int c = 0; c < 4; ++c) mma(c, 0, 0, 1), mma(c, 0, 1, 1), mma(c, 1, 0, 1);
      }
      mma_commit(barA);
    }
    ok = bar_wait_bounded(barA, parity, 4000000000LL);
    t1 = clock64();
  }
  // ---- read-out of ONE target row (lane n of each of the 4 shift-block quarters, all 4 N accumulator columns) into
  // shared memory: a warp can only read whole 32-lane quarters, so 31 of the 32 lanes it loads are discarded ----
  asm volatile("tcgen05.fence::after_thread_sync;");
  __syncthreads();
  const long long t2 = clock64();
  if (ok) {
    for (int rep = 0; rep < 64; ++rep) {
      const int sel = rep & 31;
      for (int n0 = 0; n0 < 4 * N; n0 += 32) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(tmem + ((uint32_t)(warp * 32) << 16) + n0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if ((t & 31) == sel) {
          uint4 *dst = reinterpret_cast<uint4 *>(xs + warp * kXLen + n0);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        }
      }
    }
  }
  __syncthreads();
  const long long t3 = clock64();
  if (t == 0) {
    p.cycles[blockIdx.x] = t1 - t0;
    if (blockIdx.x == 0) p.cycles[gridDim.x] = (t3 - t2) / 64;
    if (!ok) atomicExch(p.err, 1);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

template <int N, bool TS>
static void run_t(int iters, const std::vector<float> &x, const std::vector<double> &W /* [32][32] */) {
  // A = weights, row m = j * 32 + n, k = 0..7: W[n][8 j + k]; K-major canonical order, hi and lo parts
  std::vector<float> amat((size_t)2 * 1024, 0.f);
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < 8; ++k) {
      const float w = (float)W[(size_t)(m % 32) * kU + 8 * (m / 32) + k];
      uint32_t bits;
      std::memcpy(&bits, &w, 4);
      bits &= 0xFFFFE000u;
      float hi;
      std::memcpy(&hi, &bits, 4);
      const size_t off = (size_t)(k / 4) * 512 + (size_t)(m / 8) * 32 + (m % 8) * 4 + (k % 4);
      amat[off] = hi;
      amat[1024 + off] = w - hi;
    }
  float *d_x, *d_b, *d_out;
  long long *d_cyc;
  int *d_err;
  int nsm = 0;
  CU(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  const size_t nout = (size_t)2 * 128 * 4 * N;
  CU(cudaMalloc(&d_x, (kXLen + 4) * sizeof(float)));
  CU(cudaMalloc(&d_b, amat.size() * sizeof(float)));
  CU(cudaMalloc(&d_out, nout * sizeof(float)));
  CU(cudaMalloc(&d_cyc, (nsm + 1) * sizeof(long long)));
  CU(cudaMalloc(&d_err, sizeof(int)));
  CU(cudaMemset(d_err, 0, sizeof(int)));
  CU(cudaMemset(d_out, 0, nout * sizeof(float)));
  CU(cudaMemcpy(d_x, x.data(), (kXLen + 4) * sizeof(float), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_b, amat.data(), amat.size() * sizeof(float), cudaMemcpyHostToDevice));
  Params p{d_x, d_b, d_out, d_cyc, d_err, iters, 1};
  const size_t smem = (size_t)(2 * 4 * kXLen + 2 * 1024) * sizeof(float);
  CU(cudaFuncSetAttribute(k_tc_probe_t<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  k_tc_probe_t<N, TS><<<nsm, 128, smem>>>(p);
  CU(cudaDeviceSynchronize());
  CU(cudaEventRecord(e0));
  k_tc_probe_t<N, TS><<<nsm, 128, smem>>>(p);
  CU(cudaEventRecord(e1));
  CU(cudaDeviceSynchronize());
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  int err = 0;
  std::vector<float> out(nout);
  std::vector<long long> cyc(nsm + 1);
  CU(cudaMemcpy(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(out.data(), d_out, out.size() * sizeof(float), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(cyc.data(), d_cyc, (nsm + 1) * sizeof(long long), cudaMemcpyDeviceToHost));
  double maxrel[2] = {0, 0};
  const int qmax = N - 6;  // columns whose four parts all lie inside the accumulator
  for (int q = 0; q < qmax; ++q)
    for (int c = 0; c < 4; ++c)
      for (int n = 0; n < 31; ++n) {
        double ref = 0;
        for (int u = 0; u < kU; ++u) ref += W[(size_t)n * kU + u] * (double)x[4 * q + c + u];
        if (ref == 0) continue;
        for (int v = 0; v < 2; ++v) {
          float sum = 0;
          for (int j = 0; j < 4; ++j) sum += out[((size_t)v * 128 + j * 32 + n) * 4 * N + c * N + q + 2 * j];
          maxrel[v] = std::fmax(maxrel[v], std::fabs(sum - ref) / std::fabs(ref));
        }
      }
  long long cmax = 0;
  for (int k = 0; k < nsm; ++k) cmax = cyc[k] > cmax ? cyc[k] : cmax;
  const double clk_per_mma = (double)cmax / ((double)iters * 12);
  const double cells = (double)nsm * iters * 4 * qmax;
  std::printf(
      "roles exchanged, weights from %s, N=%3d err=%d  max rel error vs FP64: plain TF32 %.2e, 3-product split %.2e\n"
      "       %.1f clk per MMA (128 x %d x 8), %.0f clk per source row of %d columns = %.2f clk per cell per SM;"
      " %.3e cells/s on %d SMs\n"
      "       read-out of one finished target row (4 quarters x %d accumulator columns -> shared memory, 4 warps): %lld clk\n",
      TS ? "TMEM" : "smem", N, err, maxrel[0], maxrel[1], clk_per_mma, N, clk_per_mma * 12, 4 * qmax, clk_per_mma * 12 / (4 * qmax),
      (double)nsm * 4 * qmax / (clk_per_mma * 12) * 1.0 * clk_ghz() * 1e9, nsm, 4 * N, cyc[nsm]);
  (void)ms, (void)cells;
  cudaFree(d_x), cudaFree(d_b), cudaFree(d_out), cudaFree(d_cyc), cudaFree(d_err);
}

int main(int argc, char **argv) {
  const int iters = argc > 1 ? std::atoi(argv[1]) : 5000;
  // a band row with the dynamic range of Boltzmann-weighted partition functions, and interior-loop-like weights:
  // W[n][u] = E(n) * r^min(|n - 2u|, 6) for u <= n (size term x asymmetry term, raccess.cpp:796-816), else 0
  std::vector<float> x(kXLen + 4);
  uint64_t s = 88172645463325252ull;
  auto rnd = [&]() {
    s ^= s << 13, s ^= s >> 7, s ^= s << 17;
    return (double)(s >> 11) / 9007199254740992.0;
  };
  for (auto &v : x) v = (float)std::exp(8.0 * rnd() - 4.0);
  auto weights = [&](int N) {
    std::vector<double> W((size_t)N * kU, 0.0);
    for (int n = 0; n < N; ++n)
      for (int u = 0; u < kU; ++u) {
        const int tot = n % 32;
        if (u > tot || tot > 30) continue;
        const int asym = std::abs(tot - 2 * u);
        W[(size_t)n * kU + u] = std::exp(-(1.7 + 1.08 * std::log((double)tot + 1.0)) / 0.616) * std::pow(0.45, asym < 6 ? asym : 6);
      }
    return W;
  };
  run<32>(1, iters, x, weights(32));
  run<32>(0, iters, x, weights(32));
  run<64>(1, iters, x, weights(64));
  run<128>(1, iters, x, weights(128));
  run_t<128, false>(iters, x, weights(32));
  run_t<112, true>(iters, x, weights(32));
  run_t<64, true>(iters, x, weights(32));
  return 0;
}
