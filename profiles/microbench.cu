// Issue-rate probes behind the design of the tile kernels (profiles/r2/microbench.txt holds the output):
//   * FFMA2 in the form the stencils use (scalar multiplicand, coefficient pair from the constant bank)
//   * shared-memory load streams (LDS.32 / .64 / .128, conflict free)
//   * the stencil inner loop as it is (1 column per thread: 1 LDS.32 feeds 2 FFMA2) and with register tiling
//     over 2 or 4 adjacent columns (1 LDS.64 / LDS.128 feeds 8 / 16 FFMA2)
// Persistent grid of one CTA per SM, like the tile kernels.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

__constant__ float2 g_pair[64];

__device__ __forceinline__ void ffma2_bcast(unsigned long long &acc, float v, float2 g) {
  unsigned long long vv, gg;
  asm("mov.b64 %0, {%1, %1};" : "=l"(vv) : "f"(v));
  asm("mov.b64 %0, {%1, %2};" : "=l"(gg) : "f"(g.x), "f"(g.y));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(vv), "l"(gg));
}

constexpr int kRow = 2048;   // floats per smem row
constexpr int kRows = 16;

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_probe(float *sink, int iters, int nthr_active) {
  extern __shared__ __align__(16) float sm[];
  const int t = threadIdx.x;
  for (int k = t; k < kRow * kRows; k += blockDim.x) sm[k] = 1.0f + 1e-3f * (k & 15);
  __syncthreads();
  if (t >= nthr_active) return;
  unsigned long long a[8];
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = 0;
  float f[8];
#pragma unroll
  for (int k = 0; k < 8; k++) f[k] = 0;
  for (int it = 0; it < iters; it++) {
    const float *row = sm + (it & (kRows - 1)) * kRow;
    if (MODE == 0) {  // FFMA2 only: 8 chains, 64 FFMA2 per iteration
      const float v = row[t];
#pragma unroll
      for (int x = 0; x < 8; x++) {
#pragma unroll
        for (int k = 0; k < 8; k++) ffma2_bcast(a[k], v, g_pair[(x * 8 + k) & 63]);
      }
    } else if (MODE == 1) {  // LDS.32 stream: 32 loads + 32 FADD per iteration
#pragma unroll
      for (int x = 0; x < 32; x++) f[x & 7] += row[t + x];
    } else if (MODE == 2) {  // LDS.64 stream: 32 loads (64 floats)
#pragma unroll
      for (int x = 0; x < 32; x++) {
        const float2 v = *reinterpret_cast<const float2 *>(row + 2 * t + 2 * x);
        f[x & 7] += v.x + v.y;
      }
    } else if (MODE == 3) {  // LDS.128 stream: 32 loads (128 floats)
#pragma unroll
      for (int x = 0; x < 32; x++) {
        const float4 v = *reinterpret_cast<const float4 *>(row + 4 * (t & 255) + 4 * x);
        f[x & 7] += (v.x + v.y) + (v.z + v.w);
      }
    } else if (MODE == 4) {  // stencil as is: per element 1 LDS.32 + 2 FFMA2 (4 targets); 32 elements
#pragma unroll
      for (int x = 0; x < 32; x++) {
        const float v = row[t + 1 + x];
        ffma2_bcast(a[0], v, g_pair[(2 * x) & 63]);
        ffma2_bcast(a[1], v, g_pair[(2 * x + 1) & 63]);
      }
    } else if (MODE == 5) {  // 2 columns per thread: per LDS.64 (2 elements) 8 FFMA2; 32 elements -> 2 x 32 x 2 FFMA2
#pragma unroll
      for (int x = 0; x < 32; x += 2) {
        const float2 v = *reinterpret_cast<const float2 *>(row + 2 * t + 2 + x);
        // column A takes element j as offset x, column B as offset x - 1
        ffma2_bcast(a[0], v.x, g_pair[(2 * x) & 63]);
        ffma2_bcast(a[1], v.x, g_pair[(2 * x + 1) & 63]);
        ffma2_bcast(a[2], v.x, g_pair[(2 * x + 2) & 63]);
        ffma2_bcast(a[3], v.x, g_pair[(2 * x + 3) & 63]);
        ffma2_bcast(a[0], v.y, g_pair[(2 * x + 4) & 63]);
        ffma2_bcast(a[1], v.y, g_pair[(2 * x + 5) & 63]);
        ffma2_bcast(a[2], v.y, g_pair[(2 * x + 6) & 63]);
        ffma2_bcast(a[3], v.y, g_pair[(2 * x + 7) & 63]);
      }
    } else if (MODE == 6) {  // 4 columns per thread: per LDS.128 (4 elements) 32 FFMA2
#pragma unroll
      for (int x = 0; x < 32; x += 4) {
        const float4 v = *reinterpret_cast<const float4 *>(row + 4 * (t & 255) + 4 + x);
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
#pragma unroll
          for (int k = 0; k < 8; k++) ffma2_bcast(a[k], e[j], g_pair[(8 * x + 8 * j + k) & 63]);
        }
      }
    } else if (MODE == 7) {  // scalar FFMA with constant-bank coefficient: per element 1 LDS.32 + 4 FFMA
#pragma unroll
      for (int x = 0; x < 32; x++) {
        const float v = row[t + 1 + x];
        f[0] = fmaf(v, g_pair[(2 * x) & 63].x, f[0]);
        f[1] = fmaf(v, g_pair[(2 * x) & 63].y, f[1]);
        f[2] = fmaf(v, g_pair[(2 * x + 1) & 63].x, f[2]);
        f[3] = fmaf(v, g_pair[(2 * x + 1) & 63].y, f[3]);
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[k]));
    s += lo + hi + f[k];
  }
  if (s == 123.456f) sink[blockIdx.x * blockDim.x + t] = s;
}

template <int MODE>
double run(int threads, int iters, float *sink, int nsm) {
  const size_t smem = sizeof(float) * kRow * kRows + 4096 * 4;
  cudaFuncSetAttribute(k_probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 1e30;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    k_probe<MODE><<<nsm, threads, smem>>>(sink, iters, threads);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  return best;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int nsm = prop.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double clk = khz * 1e3;  // nominal max SM clock
  float2 pairs[64];
  for (int k = 0; k < 64; k++) pairs[k] = make_float2(1.0f + 1e-4f * k, 1.0f - 1e-4f * k);
  cudaMemcpyToSymbol(g_pair, pairs, sizeof(pairs));
  float *sink;
  cudaMalloc(&sink, sizeof(float) * nsm * 1024);
  const int iters = 20000;
  printf("%s, %d SMs, nominal %.0f MHz; per-SM rates assume the nominal clock\n", prop.name, nsm, clk / 1e6);
  struct Row { const char *name; int mode; double per_iter; const char *unit; };
  const Row rows[] = {
      {"FFMA2 (scalar x const pair)", 0, 64, "FFMA2"},
      {"LDS.32 stream", 1, 32, "LDS"},
      {"LDS.64 stream", 2, 32, "LDS"},
      {"LDS.128 stream", 3, 32, "LDS"},
      {"stencil 1 col/thread (LDS.32 + 2 FFMA2)", 4, 32 * 4, "FMA"},
      {"stencil 2 col/thread (LDS.64 + 8 FFMA2)", 5, 32 * 8, "FMA"},
      {"stencil 4 col/thread (LDS.128 + 32 FFMA2)", 6, 32 * 16, "FMA"},
      {"stencil 1 col/thread scalar FFMA (LDS.32 + 4 FFMA)", 7, 32 * 4, "FMA"},
  };
  for (int threads : {256, 352, 704, 1024}) {
    for (const Row &r : rows) {
      if (r.mode == 3 || r.mode == 6) { if (threads > 256) continue; }
      double ms = 0;
      switch (r.mode) {
        case 0: ms = run<0>(threads, iters, sink, nsm); break;
        case 1: ms = run<1>(threads, iters, sink, nsm); break;
        case 2: ms = run<2>(threads, iters, sink, nsm); break;
        case 3: ms = run<3>(threads, iters, sink, nsm); break;
        case 4: ms = run<4>(threads, iters, sink, nsm); break;
        case 5: ms = run<5>(threads, iters, sink, nsm); break;
        case 6: ms = run<6>(threads, iters, sink, nsm); break;
        default: ms = run<7>(threads, iters, sink, nsm); break;
      }
      const double warps = threads / 32.0;
      const double per_sm_clk = r.per_iter * iters * warps / (ms * 1e-3 * clk);
      printf("threads=%4d  %-52s %8.3f ms  %7.3f warp-%s/clk/SM  (%.1f thread-%s/clk/SM)\n", threads, r.name, ms,
             per_sm_clk, r.unit, per_sm_clk * 32, r.unit);
    }
  }
  return 0;
}
