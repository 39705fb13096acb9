// Where does the start-up time of a multi-GPU process go?  (profiles/r1/cuda_init_8gpu.txt)
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv) {
  size_t mb = argc > 1 ? atol(argv[1]) : 1500;
  double t0 = now();
  int n = 0;
  cudaGetDeviceCount(&n);
  double t1 = now();
  cudaSetDevice(0);
  cudaFree(0);
  double t2 = now();
  void *p = nullptr;
  cudaMallocHost(&p, mb << 20);
  double t3 = now();
  for (int d = 1; d < n; d++) { cudaSetDevice(d); cudaFree(0); }
  double t4 = now();
  void *q = nullptr;
  cudaMallocHost(&q, mb << 20);
  double t5 = now();
  printf("devices %d: cudaGetDeviceCount %.3f s, context 0 %.3f s, cudaMallocHost(%zu MB) %.3f s, contexts 1..n-1 (serial) %.3f s, "
         "cudaMallocHost again with all contexts live %.3f s\n", n, t1 - t0, t2 - t1, mb, t3 - t2, t4 - t3, t5 - t4);
  return 0;
}
