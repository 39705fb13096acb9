// Suffix array of one database page on the GPU.
//
// Replaces the reference's `sais(&encoded_sequences[0], &suffix_array[0], n)` (db_construction.cpp:334,
// sais.cpp:656) behind the same (text, n) -> SA contract.  The suffix array of a text is unique, so any
// correct construction yields the bytes the reference writes into <db>.ind.
//
// Method: prefix doubling.  Round 0 sorts the suffixes by their first 16 symbols (4 bits per symbol, one
// 64-bit key); every later round sorts by (rank[i], rank[i + h]) with h = 16, 32, ...; ranks are 1-based and
// "past the end" is 0, so a suffix that is a proper prefix of another one sorts first, as in sais().
// Random transcripts are fully ranked after round 0 or 1 (4^16 >> n); repeats cost log2(repeat length)
// rounds.  The sorts are cub::DeviceRadixSort (CUDA toolkit library code, like calling cuBLAS); keying,
// ranking and the convergence test are the kernels below.
#include <cuda_runtime.h>

#include <cub/cub.cuh>

#include <cstdint>
#include <string>

#include "../../include/priblast_acc.h"

extern "C" void prib_internal_set_error(const char *msg);  // acc_kernels.cu: text behind prib_last_error()

namespace {

#define SA_CU(call)                                                                            \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      prib_internal_set_error((std::string(#call) + ": " + cudaGetErrorString(e_)).c_str());  \
      rc = PRIB_ECUDA;                                                                         \
      goto done;                                                                               \
    }                                                                                          \
  } while (0)

// key0[i] = symbols text[i .. i+15], most significant first, 4 bits each, stored as symbol + 1 (0 = past the end)
__global__ void __launch_bounds__(256) k_sa_key0(const uint8_t *text, long long n, unsigned long long *key,
                                                 int32_t *idx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long k = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const long long p = i + j;
    const unsigned long long s = p < n ? (unsigned long long)(text[p] & 15u) + 1ull : 0ull;
    k = (k << 4) | s;
  }
  key[i] = k;
  idx[i] = (int32_t)i;
}

// flag[j] = 1 where the sorted key changes (start of a new rank class)
__global__ void __launch_bounds__(256) k_sa_flags(const unsigned long long *key, long long n, int32_t *flag) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  flag[j] = (j == 0 || key[j] != key[j - 1]) ? 1 : 0;
}

// rank[sa[j]] = class number (1-based, from the inclusive scan of the flags)
__global__ void __launch_bounds__(256) k_sa_scatter_rank(const int32_t *sa, const int32_t *cls, long long n,
                                                         int32_t *rank) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  rank[sa[j]] = cls[j];
}

// key[i] = rank[i] * (n + 1) + rank[i + h]   (second half past the end = 0); fits 64 bits for n < 2^31
__global__ void __launch_bounds__(256) k_sa_key_pair(const int32_t *rank, long long n, long long h,
                                                     unsigned long long *key, int32_t *idx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long r2 = i + h < n ? (unsigned long long)rank[i + h] : 0ull;
  key[i] = (unsigned long long)rank[i] * (unsigned long long)(n + 1) + r2;
  idx[i] = (int32_t)i;
}

int bits_for(unsigned long long v) {
  int b = 1;
  while (b < 64 && (v >> b) != 0) ++b;
  return b;
}

}  // namespace

extern "C" {

int prib_suffix_array(const unsigned char *text, int32_t n, int32_t *sa, int32_t device) {
  if (n < 0 || (n > 0 && (!text || !sa))) {
    prib_internal_set_error("bad argument");
    return PRIB_EINVAL;
  }
  if (n == 0) return PRIB_OK;
  for (int32_t k = 0; k < n; k++)
    if (text[k] > 14) {  // the database alphabet is 0..9 (encoder.hpp:36-78); keys hold 4 bits per symbol
      prib_internal_set_error("prib_suffix_array: symbols must be < 15 (encoded database text)");
      return PRIB_EINVAL;
    }
  int rc = PRIB_OK;
  const long long N = n;
  const unsigned grid = (unsigned)((N + 255) / 256);
  uint8_t *d_text = nullptr;
  unsigned long long *d_key[2] = {nullptr, nullptr};
  int32_t *d_idx[2] = {nullptr, nullptr}, *d_rank = nullptr, *d_flag = nullptr, *d_cls = nullptr;
  void *d_tmp = nullptr;
  size_t tmp_sort = 0, tmp_scan = 0, tmp_bytes = 0;
  int32_t classes = 0;
  {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
      prib_internal_set_error("no CUDA device available (this library has no CPU fallback)");
      return PRIB_ECUDA;
    }
    if (device < 0 || device >= ndev) {
      prib_internal_set_error("device ordinal out of range");
      return PRIB_EINVAL;
    }
  }
  SA_CU(cudaSetDevice(device));
  SA_CU(cudaMalloc(&d_text, (size_t)N));
  SA_CU(cudaMalloc(&d_key[0], (size_t)N * 8));
  SA_CU(cudaMalloc(&d_key[1], (size_t)N * 8));
  SA_CU(cudaMalloc(&d_idx[0], (size_t)N * 4));
  SA_CU(cudaMalloc(&d_idx[1], (size_t)N * 4));
  SA_CU(cudaMalloc(&d_rank, (size_t)N * 4));
  SA_CU(cudaMalloc(&d_flag, (size_t)N * 4));
  SA_CU(cudaMalloc(&d_cls, (size_t)N * 4));
  {
    cub::DoubleBuffer<unsigned long long> kb(d_key[0], d_key[1]);
    cub::DoubleBuffer<int32_t> vb(d_idx[0], d_idx[1]);
    SA_CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, kb, vb, n, 0, 64));
    SA_CU(cub::DeviceScan::InclusiveSum(nullptr, tmp_scan, d_flag, d_cls, n));
    tmp_bytes = tmp_sort > tmp_scan ? tmp_sort : tmp_scan;
    SA_CU(cudaMalloc(&d_tmp, tmp_bytes));
  }
  SA_CU(cudaMemcpy(d_text, text, (size_t)N, cudaMemcpyHostToDevice));
  {
    const int pair_bits = bits_for((unsigned long long)N * (unsigned long long)(N + 1) + (unsigned long long)N);
    long long h = 0;  // prefix length the current ranks distinguish (0: no ranks yet)
    for (;;) {
      cub::DoubleBuffer<unsigned long long> kb(d_key[0], d_key[1]);
      cub::DoubleBuffer<int32_t> vb(d_idx[0], d_idx[1]);
      if (h == 0) k_sa_key0<<<grid, 256>>>(d_text, N, kb.Current(), vb.Current());
      else k_sa_key_pair<<<grid, 256>>>(d_rank, N, h, kb.Current(), vb.Current());
      size_t tb = tmp_bytes;
      SA_CU(cub::DeviceRadixSort::SortPairs(d_tmp, tb, kb, vb, n, 0, h == 0 ? 64 : pair_bits));
      k_sa_flags<<<grid, 256>>>(kb.Current(), N, d_flag);
      tb = tmp_bytes;
      SA_CU(cub::DeviceScan::InclusiveSum(d_tmp, tb, d_flag, d_cls, n));
      SA_CU(cudaMemcpy(&classes, d_cls + (N - 1), 4, cudaMemcpyDeviceToHost));
      const long long covered = h == 0 ? 16 : 2 * h;
      if (classes == n || covered >= N) {  // every suffix has its own rank
        SA_CU(cudaMemcpy(sa, vb.Current(), (size_t)N * 4, cudaMemcpyDeviceToHost));
        break;
      }
      k_sa_scatter_rank<<<grid, 256>>>(vb.Current(), d_cls, N, d_rank);
      SA_CU(cudaGetLastError());
      h = covered;
    }
  }
done:
  cudaFree(d_text);
  cudaFree(d_key[0]);
  cudaFree(d_key[1]);
  cudaFree(d_idx[0]);
  cudaFree(d_idx[1]);
  cudaFree(d_rank);
  cudaFree(d_flag);
  cudaFree(d_cls);
  cudaFree(d_tmp);
  return rc;
}

}  // extern "C"
