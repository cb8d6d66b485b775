// yy_common.cuh -- error plumbing, launch accounting, Philox, small device helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/yinyang_b200.h"
#include "yy_rules.cuh"

namespace yy {

// ---- error state (thread-local message; integer status across the ABI, never exceptions) ----
char* error_buffer();
int set_error(int status, const char* fmt, ...);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define YY_CUDA_OK(expr)                                                                      \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return yy::set_error(YY_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,          \
                           cudaGetErrorString(_e));                                           \
  } while (0)

#define YY_LAUNCH_CHECK()                                                                     \
  do {                                                                                        \
    yy::count_launch();                                                                       \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess)                                                                    \
      return yy::set_error(YY_ERR_CUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__,      \
                           cudaGetErrorString(_e));                                           \
  } while (0)

inline int words_for_cells(int cells) { return (cells + 63) >> 6; }
// template width used for a board: 1, 2 or 4 words (cells <= 256)
inline int nw_for_cells(int cells) { return cells <= 64 ? 1 : (cells <= 128 ? 2 : 4); }
inline bool board_supported(int rows, int cols) {
  return rows >= 1 && cols >= 1 && rows <= 32 && cols <= 32 && rows * cols <= 256;
}

#define YY_DISPATCH_NW(cells, ...)                          \
  do {                                                      \
    int _nw = yy::nw_for_cells(cells);                      \
    if (_nw == 1) { constexpr int NW = 1; __VA_ARGS__; }    \
    else if (_nw == 2) { constexpr int NW = 2; __VA_ARGS__; } \
    else { constexpr int NW = 4; __VA_ARGS__; }             \
  } while (0)

// ---- board load/store between global word arrays ([i*W + w]) and BB<NW> ----
template <int NW>
__device__ __forceinline__ BB<NW> load_bb(const uint64_t* __restrict__ base, long long i, int W) {
  BB<NW> r;
#pragma unroll
  for (int k = 0; k < NW; ++k) r.w[k] = k < W ? base[i * W + k] : 0ull;
  return r;
}
template <int NW>
__device__ __forceinline__ void store_bb(uint64_t* __restrict__ base, long long i, int W, const BB<NW>& v) {
#pragma unroll
  for (int k = 0; k < NW; ++k)
    if (k < W) base[i * W + k] = v.w[k];
}

// ---- Philox4x32-10 counter-based RNG ----
struct Philox {
  uint32_t key[2];
  uint32_t ctr[4];
  uint32_t out[4];
  int have;
  __host__ __device__ Philox(uint64_t seed, uint64_t stream, uint64_t substream) {
    key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
    ctr[0] = 0; ctr[1] = (uint32_t)substream; ctr[2] = (uint32_t)stream; ctr[3] = (uint32_t)(stream >> 32) ^ (uint32_t)(substream >> 32);
    have = 0;
  }
  __host__ __device__ void round_once(uint32_t k0, uint32_t k1) {
    uint64_t p0 = (uint64_t)0xD2511F53u * ctr[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * ctr[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ ctr[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ ctr[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
  }
  __host__ __device__ void refill() {
    uint32_t save[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) { round_once(k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    out[0] = ctr[0]; out[1] = ctr[1]; out[2] = ctr[2]; out[3] = ctr[3];
    ctr[0] = save[0] + 1; ctr[1] = save[1]; ctr[2] = save[2]; ctr[3] = save[3];
    have = 4;
  }
  __host__ __device__ uint32_t next() { if (!have) refill(); return out[--have]; }
  // uniform in (0,1), 53-bit
  __host__ __device__ double uniform() {
    uint64_t hi = next(), lo = next();
    uint64_t v = ((hi << 32) | lo) >> 11;
    return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
  }
  __host__ __device__ uint32_t below(uint32_t n) { return (uint32_t)(((uint64_t)next() * n) >> 32); }
};

}  // namespace yy
