// Shared helpers for the xmlb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#define XMLB_OK 0
#define XMLB_EINVAL (-1)
#define XMLB_EUNSUPPORTED (-2)

// thread-local last-error text, read through xmlb_last_error() (see include/xmlb200.h)
void xmlb_set_error(const char* fmt, ...);

#define XMLB_REQUIRE(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      xmlb_set_error(__VA_ARGS__);              \
      return XMLB_EINVAL;                       \
    }                                           \
  } while (0)

#define XMLB_CUDA(expr)                                                                  \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      xmlb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                    \
    }                                                                                    \
  } while (0)

#define XMLB_LAUNCH_CHECK()                                                              \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      xmlb_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                    \
    }                                                                                    \
  } while (0)

// launch counter (bench.py reports it as gpu_launches)
void xmlb_count_launch(int n);

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

#define MASK_FILL (-1e10f)      // reference model_xml.py:640-641
#define ATT_MASK_FILL (-10000.f)  // reference model_components.py:277

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// reference mask_logits(x, m) = x*m + (1-m)*(-1e10), evaluated literally in fp32
__device__ __forceinline__ float mask_logit(float x, float m) {
  return __fadd_rn(__fmul_rn(x, m), __fmul_rn(__fsub_rn(1.f, m), MASK_FILL));
}

// float atomic max that is correct for mixed signs (destination must be initialised)
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// Counter-based dropout mask (xmlb_dropout, xmlb_attention_train and their backward passes): keep(i) depends only on
// (seed, i) -- two rounds of a 32-bit multiply-xorshift mix (the murmur3 finaliser) over both halves of the 64-bit
// element index and seed, compared with p * 2^32 -- so no mask is ever stored.
__device__ __forceinline__ uint32_t mix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}
__device__ __forceinline__ bool dropout_keep(unsigned long long seed, unsigned long long i, uint32_t threshold) {
  uint32_t h = mix32((uint32_t)i ^ (uint32_t)seed);
  h = mix32(h + 0x9E3779B9u * (uint32_t)(i >> 32) + (uint32_t)(seed >> 32));
  return h >= threshold;
}
static inline uint32_t dropout_threshold(float p) {
  const double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
}
