// Shared device/host helpers for libscrna_b200 (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/scrna_b200.h"

namespace scrna {

constexpr float kEpsMU = 1.1920928955078125e-07f;  // float32 eps: sklearn _nmf.py:32 EPSILON

// Device-resident control block of one NMF fit (mirrors scrna_nmf_ctrl_t in include/scrna_b200.h).
struct NmfCtrl {
  int iter;     // completed iterations
  int done;     // stop rule fired: every kernel of the loop early-exits
  int n_iter;   // iteration count at which the rule fired (0 while running)
  int flags;    // SCRNA_FLAG_*
  double viol_init;
  double viol_last;
  double err_last;
  double reserved[3];
};
static_assert(sizeof(NmfCtrl) == 64, "ctrl block layout is part of the C ABI");

__device__ __forceinline__ bool ctrl_done(const NmfCtrl* c) { return c != nullptr && c->done != 0; }

// ---- streaming (read-once) loads of X: bypass L1 allocation, keep the factors there ----------
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 v;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ld_stream(const float2* p) {
  float2 v;
  asm("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// C consecutive floats as a small register vector
template <int C>
struct Vec;
template <>
struct Vec<4> {
  float v[4];
  __device__ __forceinline__ static Vec<4> stream(const float* p) {
    float4 t = ld_stream(reinterpret_cast<const float4*>(p));
    return Vec<4>{{t.x, t.y, t.z, t.w}};
  }
  __device__ __forceinline__ static Vec<4> cached(const float* p) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    return Vec<4>{{t.x, t.y, t.z, t.w}};
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <>
struct Vec<2> {
  float v[2];
  __device__ __forceinline__ static Vec<2> stream(const float* p) {
    float2 t = ld_stream(reinterpret_cast<const float2*>(p));
    return Vec<2>{{t.x, t.y}};
  }
  __device__ __forceinline__ static Vec<2> cached(const float* p) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    return Vec<2>{{t.x, t.y}};
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  }
};
template <>
struct Vec<1> {
  float v[1];
  __device__ __forceinline__ static Vec<1> stream(const float* p) { return Vec<1>{{ld_stream(p)}}; }
  __device__ __forceinline__ static Vec<1> cached(const float* p) { return Vec<1>{{__ldg(p)}}; }
  __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
};

// C consecutive elements of a row of X (fp32 or fp16 storage) as fp32, streaming
template <int C>
__device__ __forceinline__ Vec<C> load_x(const float* p) { return Vec<C>::stream(p); }
template <int C>
__device__ __forceinline__ Vec<C> load_x(const __half* p);
template <>
__device__ __forceinline__ Vec<4> load_x<4>(const __half* p) {
  uint32_t a, b;
  asm("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p));
  const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&a));
  const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&b));
  return Vec<4>{{lo.x, lo.y, hi.x, hi.y}};
}
template <>
__device__ __forceinline__ Vec<2> load_x<2>(const __half* p) {
  uint32_t a;
  asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(a) : "l"(p));
  const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&a));
  return Vec<2>{{v.x, v.y}};
}
template <>
__device__ __forceinline__ Vec<1> load_x<1>(const __half* p) {
  uint16_t a;
  asm("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(a) : "l"(p));
  return Vec<1>{{__half2float(__ushort_as_half(a))}};
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block-wide sum (fixed tree); result valid in thread 0. `scratch` >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    double t = lane < nw ? scratch[lane] : 0.0;
    t = warp_sum(t);
    v = t;
  }
  return v;
}

inline int cuda_status(cudaError_t e) { return e == cudaSuccess ? SCRNA_OK : SCRNA_ERR_CUDA; }
inline int last_launch_status() { return cuda_status(cudaGetLastError()); }

inline int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  }
  return sms;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace scrna
