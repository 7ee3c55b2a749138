// Shared host/device helpers for librfk (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "rfk.h"

namespace rfk {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return RFK_ECUDA;
  }
  return RFK_OK;
}

#define RFK_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      rfk::set_error(__VA_ARGS__);        \
      return RFK_EINVAL;                  \
    }                                     \
  } while (0)

int sm_count();
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define RFK_LAUNCH(kernel, grid, block, smem, stream, ...) \
  rfk::launch_kernel(kernel, dim3(grid), dim3(block), smem, stream, __VA_ARGS__)

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// grid for a bandwidth-bound grid-stride kernel: a whole number of waves over the SMs
inline int stream_grid(long long work_items, int threads, int ctas_per_sm) {
  long long need = (work_items + threads - 1) / threads;
  long long wave = (long long)sm_count() * ctas_per_sm;
  if (need <= wave) return (int)(need > 0 ? need : 1);
  return (int)wave;
}

// Programmatic dependent launch (PDL).  Every kernel of this library is launched with the programmatic-stream-
// serialization attribute, calls pdl_wait() before it reads anything a preceding kernel may have written, and calls
// pdl_trigger() early so that the next kernel in the stream can be scheduled and run its own prologue (barrier init,
// TMEM allocation, weight prefetch) while this one is still working.  Opt-in with RFK_PDL=1 (plain launches otherwise).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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

// streaming 128-bit access: read-once / write-once data should not pollute L1
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// log-scale clamp of the affine coupling (Flow/glow_modules.py:252-268)
__device__ __forceinline__ float clamp_ls(float s, int kind, float a, float b) {
  switch (kind) {
    case RFK_CLAMP_REALNVP: return a * tanhf(s) + b;
    case RFK_CLAMP_GLOW: {
      // log(sigmoid(s+2)) = -softplus(-(s+2)), evaluated without overflow
      float t = -(s + 2.0f);
      return -(t > 15.0f ? t : log1pf(expf(t)));
    }
    case RFK_CLAMP_SOFT: return 2.5f * 0.636f * atanf(s * (1.0f / 2.5f));
    default: return s;
  }
}

__device__ __forceinline__ float std_from_raw(float raw, int kind) {
  if (kind == RFK_STD_EXP) return expf(raw);
  // torch softplus (beta=1, threshold=20) + 1e-8 (Flow/glow_modules.py:340-341)
  float sp = raw > 20.0f ? raw : log1pf(expf(raw));
  return sp + 1e-8f;
}

}  // namespace rfk
