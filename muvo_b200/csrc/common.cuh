// Shared helpers for libmuvo_b200 (sm_100a).  Internal header, not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/muvo_b200.h"

#define MUVO_LAUNCH_CHECK()                                   \
  do {                                                        \
    cudaError_t e__ = cudaPeekAtLastError();                  \
    if (e__ != cudaSuccess) return (int)cudaGetLastError();   \
  } while (0)
// launch check + optional profiling mark named after the kernel
#define MUVO_AFTER_LAUNCH(name, st)                           \
  do {                                                        \
    MUVO_LAUNCH_CHECK();                                      \
    ::muvo::prof_mark(name, st);                              \
  } while (0)

namespace muvo {

constexpr int kNumSMsB200 = 148;
extern int g_tuning[8];   // muvo_debug_set_tuning (points.cu)

// Optional per-kernel timing (muvo_profile_begin/end): when active on this host thread, every launch site
// records a CUDA event on the launching stream right after its kernel.  Inactive -> a single branch.
constexpr int kProfMax = 64;
struct Profile {
  cudaEvent_t ev[kProfMax];
  const char* name[kProfMax];
  int n = 0;
  int created = 0;
  bool on = false;
};
Profile& profile_state();
inline void prof_mark(const char* name, cudaStream_t st) {
  Profile& p = profile_state();
  if (!p.on || p.n >= kProfMax) return;
  if (p.n >= p.created) { if (cudaEventCreate(&p.ev[p.created]) != cudaSuccess) return; ++p.created; }
  p.name[p.n] = name;
  cudaEventRecord(p.ev[p.n], st);
  ++p.n;
}

// A failed runtime call (cudaFuncSetAttribute, cudaMemsetAsync, ...) is reported through the return value; the runtime's
// "last error" is cleared as well, otherwise the NEXT library call's launch check would report it again as its own.
static inline int cuda_fail(cudaError_t e) { (void)cudaGetLastError(); return (int)e; }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
__host__ __device__ static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Streaming (read-once) loads / write-once stores: keep them out of L1 and mark them
// first-to-evict in L2 so that the small L2-resident tables survive next to them.
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream_u4(uint4* p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_f4(float4* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

}  // namespace muvo
