// "Next" row N4, last part: the two torch_scatter reductions of the PointPillar encoder (muvo/models/common.py:698-761,
// POINT_PILLAR.ENABLED, off by default): scatter_mean(xyz, inverse_indices, dim=0) (:731) and
// scatter_max(feat, inverse_indices, dim=0)[0] (:703), with the gradients autograd needs.
//
// torch_scatter is not vendored in the reference and not installed in this image; the semantics restated here are its
// documented ones: out[m] = mean / max over {n : index[n] == m}; rows that receive nothing are 0; the mean divides by
// max(count, 1); scatter_max also returns arg[m, f] = a source row that attains the maximum (N for empty rows).
// Sizes are small (1e5 points x 3..32 features), so these are latency-sized kernels; what matters is one launch per
// reduction and no host round trip.  scatter_max is deterministic (packed 64-bit atomicMax: value, then LOWEST row on
// ties); scatter_mean adds floats with atomics like torch_scatter does, i.e. to ~1e-7 relative, order dependent.
#include "common.cuh"

namespace muvo {
namespace {

constexpr int kPillarBlock = 256;

// order-preserving map float -> uint32 (larger float <=> larger key); NaN sorts above +inf like torch's max
__device__ __forceinline__ uint32_t f32_key(float v) {
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_f32(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

template <typename IT>
__global__ void __launch_bounds__(kPillarBlock)
k_pillar_sum(const float* __restrict__ src, const IT* __restrict__ index, int64_t N, int F, int64_t M, float* __restrict__ sum,
             int32_t* __restrict__ count, int32_t* __restrict__ bad) {
  const int64_t t = (int64_t)blockIdx.x * kPillarBlock + threadIdx.x;
  if (t >= N * F) return;
  const int64_t n = t / F;
  const int f = (int)(t - n * F);
  const int64_t m = (int64_t)index[n];
  if (m < 0 || m >= M) { if (f == 0) atomicExch(bad, 1); return; }
  atomicAdd(sum + m * F + f, src[t]);
  if (f == 0) atomicAdd(count + m, 1);
}

__global__ void __launch_bounds__(kPillarBlock)
k_pillar_mean_finish(float* __restrict__ out, const int32_t* __restrict__ count, int64_t M, int F) {
  const int64_t t = (int64_t)blockIdx.x * kPillarBlock + threadIdx.x;
  if (t >= M * F) return;
  const int c = count[t / F];
  out[t] = out[t] / (float)(c > 1 ? c : 1);
}

template <typename IT>
__global__ void __launch_bounds__(kPillarBlock)
k_pillar_mean_bwd(const float* __restrict__ gout, const IT* __restrict__ index, const int32_t* __restrict__ count, int64_t N, int F,
                  float* __restrict__ gsrc) {
  const int64_t t = (int64_t)blockIdx.x * kPillarBlock + threadIdx.x;
  if (t >= N * F) return;
  const int64_t n = t / F;
  const int f = (int)(t - n * F);
  const int64_t m = (int64_t)index[n];
  const int c = count[m];
  gsrc[t] = gout[m * F + f] / (float)(c > 1 ? c : 1);
}

template <typename IT>
__global__ void __launch_bounds__(kPillarBlock)
k_pillar_max(const float* __restrict__ src, const IT* __restrict__ index, int64_t N, int F, int64_t M,
             unsigned long long* __restrict__ packed, int32_t* __restrict__ bad) {
  const int64_t t = (int64_t)blockIdx.x * kPillarBlock + threadIdx.x;
  if (t >= N * F) return;
  const int64_t n = t / F;
  const int f = (int)(t - n * F);
  const int64_t m = (int64_t)index[n];
  if (m < 0 || m >= M) { if (f == 0) atomicExch(bad, 1); return; }
  const unsigned long long w = ((unsigned long long)f32_key(src[t]) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)n);
  atomicMax(packed + m * F + f, w);
}

__global__ void __launch_bounds__(kPillarBlock)
k_pillar_max_finish(const unsigned long long* __restrict__ packed, int64_t M, int F, int64_t N, float* __restrict__ out,
                    int64_t* __restrict__ arg) {
  const int64_t t = (int64_t)blockIdx.x * kPillarBlock + threadIdx.x;
  if (t >= M * F) return;
  const unsigned long long w = packed[t];
  if (w == 0ull) { out[t] = 0.f; if (arg) arg[t] = N; return; }        // nothing scattered here
  out[t] = key_f32((uint32_t)(w >> 32));
  if (arg) arg[t] = (int64_t)(0xffffffffu - (uint32_t)w);
}

template <typename IT>
__global__ void __launch_bounds__(kPillarBlock)
k_pillar_max_bwd(const float* __restrict__ gout, const IT* __restrict__ index, const int64_t* __restrict__ arg, int64_t N, int F,
                 float* __restrict__ gsrc) {
  const int64_t t = (int64_t)blockIdx.x * kPillarBlock + threadIdx.x;
  if (t >= N * F) return;
  const int64_t n = t / F;
  const int f = (int)(t - n * F);
  const int64_t m = (int64_t)index[n];
  gsrc[t] = arg[m * F + f] == n ? gout[m * F + f] : 0.f;
}

static unsigned pillar_blocks(int64_t items) { return (unsigned)ceil_div64(items > 0 ? items : 1, kPillarBlock); }

}  // namespace
}  // namespace muvo

using namespace muvo;

extern "C" {

int muvo_pillar_workspace_bytes(int64_t n_out, int32_t n_feat, size_t* bytes_out_h) {
  if (!bytes_out_h) return MUVO_E_NULL;
  if (n_out < 0 || n_feat <= 0) return MUVO_E_ARG;
  *bytes_out_h = align_up((size_t)n_out * n_feat * 8, 256) + 256;       // packed words of scatter_max + the error flag
  return MUVO_OK;
}

int muvo_pillar_scatter_mean(const float* src, const void* index, int32_t index_dtype, int64_t n_src, int32_t n_feat, int64_t n_out,
                             float* out, int32_t* count_out, void* ws, size_t ws_bytes, int32_t* bad_index_flag, void* stream) {
  if (n_src < 0 || n_feat <= 0 || n_out < 0) return MUVO_E_ARG;
  if (n_src >= ((int64_t)1 << 32)) return MUVO_E_SHAPE;
  if (n_out == 0) return MUVO_OK;
  if (!out || !count_out || !bad_index_flag) return MUVO_E_NULL;
  if (n_src > 0 && (!src || !index)) return MUVO_E_NULL;
  (void)ws; (void)ws_bytes;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(out, 0, (size_t)n_out * n_feat * sizeof(float), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(count_out, 0, (size_t)n_out * sizeof(int32_t), st);
  if (e != cudaSuccess) return (int)e;
  prof_mark("<pillar>", st);
  if (n_src > 0) {
    if (index_dtype == MUVO_I64)
      k_pillar_sum<int64_t><<<pillar_blocks(n_src * n_feat), kPillarBlock, 0, st>>>(src, (const int64_t*)index, n_src, n_feat, n_out, out, count_out, bad_index_flag);
    else if (index_dtype == MUVO_I32)
      k_pillar_sum<int32_t><<<pillar_blocks(n_src * n_feat), kPillarBlock, 0, st>>>(src, (const int32_t*)index, n_src, n_feat, n_out, out, count_out, bad_index_flag);
    else return MUVO_E_ARG;
    MUVO_AFTER_LAUNCH("k_pillar_sum", st);
  }
  k_pillar_mean_finish<<<pillar_blocks(n_out * n_feat), kPillarBlock, 0, st>>>(out, count_out, n_out, n_feat);
  MUVO_AFTER_LAUNCH("k_pillar_mean_finish", st);
  return MUVO_OK;
}

int muvo_pillar_scatter_mean_bwd(const float* grad_out, const void* index, int32_t index_dtype, const int32_t* count, int64_t n_src,
                                 int32_t n_feat, float* grad_src, void* stream) {
  if (n_src < 0 || n_feat <= 0) return MUVO_E_ARG;
  if (n_src == 0) return MUVO_OK;
  if (!grad_out || !index || !count || !grad_src) return MUVO_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  if (index_dtype == MUVO_I64)
    k_pillar_mean_bwd<int64_t><<<pillar_blocks(n_src * n_feat), kPillarBlock, 0, st>>>(grad_out, (const int64_t*)index, count, n_src, n_feat, grad_src);
  else if (index_dtype == MUVO_I32)
    k_pillar_mean_bwd<int32_t><<<pillar_blocks(n_src * n_feat), kPillarBlock, 0, st>>>(grad_out, (const int32_t*)index, count, n_src, n_feat, grad_src);
  else return MUVO_E_ARG;
  MUVO_AFTER_LAUNCH("k_pillar_mean_bwd", st);
  return MUVO_OK;
}

int muvo_pillar_scatter_max(const float* src, const void* index, int32_t index_dtype, int64_t n_src, int32_t n_feat, int64_t n_out,
                            float* out, int64_t* arg_out, void* ws, size_t ws_bytes, int32_t* bad_index_flag, void* stream) {
  if (n_src < 0 || n_feat <= 0 || n_out < 0) return MUVO_E_ARG;
  if (n_src >= ((int64_t)1 << 32) - 1) return MUVO_E_SHAPE;           // the row index shares a 64-bit word with the value
  if (n_out == 0) return MUVO_OK;
  if (!out || !ws || !bad_index_flag) return MUVO_E_NULL;
  if (n_src > 0 && (!src || !index)) return MUVO_E_NULL;
  if (ws_bytes < (size_t)n_out * n_feat * 8) return MUVO_E_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(ws) % 8) return MUVO_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* packed = (unsigned long long*)ws;
  cudaError_t e = cudaMemsetAsync(packed, 0, (size_t)n_out * n_feat * 8, st);
  if (e != cudaSuccess) return (int)e;
  prof_mark("<pillar>", st);
  if (n_src > 0) {
    if (index_dtype == MUVO_I64)
      k_pillar_max<int64_t><<<pillar_blocks(n_src * n_feat), kPillarBlock, 0, st>>>(src, (const int64_t*)index, n_src, n_feat, n_out, packed, bad_index_flag);
    else if (index_dtype == MUVO_I32)
      k_pillar_max<int32_t><<<pillar_blocks(n_src * n_feat), kPillarBlock, 0, st>>>(src, (const int32_t*)index, n_src, n_feat, n_out, packed, bad_index_flag);
    else return MUVO_E_ARG;
    MUVO_AFTER_LAUNCH("k_pillar_max", st);
  }
  k_pillar_max_finish<<<pillar_blocks(n_out * n_feat), kPillarBlock, 0, st>>>(packed, n_out, n_feat, n_src, out, arg_out);
  MUVO_AFTER_LAUNCH("k_pillar_max_finish", st);
  return MUVO_OK;
}

int muvo_pillar_scatter_max_bwd(const float* grad_out, const void* index, int32_t index_dtype, const int64_t* arg, int64_t n_src,
                                int32_t n_feat, float* grad_src, void* stream) {
  if (n_src < 0 || n_feat <= 0) return MUVO_E_ARG;
  if (n_src == 0) return MUVO_OK;
  if (!grad_out || !index || !arg || !grad_src) return MUVO_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  if (index_dtype == MUVO_I64)
    k_pillar_max_bwd<int64_t><<<pillar_blocks(n_src * n_feat), kPillarBlock, 0, st>>>(grad_out, (const int64_t*)index, arg, n_src, n_feat, grad_src);
  else if (index_dtype == MUVO_I32)
    k_pillar_max_bwd<int32_t><<<pillar_blocks(n_src * n_feat), kPillarBlock, 0, st>>>(grad_out, (const int32_t*)index, arg, n_src, n_feat, grad_src);
  else return MUVO_E_ARG;
  MUVO_AFTER_LAUNCH("k_pillar_max_bwd", st);
  return MUVO_OK;
}

}  // extern "C"
