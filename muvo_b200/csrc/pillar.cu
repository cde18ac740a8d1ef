// "Next" row N4, last part: the two torch_scatter reductions of the PointPillar encoder (muvo/models/common.py:698-761,
// POINT_PILLAR.ENABLED, off by default): scatter_mean(xyz, inverse_indices, dim=0) (:731) and
// scatter_max(feat, inverse_indices, dim=0)[0] (:703), with the gradients autograd needs.
//
// torch_scatter is not vendored in the reference and not installed in this image; the semantics restated here are its
// documented ones: out[m] = mean / max over {n : index[n] == m}; rows that receive nothing are 0; the mean divides by
// max(count, 1); scatter_max also returns arg[m, f] = a source row that attains the maximum (N for empty rows).
// Sizes are small (1e5 points x 3..32 features), so these are latency-sized kernels; what matters is one launch per
// reduction and no host round trip.  Both are deterministic: scatter_max is a packed 64-bit atomicMax (value, then LOWEST
// row on ties); scatter_mean accumulates in 64-bit FIXED POINT (integer atomics are associative, float atomics are not):
// every value is scaled by a power of two chosen from max|src| and N so that the sum of N such values cannot overflow,
// which keeps 2^-38 of the largest magnitude or better -- finer than the float32 result it is rounded to at the end
// (torch_scatter's float atomics are order dependent in the last bits).  NaN / Inf inputs propagate through side flags.
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

// max |src| over the finite values, as float bits (non-negative floats order like their bit patterns)
__global__ void __launch_bounds__(kPillarBlock)
k_pillar_absmax(const float* __restrict__ src, int64_t n, uint32_t* __restrict__ absmax_bits) {
  uint32_t m = 0u;
  for (int64_t t = (int64_t)blockIdx.x * kPillarBlock + threadIdx.x; t < n; t += (int64_t)gridDim.x * kPillarBlock) {
    const uint32_t u = __float_as_uint(src[t]) & 0x7fffffffu;
    if (u < 0x7f800000u && u > m) m = u;
  }
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m) atomicMax(absmax_bits, m);
}

// log2 of the fixed-point scale: |v| * 2^shift < 2^(61 - ceil(log2(N + 1))) for every finite v, so N of them cannot overflow
__device__ __forceinline__ int pillar_shift(uint32_t absmax_bits, int64_t N) {
  const int e_max = (int)(absmax_bits >> 23) - 126;                 // |v| < 2^e_max  (denormals / zero: e_max = -126)
  const int lgn = 64 - __clzll((unsigned long long)N);              // N < 2^lgn
  int sh = 61 - lgn - e_max;
  return sh > 126 ? 126 : sh;                                       // (the scale itself must be a finite double: always true here)
}

template <typename IT>
__global__ void __launch_bounds__(kPillarBlock)
k_pillar_sum(const float* __restrict__ src, const IT* __restrict__ index, int64_t N, int F, int64_t M,
             unsigned long long* __restrict__ sum, uint32_t* __restrict__ special, int32_t* __restrict__ count,
             const uint32_t* __restrict__ absmax_bits, int32_t* __restrict__ bad) {
  const int64_t t = (int64_t)blockIdx.x * kPillarBlock + threadIdx.x;
  if (t >= N * F) return;
  const int64_t n = t / F;
  const int f = (int)(t - n * F);
  const int64_t m = (int64_t)index[n];
  if (m < 0 || m >= M) { if (f == 0) atomicExch(bad, 1); return; }
  const float v = src[t];
  if (f == 0) atomicAdd(count + m, 1);
  if ((__float_as_uint(v) & 0x7fffffffu) >= 0x7f800000u) {          // NaN / +Inf / -Inf: remembered per output element
    atomicOr(special + m * F + f, v != v ? 1u : (v > 0.f ? 2u : 4u));
    return;
  }
  const long long q = __double2ll_rn(ldexp((double)v, pillar_shift(*absmax_bits, N)));
  atomicAdd(sum + m * F + f, (unsigned long long)q);                // two's complement: signed sums through the unsigned atomic
}

__global__ void __launch_bounds__(kPillarBlock)
k_pillar_mean_finish(const unsigned long long* __restrict__ sum, const uint32_t* __restrict__ special, float* __restrict__ out,
                     const int32_t* __restrict__ count, const uint32_t* __restrict__ absmax_bits, int64_t N, int64_t M, int F) {
  const int64_t t = (int64_t)blockIdx.x * kPillarBlock + threadIdx.x;
  if (t >= M * F) return;
  const int c = count[t / F];
  const uint32_t sp = special[t];
  float r;
  if (sp) r = (sp & 1u) || ((sp & 2u) && (sp & 4u)) ? __int_as_float(0x7fc00000) : ((sp & 2u) ? __int_as_float(0x7f800000) : __int_as_float(0xff800000));
  else r = (float)(ldexp((double)(long long)sum[t], -pillar_shift(*absmax_bits, N)) / (double)(c > 1 ? c : 1));
  out[t] = r;
}

template <typename IT>
__global__ void __launch_bounds__(kPillarBlock)
k_pillar_mean_bwd(const float* __restrict__ gout, const IT* __restrict__ index, const int32_t* __restrict__ count, int64_t N, int F,
                  int64_t M, float* __restrict__ gsrc) {
  const int64_t t = (int64_t)blockIdx.x * kPillarBlock + threadIdx.x;
  if (t >= N * F) return;
  const int64_t n = t / F;
  const int f = (int)(t - n * F);
  const int64_t m = (int64_t)index[n];
  if (m < 0 || m >= M) { gsrc[t] = 0.f; return; }                    // out-of-range rows were skipped by the forward
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
                 int64_t M, float* __restrict__ gsrc) {
  const int64_t t = (int64_t)blockIdx.x * kPillarBlock + threadIdx.x;
  if (t >= N * F) return;
  const int64_t n = t / F;
  const int f = (int)(t - n * F);
  const int64_t m = (int64_t)index[n];
  if (m < 0 || m >= M) { gsrc[t] = 0.f; return; }                    // out-of-range rows were skipped by the forward
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
  // scatter_max: packed words (8 B / element); scatter_mean: fixed-point sums (8 B) + NaN/Inf flags (4 B) + max|src|
  *bytes_out_h = align_up((size_t)n_out * n_feat * 8, 256) + align_up((size_t)n_out * n_feat * 4, 256) + 256;
  return MUVO_OK;
}

int muvo_pillar_scatter_mean(const float* src, const void* index, int32_t index_dtype, int64_t n_src, int32_t n_feat, int64_t n_out,
                             float* out, int32_t* count_out, void* ws, size_t ws_bytes, int32_t* bad_index_flag, void* stream) {
  if (n_src < 0 || n_feat <= 0 || n_out < 0) return MUVO_E_ARG;
  if (n_src >= ((int64_t)1 << 32)) return MUVO_E_SHAPE;
  if (n_out == 0) return MUVO_OK;
  if (!out || !count_out || !bad_index_flag || !ws) return MUVO_E_NULL;
  if (n_src > 0 && (!src || !index)) return MUVO_E_NULL;
  if (reinterpret_cast<uintptr_t>(ws) % 8) return MUVO_E_ALIGN;
  const size_t sum_bytes = align_up((size_t)n_out * n_feat * 8, 256), sp_bytes = align_up((size_t)n_out * n_feat * 4, 256);
  if (ws_bytes < sum_bytes + sp_bytes + 256) return MUVO_E_WORKSPACE;
  unsigned long long* sum = (unsigned long long*)ws;
  uint32_t* special = (uint32_t*)((char*)ws + sum_bytes);
  uint32_t* absmax = (uint32_t*)((char*)ws + sum_bytes + sp_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(ws, 0, sum_bytes + sp_bytes + 256, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(count_out, 0, (size_t)n_out * sizeof(int32_t), st);
  if (e != cudaSuccess) return ::muvo::cuda_fail(e);
  prof_mark("<pillar>", st);
  if (n_src > 0) {
    k_pillar_absmax<<<kNumSMsB200 * 4, kPillarBlock, 0, st>>>(src, n_src * n_feat, absmax);
    MUVO_AFTER_LAUNCH("k_pillar_absmax", st);
    if (index_dtype == MUVO_I64)
      k_pillar_sum<int64_t><<<pillar_blocks(n_src * n_feat), kPillarBlock, 0, st>>>(src, (const int64_t*)index, n_src, n_feat, n_out, sum, special, count_out, absmax, bad_index_flag);
    else if (index_dtype == MUVO_I32)
      k_pillar_sum<int32_t><<<pillar_blocks(n_src * n_feat), kPillarBlock, 0, st>>>(src, (const int32_t*)index, n_src, n_feat, n_out, sum, special, count_out, absmax, bad_index_flag);
    else return MUVO_E_ARG;
    MUVO_AFTER_LAUNCH("k_pillar_sum", st);
  }
  k_pillar_mean_finish<<<pillar_blocks(n_out * n_feat), kPillarBlock, 0, st>>>(sum, special, out, count_out, absmax, n_src, n_out, n_feat);
  MUVO_AFTER_LAUNCH("k_pillar_mean_finish", st);
  return MUVO_OK;
}

int muvo_pillar_scatter_mean_bwd(const float* grad_out, const void* index, int32_t index_dtype, const int32_t* count, int64_t n_src,
                                 int32_t n_feat, int64_t n_out, float* grad_src, void* stream) {
  if (n_src < 0 || n_feat <= 0 || n_out < 0) return MUVO_E_ARG;
  if (n_src == 0) return MUVO_OK;
  if (!grad_out || !index || !count || !grad_src) return MUVO_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  if (index_dtype == MUVO_I64)
    k_pillar_mean_bwd<int64_t><<<pillar_blocks(n_src * n_feat), kPillarBlock, 0, st>>>(grad_out, (const int64_t*)index, count, n_src, n_feat, n_out, grad_src);
  else if (index_dtype == MUVO_I32)
    k_pillar_mean_bwd<int32_t><<<pillar_blocks(n_src * n_feat), kPillarBlock, 0, st>>>(grad_out, (const int32_t*)index, count, n_src, n_feat, n_out, grad_src);
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
  if (e != cudaSuccess) return ::muvo::cuda_fail(e);
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
                                int32_t n_feat, int64_t n_out, float* grad_src, void* stream) {
  if (n_src < 0 || n_feat <= 0 || n_out < 0) return MUVO_E_ARG;
  if (n_src == 0) return MUVO_OK;
  if (!grad_out || !index || !arg || !grad_src) return MUVO_E_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  if (index_dtype == MUVO_I64)
    k_pillar_max_bwd<int64_t><<<pillar_blocks(n_src * n_feat), kPillarBlock, 0, st>>>(grad_out, (const int64_t*)index, arg, n_src, n_feat, n_out, grad_src);
  else if (index_dtype == MUVO_I32)
    k_pillar_max_bwd<int32_t><<<pillar_blocks(n_src * n_feat), kPillarBlock, 0, st>>>(grad_out, (const int32_t*)index, arg, n_src, n_feat, n_out, grad_src);
  else return MUVO_E_ARG;
  MUVO_AFTER_LAUNCH("k_pillar_max_bwd", st);
  return MUVO_OK;
}

}  // extern "C"
