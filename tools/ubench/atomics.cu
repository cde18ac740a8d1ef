// Microbenchmark: throughput of the atomic / store patterns the point pass can choose from (B200, sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomics atomics.cu ; run: ./atomics
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;

enum Op { ATOM64 = 0, RED64, ST64, ATOM32, RED32, SMEM32, SMEM64, LD64 };
enum Pat { COAL = 0, STRIDE4, PAIRS, RANDOM, RUN2_STRIDE, QUADS };

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int OP, int PAT>
__global__ void __launch_bounds__(256) k(u64* tab, uint32_t mask_words, int iters, u64* sink) {
  extern __shared__ u64 sm[];
  const uint32_t gt = blockIdx.x * 256 + threadIdx.x, lane = threadIdx.x & 31, gw = gt >> 5;
  u64 acc = 0;
  if (OP == SMEM32 || OP == SMEM64) { for (int i = threadIdx.x; i < 4096; i += 256) sm[i] = 0; __syncthreads(); }
#pragma unroll 4
  for (int it = 0; it < iters; ++it) {
    const uint32_t wbase = hash32(gw * 7919u + it) ;   // a warp-level base (moves every iteration)
    uint32_t idx;
    if (PAT == COAL) idx = wbase * 32u + lane;
    else if (PAT == STRIDE4) idx = wbase * 128u + lane * 4u;
    else if (PAT == PAIRS) idx = wbase * 16u + (lane >> 1);
    else if (PAT == QUADS) idx = wbase * 8u + (lane >> 2);
    else if (PAT == RUN2_STRIDE) idx = wbase * 64u + (lane >> 1) * 3u;
    else idx = hash32(gt * 2654435761u + it * 40503u);
    idx &= mask_words;
    const u64 v = ((u64)hash32(gt + it * 977u) << 32) | gt;
    if (OP == ATOM64) { u64 o; asm volatile("atom.global.max.u64 %0, [%1], %2;" : "=l"(o) : "l"(tab + idx), "l"(v) : "memory"); acc += o; }
    else if (OP == RED64) { asm volatile("red.global.max.u64 [%0], %1;" ::"l"(tab + idx), "l"(v) : "memory"); }
    else if (OP == ST64) { asm volatile("st.global.cg.u64 [%0], %1;" ::"l"(tab + idx), "l"(v) : "memory"); }
    else if (OP == LD64) { u64 o; asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(o) : "l"(tab + idx) : "memory"); acc += o; }
    else if (OP == ATOM32) { uint32_t o; asm volatile("atom.global.max.u32 %0, [%1], %2;" : "=r"(o) : "l"((uint32_t*)tab + idx), "r"((uint32_t)(v >> 32)) : "memory"); acc += o; }
    else if (OP == RED32) { asm volatile("red.global.max.u32 [%0], %1;" ::"l"((uint32_t*)tab + idx), "r"((uint32_t)(v >> 32)) : "memory"); }
    else if (OP == SMEM32) { acc += atomicMax((uint32_t*)sm + (idx & 8191u), (uint32_t)(v >> 32)); }
    else if (OP == SMEM64) { acc += atomicMax(sm + (idx & 4095u), v); }
  }
  if (acc == 0x1234567u) sink[0] = acc;
}

template <int OP, int PAT>
static void run(const char* name, u64* tab, size_t words, u64* sink) {
  const int grid = 148 * 4, iters = 2048;
  cudaFuncSetAttribute(k<OP, PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP, PAT><<<grid, 256, 32768>>>(tab, (uint32_t)(words - 1), 64, sink);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<OP, PAT><<<grid, 256, 32768>>>(tab, (uint32_t)(words - 1), iters, sink);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double n = (double)grid * 256 * iters;
  printf("%-28s table %6.1f MB : %8.3f ms  %8.2f Gop/s  %6.3f lane-ops/clk/SM @1.965GHz  err=%d\n", name, words * 8 / 1e6, ms,
         n / ms / 1e6, n / (ms * 1e-3) / 148 / 1.965e9, (int)cudaGetLastError());
}

int main() {
  u64 *tab, *sink;
  const size_t words_small = (size_t)1 << 20;   // 8 MB (L2 resident)
  const size_t words_big = (size_t)1 << 27;     // 1 GB  (DRAM)
  cudaMalloc(&tab, words_big * 8); cudaMemset(tab, 0, words_big * 8); cudaMalloc(&sink, 64);
#define R(OP, PAT) run<OP, PAT>(#OP " " #PAT, tab, words_small, sink)
  R(ATOM64, COAL); R(ATOM64, STRIDE4); R(ATOM64, PAIRS); R(ATOM64, QUADS); R(ATOM64, RUN2_STRIDE); R(ATOM64, RANDOM);
  R(RED64, COAL); R(RED64, STRIDE4); R(RED64, PAIRS); R(RED64, QUADS); R(RED64, RUN2_STRIDE); R(RED64, RANDOM);
  R(ATOM32, COAL); R(ATOM32, STRIDE4); R(ATOM32, PAIRS); R(ATOM32, RANDOM);
  R(RED32, COAL); R(RED32, RANDOM);
  R(ST64, COAL); R(ST64, STRIDE4); R(ST64, PAIRS); R(ST64, RANDOM);
  R(LD64, COAL); R(LD64, STRIDE4); R(LD64, RANDOM);
  R(SMEM32, COAL); R(SMEM32, RANDOM); R(SMEM32, PAIRS); R(SMEM64, COAL); R(SMEM64, RANDOM);
  printf("--- DRAM-sized table (1 GB)\n");
  run<ATOM64, COAL>("ATOM64 COAL big", tab, words_big, sink);
  run<ATOM64, RANDOM>("ATOM64 RANDOM big", tab, words_big, sink);
  run<RED64, RANDOM>("RED64 RANDOM big", tab, words_big, sink);
  return 0;
}
