// ABI version + error strings.
#include "common.cuh"
#include <cstring>
#include <thread>
#include <vector>

namespace muvo {
Profile& profile_state() {
  static thread_local Profile p;
  return p;
}
}  // namespace muvo

extern "C" {

int muvo_profile_begin(void* stream) {
  muvo::Profile& p = muvo::profile_state();
  p.n = 0;
  p.on = true;
  muvo::prof_mark("<begin>", (cudaStream_t)stream);
  return p.n == 1 ? MUVO_OK : MUVO_E_ARG;
}

int muvo_profile_end(void* stream, int32_t capacity, float* ms_out_h, const char** names_out_h, int32_t* n_out_h) {
  muvo::Profile& p = muvo::profile_state();
  if (!p.on) return MUVO_E_ARG;
  p.on = false;
  if (!n_out_h) return MUVO_E_NULL;
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  if (e != cudaSuccess) return ::muvo::cuda_fail(e);
  // marks whose name starts with '<' are baselines (recorded right before the first launch of an entry point):
  // they are not reported, but the next kernel's time is measured from them, so host-side gaps are excluded.
  int n = 0;
  for (int i = 1; i < p.n && n < capacity; ++i) {
    if (p.name[i][0] == '<') continue;
    float ms = 0.f;
    e = cudaEventElapsedTime(&ms, p.ev[i - 1], p.ev[i]);
    if (e != cudaSuccess) return ::muvo::cuda_fail(e);
    if (ms_out_h) ms_out_h[n] = ms;
    if (names_out_h) names_out_h[n] = p.name[i];
    ++n;
  }
  *n_out_h = n;
  return MUVO_OK;
}

int muvo_host_copy(void* dst_h, const void* src_h, size_t bytes, int32_t n_threads) {
  if (bytes == 0) return MUVO_OK;
  if (!dst_h || !src_h) return MUVO_E_NULL;
  if (n_threads < 0) return MUVO_E_ARG;
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 1;
  size_t nt = n_threads > 0 ? (size_t)n_threads : (hw > 16 ? 16 : hw);
  const size_t min_chunk = (size_t)1 << 20;                       // below 1 MiB per thread the spawn costs more than it saves
  if (nt > (bytes + min_chunk - 1) / min_chunk) nt = (bytes + min_chunk - 1) / min_chunk;
  if (nt <= 1) { std::memcpy(dst_h, src_h, bytes); return MUVO_OK; }
  const size_t chunk = ((bytes + nt - 1) / nt + 4095) & ~(size_t)4095;
  std::vector<std::thread> th;
  th.reserve(nt - 1);
  char* d = (char*)dst_h;
  const char* s = (const char*)src_h;
  for (size_t t = 1; t < nt; ++t) {
    const size_t o = t * chunk;
    if (o >= bytes) break;
    const size_t n = bytes - o < chunk ? bytes - o : chunk;
    th.emplace_back([=] { std::memcpy(d + o, s + o, n); });
  }
  std::memcpy(d, s, chunk < bytes ? chunk : bytes);
  for (auto& t : th) t.join();
  return MUVO_OK;
}

int muvo_abi_version(void) { return MUVO_B200_ABI_VERSION; }

const char* muvo_strerror(int code) {
  switch (code) {
    case MUVO_OK: return "ok";
    case MUVO_E_NULL: return "required pointer is NULL";
    case MUVO_E_ARG: return "invalid argument or unsupported dtype";
    case MUVO_E_SHAPE: return "shape overflow";
    case MUVO_E_WORKSPACE: return "workspace too small";
    case MUVO_E_ALIGN: return "pointer not aligned as documented";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown muvo error";
}

}  // extern "C"
