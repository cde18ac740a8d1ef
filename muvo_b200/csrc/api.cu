// ABI version + error strings.
#include "common.cuh"

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
  if (e != cudaSuccess) return (int)e;
  // marks whose name starts with '<' are baselines (recorded right before the first launch of an entry point):
  // they are not reported, but the next kernel's time is measured from them, so host-side gaps are excluded.
  int n = 0;
  for (int i = 1; i < p.n && n < capacity; ++i) {
    if (p.name[i][0] == '<') continue;
    float ms = 0.f;
    e = cudaEventElapsedTime(&ms, p.ev[i - 1], p.ev[i]);
    if (e != cudaSuccess) return (int)e;
    if (ms_out_h) ms_out_h[n] = ms;
    if (names_out_h) names_out_h[n] = p.name[i];
    ++n;
  }
  *n_out_h = n;
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
