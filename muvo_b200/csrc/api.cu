// ABI version + error strings.
#include "common.cuh"

extern "C" {

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
