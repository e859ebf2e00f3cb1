// ABI bookkeeping for libtrb.so.
#include "trb_common.cuh"

namespace trb {
thread_local int g_last_cuda_error = 0;
}

extern "C" int trb_abi_version(void) { return TRB_ABI_VERSION; }

extern "C" int trb_last_cuda_error(void) { return trb::g_last_cuda_error; }

// sizeof() of the structs that cross the ABI, so that a binding can refuse a stale library.
extern "C" int trb_abi_struct_size(int which) {
  switch (which) {
    case 0: return (int)sizeof(trb_view);
    case 1: return (int)sizeof(trb_shade_config);
    case 2: return (int)sizeof(trb_render_config);
    case 3: return (int)sizeof(trb_uv_texture);
    case 4: return (int)sizeof(trb_peer_sum);
    case 5: return (int)sizeof(trb_render_extras);
    default: return -1;
  }
}

extern "C" const char* trb_status_string(int status) {
  switch (status) {
    case TRB_OK: return "ok";
    case TRB_ERR_BAD_ARG: return "bad argument";
    case TRB_ERR_K_TOO_LARGE: return "faces_per_pixel exceeds 150";
    case TRB_ERR_WORKSPACE: return "workspace too small";
    case TRB_ERR_CUDA: return cudaGetErrorString((cudaError_t)trb::g_last_cuda_error);
    default: return "unknown status";
  }
}
