// C-ABI plumbing shared by every entry point of libmhb200.so: status / error text helpers.
#include <stdarg.h>

#include "common.cuh"

namespace mhb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int32_t cuda_status(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return MHB_OK;
    set_error("%s: CUDA error %d (%s)", what, static_cast<int>(e), cudaGetErrorString(e));
    return static_cast<int32_t>(e);
}

}  // namespace mhb

extern "C" int32_t mhb_abi_version(void) { return MHB_ABI_VERSION; }

extern "C" const char* mhb_last_error(void) { return mhb::g_err; }

extern "C" int64_t mhb_n_windows(int64_t series_len, int32_t wsize, int32_t wstep) {
    if (wsize < 1 || wstep < 1) return 0;
    return mhb::n_windows_host(series_len, wsize, wstep);
}
