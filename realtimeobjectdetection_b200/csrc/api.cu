// api.cu -- library-wide pieces of the C ABI (include/rtod.h): version and thread-local error text
#include "common.cuh"

namespace rtod {
char* last_error_buffer() {
    static thread_local char buf[kErrBuf] = "";
    return buf;
}
}  // namespace rtod

extern "C" int rtod_abi_version(void) { return RTOD_ABI_VERSION; }
extern "C" const char* rtod_last_error(void) { return rtod::last_error_buffer(); }

// ---- rtod_sm_clock_probe: SM cycles per wall-clock window, sampled by one resident thread -------------
namespace {
__global__ void sm_clock_probe_kernel(float* out_mhz, int samples, unsigned long long interval_ns) {
    for (int i = 0; i < samples; ++i) {
        const unsigned long long t0 = rtod::global_timer_ns();
        const long long c0 = clock64();
        unsigned long long t1 = t0;
        while (t1 - t0 < interval_ns) {
            __nanosleep(1000);
            t1 = rtod::global_timer_ns();
        }
        const long long c1 = clock64();
        out_mhz[i] = (float)((double)(c1 - c0) * 1000.0 / (double)(t1 - t0));
    }
}
}  // namespace

extern "C" int rtod_sm_clock_probe(float* out_mhz, int samples, int interval_us, void* stream) {
    if (!out_mhz || samples < 1 || samples > 100000 || interval_us < 1 || interval_us > 100000)
        return rtod::fail(RTOD_ERR_BAD_ARG, "rtod_sm_clock_probe: samples in [1,1e5], interval_us in [1,1e5]");
    sm_clock_probe_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(out_mhz, samples, (unsigned long long)interval_us * 1000ull);
    RTOD_LAUNCH_OK("sm_clock_probe_kernel");
    return RTOD_OK;
}
