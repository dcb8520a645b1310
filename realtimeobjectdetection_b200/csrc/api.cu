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
