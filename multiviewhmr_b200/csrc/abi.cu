// ABI bookkeeping: version and the thread-local last-error buffer.
#include "mvhmr_common.cuh"

namespace mvhmr {
char *err_buf()
{
    static thread_local char buf[kErrBufLen] = {0};
    return buf;
}
}  // namespace mvhmr

extern "C" int mvhmr_abi_version(void) { return MVHMR_ABI_VERSION; }

extern "C" const char *mvhmr_last_error(void) { return mvhmr::err_buf(); }
