// Library-level glue of libradarslam_b200.so: version and thread-local error text.
#include <stdarg.h>
#include <string.h>
#include "rs_common.cuh"

static thread_local char g_err[512] = "";

void rs_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int rs_version(void) { return 100; }

extern "C" const char* rs_last_error(void) { return g_err; }
