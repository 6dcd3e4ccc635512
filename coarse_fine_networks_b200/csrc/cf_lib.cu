// Library-level state of libcfnet_b200: error string, ABI version, launch counter.
#include <stdarg.h>
#include "cf_common.cuh"
#include "../../include/cfnet_b200.h"

static thread_local char g_err[512] = "";
unsigned long long g_cf_launches = 0;

void cf_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" {
const char* cf_last_error(void) { return g_err; }
int cf_abi_version(void) { return 1; }
unsigned long long cf_launch_count(void) { return g_cf_launches; }
}
