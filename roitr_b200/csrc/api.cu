// Error reporting and ABI version of libroitr_b200.
#include <stdarg.h>
#include <stdio.h>

#include "../../include/roitr_b200.h"
#include "common.cuh"

static thread_local char g_err[512] = "";

void roitr_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* roitr_last_error(void) { return g_err; }
extern "C" int roitr_abi_version(void) { return 1; }
