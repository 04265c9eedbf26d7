// Version / error plumbing of the C ABI (include/icd_b200.h).
#include "common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

void icd_set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" {
int icd_version(void) { return ICD_B200_ABI_VERSION; }
const char* icd_last_error_string(void) { return g_err; }
int icd_sizeof_att_desc(void) { return (int)sizeof(icd_att_desc_t); }
int icd_sizeof_base_desc(void) { return (int)sizeof(icd_base_desc_t); }
int icd_sizeof_beam_desc(void) { return (int)sizeof(icd_beam_desc_t); }
}
