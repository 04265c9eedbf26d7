// bf16 tensor-core (tcgen05 + TMA) contraction tier — placeholder until the kernel lands.
#include "common.cuh"

extern "C" int icd_has_tensor_core_gemm(void) { return 0; }

int icd_gemm_tc_launch(const icd_gemm_desc_t* d, cudaStream_t s) {
    (void)d; (void)s;
    icd_set_error("gemm: ICD_PREC_BF16 requested but the tcgen05 tier is not built in this library");
    return -2;
}
