// capi.cu -- ABI version and error strings of libmonodetr_b200.so (see include/monodetr_b200.h).
#include <cuda_runtime.h>

#include "../../include/monodetr_b200.h"

extern "C" {

int mdb_abi_version(void) { return 2; }

// Reproducible-accumulation mode (process-wide, like the arithmetic mode): read by the two large scatter / accumulate sites,
// the MSDeformAttn value gradient (msda.cu) and the split-K weight gradient (conv_gemm.cu).
static int g_deterministic = 0;
int mdb_set_deterministic(int on) { g_deterministic = on ? 1 : 0; return 0; }
int mdb_get_deterministic(void) { return g_deterministic; }

const char* mdb_error_string(int code) {
    if (code == 0) return "ok";
    if (code == MDB_EINVAL) return "monodetr_b200: invalid argument (size, null or misaligned pointer)";
    if (code == MDB_EUNSUPPORTED) return "monodetr_b200: shape not supported by the sm_100a kernels";
    if (code == MDB_EWORKSPACE) return "monodetr_b200: scratch workspace missing or too small (mdb_set_workspace)";
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "monodetr_b200: unknown error";
}

}  // extern "C"
