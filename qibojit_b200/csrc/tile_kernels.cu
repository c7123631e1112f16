// tile_kernels.cu -- shared-memory staged dense k-target kernel (placeholder: routing off).
#include "common.cuh"
namespace qj {
bool tile_kernel_applies(const qj_handle *, const GateCall &) { return false; }
int launch_dense_tile(qj_handle *, const GateCall &) { return fail(QJ_ERR_UNSUPPORTED, "tile kernel not built"); }
}  // namespace qj
