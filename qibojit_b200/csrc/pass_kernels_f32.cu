// complex64 instantiation of the multi-gate tile pass kernel (pass_device.cuh), in its own translation
// unit so that it compiles in parallel with the complex128 one in pass_kernels.cu.
#include "pass_device.cuh"

namespace qj {

int launch_k_pass_f32(int device, unsigned grid, int threads, size_t smem, cudaStream_t stream, void *state,
                      const void *geom, const void *tables, const void *pp) {
    static bool configured[kMaxDevices] = {false};   // the attribute is per device, not per process
    if (!configured[device]) {
        QJ_CUDA_OK(cudaFuncSetAttribute(k_pass<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10));
        configured[device] = true;
    }
    k_pass<float><<<grid, threads, smem, stream>>>(
        reinterpret_cast<Cx<float> *>(state), *reinterpret_cast<const PassGeom *>(geom),
        reinterpret_cast<const Cx<float> *>(tables), *reinterpret_cast<const ProgParam *>(pp));
    return QJ_OK;
}

}  // namespace qj
