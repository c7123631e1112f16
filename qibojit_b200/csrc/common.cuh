// common.cuh -- shared device/host helpers of the qibojit_b200 CUDA library (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <string>

#include "../../include/qibojit_b200.h"

namespace qj {

// ------------------------------------------------------------------ error plumbing
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);

#define QJ_CUDA_OK(expr)                                                                  \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess)                                                            \
            return ::qj::fail(QJ_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

#define QJ_REQUIRE(cond, msg)                                      \
    do {                                                           \
        if (!(cond)) return ::qj::fail(QJ_ERR_INVALID, (msg));      \
    } while (0)

}  // namespace qj

// ------------------------------------------------------------------ the handle
struct qj_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    int route = 0;
    int64_t launches = 0;
    // small device scratch for reductions (partials, norm, argmax): never state-sized
    double *scratch = nullptr;
    size_t scratch_doubles = 0;
    // device arena + pinned mirror for gate matrices too large for kernel parameters
    void *gate_dev = nullptr;
    void *gate_pin = nullptr;
    size_t gate_slot_bytes = 0;
    int gate_slots = 0;
    int gate_next = 0;
    cudaEvent_t *gate_done = nullptr;
};

namespace qj {

// Every entry point runs on its handle's device whatever device the calling thread had current
// (the reference drives all devices from joblib threads of one process, gpu.py:688-694) and
// leaves the caller's current device as it found it.
constexpr int kMaxDevices = 64;
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(const qj_handle *h) {
        if (h && cudaGetDevice(&prev) == cudaSuccess && prev != h->device) {
            cudaSetDevice(h->device);
            switched = true;
        }
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

constexpr int kMaxPos = QJ_MAX_QUBITS;
constexpr int kMaxDirectTargets = 5;  // register ("direct") kernels: 2^5 amplitudes per thread

// Geometry of one gate pass, in units of the vector a thread moves per access.
//   group g in [0, 2^nfree)  --expand-->  index with zeros at pos[], | cmask  = base
//   element e of the tuple lives at base + off[e]
struct GateGeom {
    int nfree;
    int npos;
    int pos[kMaxPos];
    int64_t cmask;
    int64_t off[1 << kMaxDirectTargets];
};

// Row-major complex matrix carried in kernel parameters (constant bank operands).
template <typename T, int DIM>
struct CMat {
    T v[2 * DIM * DIM];
};

template <typename T>
struct __align__(2 * sizeof(T)) Cx {  // 16-byte (complex128) / 8-byte (complex64) vector accesses
    T re, im;
};

// --------------------------------------------------------------------------- vectors
// One memory access of a thread: V complex amplitudes.
template <typename T, int V>
struct VecOf;
template <>
struct VecOf<double, 1> {
    using type = double2;
};
template <>
struct VecOf<float, 1> {
    using type = float2;
};
template <>
struct VecOf<float, 2> {
    using type = float4;
};

template <typename T, int V>
struct Amp {  // V complex numbers in registers
    T re[V], im[V];
};

__device__ __forceinline__ Amp<double, 1> ld_amp(const double2 *p) {
    double2 v = *p;
    Amp<double, 1> a;
    a.re[0] = v.x; a.im[0] = v.y;
    return a;
}
__device__ __forceinline__ Amp<float, 1> ld_amp(const float2 *p) {
    float2 v = *p;
    Amp<float, 1> a;
    a.re[0] = v.x; a.im[0] = v.y;
    return a;
}
__device__ __forceinline__ Amp<float, 2> ld_amp(const float4 *p) {
    float4 v = *p;
    Amp<float, 2> a;
    a.re[0] = v.x; a.im[0] = v.y; a.re[1] = v.z; a.im[1] = v.w;
    return a;
}
__device__ __forceinline__ void st_amp(double2 *p, const Amp<double, 1> &a) {
    *p = make_double2(a.re[0], a.im[0]);
}
__device__ __forceinline__ void st_amp(float2 *p, const Amp<float, 1> &a) {
    *p = make_float2(a.re[0], a.im[0]);
}
__device__ __forceinline__ void st_amp(float4 *p, const Amp<float, 2> &a) {
    *p = make_float4(a.re[0], a.im[0], a.re[1], a.im[1]);
}

// insert a zero bit at every position of geo.pos (ascending)
__device__ __forceinline__ int64_t expand_index(int64_t g, const GateGeom &geo) {
#pragma unroll 1
    for (int j = 0; j < geo.npos; j++) {
        const int p = geo.pos[j];
        const int64_t lo = g & ((int64_t(1) << p) - 1);
        g = ((g >> p) << (p + 1)) | lo;
    }
    return g;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum, result valid in thread 0; fixed reduction tree (deterministic)
template <int THREADS>
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double s_part[THREADS / 32];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 32) {
        r = (threadIdx.x < THREADS / 32) ? s_part[threadIdx.x] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

// ------------------------------------------------------------- launch entry points
// (implemented in gate_kernels.cu / tile_kernels.cu / ops_kernels.cu; dtype-dispatched)
struct GateCall {
    void *state;
    int dtype;
    int nqubits;
    int ntargets;
    int ncontrols;
    int tbits[QJ_MAX_TARGETS];   // index bit addressed by matrix-index bit u
    int cbits[QJ_MAX_QUBITS];    // control index bits
    const void *gate;            // host, row-major 2^k x 2^k in the state dtype (may be null)
};

enum SpecialOp { OP_X = 1, OP_Y = 2, OP_Z = 3, OP_ZPOW = 4, OP_SWAP = 5, OP_FSIM = 6, OP_PHASE = 7 };

int launch_dense_direct(qj_handle *h, const GateCall &c);
int launch_special(qj_handle *h, const GateCall &c, int op);
int launch_dense_generic(qj_handle *h, const GateCall &c);   // k > kMaxDirectTargets
int launch_dense_tile(qj_handle *h, const GateCall &c);      // smem/TMA staged
bool tile_kernel_applies(const qj_handle *h, const GateCall &c);

int stage_gate_matrix(qj_handle *h, const void *host, size_t bytes, void **dev_out, int *slot);
void gate_slot_release(qj_handle *h, int slot);

}  // namespace qj
