// capi.cu -- extern "C" boundary (include/qibojit_b200.h): handle management, argument
// normalisation from the reference's kernel arguments, kernel routing.

#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace qj {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }
int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}

// ---------------------------------------------------------------- gate arena
// Matrices that do not fit kernel parameters (k >= 6) are staged through a small ring of
// pinned-host / device slots; a slot is reused only after the kernel that read it finished.
int stage_gate_matrix(qj_handle *h, const void *host, size_t bytes, void **dev_out, int *slot_out) {
    if (bytes > h->gate_slot_bytes) return fail(QJ_ERR_INVALID, "gate matrix larger than the staging slot");
    if (h->gate_dev == nullptr) {  // lazily: most circuits never need it
        QJ_CUDA_OK(cudaMalloc(&h->gate_dev, h->gate_slot_bytes * h->gate_slots));
        QJ_CUDA_OK(cudaMallocHost(&h->gate_pin, h->gate_slot_bytes * h->gate_slots));
    }
    const int slot = h->gate_next;
    h->gate_next = (h->gate_next + 1) % h->gate_slots;
    QJ_CUDA_OK(cudaEventSynchronize(h->gate_done[slot]));
    char *pin = static_cast<char *>(h->gate_pin) + size_t(slot) * h->gate_slot_bytes;
    char *dev = static_cast<char *>(h->gate_dev) + size_t(slot) * h->gate_slot_bytes;
    memcpy(pin, host, bytes);
    QJ_CUDA_OK(cudaMemcpyAsync(dev, pin, bytes, cudaMemcpyHostToDevice, h->stream));
    *dev_out = dev;
    *slot_out = slot;
    return QJ_OK;
}

void gate_slot_release(qj_handle *h, int slot) {
    if (slot >= 0) cudaEventRecord(h->gate_done[slot], h->stream);
}

namespace {

int check_common(qj_handle *h, void *state, int dtype, int nqubits) {
    QJ_REQUIRE(h != nullptr, "null handle");
    QJ_REQUIRE(state != nullptr, "null state pointer");
    QJ_REQUIRE(dtype == QJ_C64 || dtype == QJ_C128, "dtype must be QJ_C64 or QJ_C128");
    QJ_REQUIRE(nqubits >= 1 && nqubits <= QJ_MAX_QUBITS, "nqubits out of range");
    QJ_REQUIRE((reinterpret_cast<uintptr_t>(state) & 15) == 0 || nqubits == 0, "state must be 16-byte aligned");
    return QJ_OK;
}

// controls = qubits \ targets  (qubits: sorted bit positions of controls U targets)
int fill_controls(GateCall &c, const int32_t *qubits, int nactive) {
    c.ncontrols = 0;
    if (qubits == nullptr || nactive <= c.ntargets) return QJ_OK;
    QJ_REQUIRE(nactive <= QJ_MAX_QUBITS, "too many active qubits");
    int found = 0;
    for (int i = 0; i < nactive; i++) {
        bool is_target = false;
        for (int u = 0; u < c.ntargets; u++) is_target |= (c.tbits[u] == qubits[i]);
        if (is_target) { found++; continue; }
        QJ_REQUIRE(qubits[i] >= 0 && qubits[i] < c.nqubits, "control qubit out of range");
        c.cbits[c.ncontrols++] = qubits[i];
    }
    QJ_REQUIRE(found == c.ntargets, "`qubits` must contain every target bit");
    return QJ_OK;
}

int route_dense(qj_handle *h, const GateCall &c) {
    if (c.ntargets > kMaxDirectTargets) return launch_dense_generic(h, c);   // (the tile kernel stages at most 2^5 amplitudes per group)
    if (h->route == 2 && tile_kernel_applies(h, c)) return launch_dense_tile(h, c);
    // automatic: measured on B200 (profiles/r1_sweep_*.txt) the register kernels are at or above
    // the copy-bandwidth roofline for k <= 4 wherever the targets sit -- except complex64 gates that
    // act on index bit 0 (the two amplitudes of a 16-byte vector: the register kernel falls back to
    // 8-byte accesses), where the TMA-staged tile kernel wins: k = 2 on bits {0,1} 5.86 vs 4.31 TB/s,
    // k = 3 on {0,1,2} 3.81 vs 3.56, k = 4 on {0,1,2,3} 2.16 vs 1.96.  Those cases go to the tile kernel.
    if (h->route == 0 && c.dtype == QJ_C64 && c.ntargets >= 2 && c.ntargets <= 4 && c.ncontrols == 0) {
        bool bit0 = false;
        for (int u = 0; u < c.ntargets; u++) bit0 |= c.tbits[u] == 0;
        if (bit0 && tile_kernel_applies(h, c)) return launch_dense_tile(h, c);
    }
    return launch_dense_direct(h, c);
}

int one_target(qj_handle *h, void *state, int dtype, int nqubits, int m, const void *gate,
               const int32_t *qubits, int nactive, int op) {
    int rc = check_common(h, state, dtype, nqubits);
    if (rc) return rc;
    QJ_REQUIRE(m >= 0 && m < nqubits, "target bit out of range");
    GateCall c;
    memset(&c, 0, sizeof(c));
    c.state = state; c.dtype = dtype; c.nqubits = nqubits; c.ntargets = 1; c.tbits[0] = m; c.gate = gate;
    rc = fill_controls(c, qubits, nactive);
    if (rc) return rc;
    if (op == 0) {
        QJ_REQUIRE(gate != nullptr, "null gate matrix");
        return route_dense(h, c);
    }
    if (op == OP_ZPOW) QJ_REQUIRE(gate != nullptr, "null phase");
    return launch_special(h, c, op);
}

int two_target(qj_handle *h, void *state, int dtype, int nqubits, int m1, int m2, int swap_targets,
               const void *gate, const int32_t *qubits, int nactive, int op) {
    int rc = check_common(h, state, dtype, nqubits);
    if (rc) return rc;
    QJ_REQUIRE(nqubits >= 2, "two-qubit gate on a one-qubit register");
    QJ_REQUIRE(m1 >= 0 && m2 < nqubits && m1 < m2, "need 0 <= m1 < m2 < nqubits");
    GateCall c;
    memset(&c, 0, sizeof(c));
    c.state = state; c.dtype = dtype; c.nqubits = nqubits; c.ntargets = 2; c.gate = gate;
    // gates.py:119-122: matrix-index bit 0 addresses uk1 (= tk1 unless swap_targets)
    c.tbits[0] = swap_targets ? m2 : m1;
    c.tbits[1] = swap_targets ? m1 : m2;
    rc = fill_controls(c, qubits, nactive);
    if (rc) return rc;
    if (op == 0) {
        QJ_REQUIRE(gate != nullptr, "null gate matrix");
        return route_dense(h, c);
    }
    if (op == OP_FSIM) QJ_REQUIRE(gate != nullptr, "null fsim parameters");
    return launch_special(h, c, op);
}

}  // namespace
}  // namespace qj

using namespace qj;

extern "C" {

const char *qj_last_error(void) { return g_last_error.c_str(); }
const char *qj_version(void) { return "qibojit_b200 0.1.0 (sm_100a)"; }

int qj_create(int device, void *stream, qj_handle **out) {
    QJ_REQUIRE(out != nullptr, "null output pointer");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(QJ_ERR_NODEVICE, "no CUDA device available (this library has no CPU fallback)");
    QJ_REQUIRE(device >= 0 && device < ndev && device < kMaxDevices, "device ordinal out of range");
    qj_handle *h = new qj_handle();
    // allocate on `device`, give the caller's current device back on return
    int prev_device = -1;
    cudaGetDevice(&prev_device);
    struct Restore {
        int prev;
        ~Restore() { if (prev >= 0) cudaSetDevice(prev); }
    } restore{prev_device};
    QJ_CUDA_OK(cudaSetDevice(device));
    h->device = device;
    cudaDeviceProp prop;
    QJ_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    // NULL is the CUDA default stream -- what torch uses unless told otherwise -- so work
    // enqueued here is ordered with the caller's copies and reads of the state.
    h->stream = static_cast<cudaStream_t>(stream);
    h->scratch_doubles = size_t(2) << 20;  // 16 MiB
    QJ_CUDA_OK(cudaMalloc(&h->scratch, h->scratch_doubles * sizeof(double)));
    h->gate_slots = 4;
    h->gate_slot_bytes = size_t(16) << (2 * QJ_MAX_TARGETS);  // complex128 2^k x 2^k, k = QJ_MAX_TARGETS
    h->gate_done = new cudaEvent_t[h->gate_slots];
    for (int i = 0; i < h->gate_slots; i++)
        QJ_CUDA_OK(cudaEventCreateWithFlags(&h->gate_done[i], cudaEventDisableTiming));
    *out = h;
    return QJ_OK;
}

int qj_destroy(qj_handle *h) {
    qj::DeviceGuard device_guard(h);
    if (!h) return QJ_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (int i = 0; i < h->gate_slots; i++) cudaEventDestroy(h->gate_done[i]);
    delete[] h->gate_done;
    cudaFree(h->scratch);
    if (h->gate_dev) cudaFree(h->gate_dev);
    if (h->gate_pin) cudaFreeHost(h->gate_pin);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return QJ_OK;
}

int qj_set_stream(qj_handle *h, void *stream) {
    qj::DeviceGuard device_guard(h);
    QJ_REQUIRE(h != nullptr, "null handle");
    if (h->own_stream) {
        cudaStreamSynchronize(h->stream);
        cudaStreamDestroy(h->stream);
        h->own_stream = false;
    }
    h->stream = static_cast<cudaStream_t>(stream);
    return QJ_OK;
}

int qj_sync(qj_handle *h) {
    qj::DeviceGuard device_guard(h);
    QJ_REQUIRE(h != nullptr, "null handle");
    QJ_CUDA_OK(cudaStreamSynchronize(h->stream));
    return QJ_OK;
}

int64_t qj_launch_count(qj_handle *h) { return h ? h->launches : 0; }

int qj_set_route(qj_handle *h, int route) {
    qj::DeviceGuard device_guard(h);
    QJ_REQUIRE(h != nullptr, "null handle");
    QJ_REQUIRE(route >= 0 && route <= 2, "route must be 0, 1 or 2");
    h->route = route;
    return QJ_OK;
}

int qj_apply_gate(qj_handle *h, void *state, int dtype, int nqubits, int m, const void *gate,
                  const int32_t *qubits, int nactive) {
    qj::DeviceGuard device_guard(h);
    return one_target(h, state, dtype, nqubits, m, gate, qubits, nactive, 0);
}
int qj_apply_x(qj_handle *h, void *state, int dtype, int nqubits, int m, const int32_t *qubits, int nactive) {
    qj::DeviceGuard device_guard(h);
    return one_target(h, state, dtype, nqubits, m, nullptr, qubits, nactive, OP_X);
}
int qj_apply_y(qj_handle *h, void *state, int dtype, int nqubits, int m, const int32_t *qubits, int nactive) {
    qj::DeviceGuard device_guard(h);
    return one_target(h, state, dtype, nqubits, m, nullptr, qubits, nactive, OP_Y);
}
int qj_apply_z(qj_handle *h, void *state, int dtype, int nqubits, int m, const int32_t *qubits, int nactive) {
    qj::DeviceGuard device_guard(h);
    return one_target(h, state, dtype, nqubits, m, nullptr, qubits, nactive, OP_Z);
}
int qj_apply_z_pow(qj_handle *h, void *state, int dtype, int nqubits, int m, const void *phase,
                   const int32_t *qubits, int nactive) {
    qj::DeviceGuard device_guard(h);
    return one_target(h, state, dtype, nqubits, m, phase, qubits, nactive, OP_ZPOW);
}

int qj_apply_phase(qj_handle *h, void *state, int dtype, int nqubits, const void *phase) {
    qj::DeviceGuard device_guard(h);
    int rc = check_common(h, state, dtype, nqubits);
    if (rc) return rc;
    QJ_REQUIRE(phase != nullptr, "null phase");
    GateCall c;
    memset(&c, 0, sizeof(c));
    c.state = state; c.dtype = dtype; c.nqubits = nqubits; c.ntargets = 0; c.gate = phase;
    return launch_special(h, c, OP_PHASE);
}

int qj_apply_two_qubit_gate(qj_handle *h, void *state, int dtype, int nqubits, int m1, int m2,
                            int swap_targets, const void *gate, const int32_t *qubits, int nactive) {
    qj::DeviceGuard device_guard(h);
    return two_target(h, state, dtype, nqubits, m1, m2, swap_targets, gate, qubits, nactive, 0);
}
int qj_apply_swap(qj_handle *h, void *state, int dtype, int nqubits, int m1, int m2,
                  const int32_t *qubits, int nactive) {
    qj::DeviceGuard device_guard(h);
    return two_target(h, state, dtype, nqubits, m1, m2, 0, nullptr, qubits, nactive, OP_SWAP);
}
int qj_apply_fsim(qj_handle *h, void *state, int dtype, int nqubits, int m1, int m2, int swap_targets,
                  const void *gate, const int32_t *qubits, int nactive) {
    qj::DeviceGuard device_guard(h);
    return two_target(h, state, dtype, nqubits, m1, m2, swap_targets, gate, qubits, nactive, OP_FSIM);
}

int qj_apply_multi_qubit_gate(qj_handle *h, void *state, int dtype, int nqubits, const void *gate,
                              const int32_t *qubits, int nactive, const int64_t *targets, int ntargets) {
    qj::DeviceGuard device_guard(h);
    int rc = check_common(h, state, dtype, nqubits);
    if (rc) return rc;
    QJ_REQUIRE(gate != nullptr && targets != nullptr, "null gate or targets");
    if (ntargets < 1 || ntargets > QJ_MAX_TARGETS)
        return fail(QJ_ERR_INVALID, "Number of target qubits must be <= " + std::to_string(QJ_MAX_TARGETS) +
                                        " but is " + std::to_string(ntargets) + ".");
    QJ_REQUIRE(ntargets <= nqubits, "more targets than qubits");
    GateCall c;
    memset(&c, 0, sizeof(c));
    c.state = state; c.dtype = dtype; c.nqubits = nqubits; c.ntargets = ntargets; c.gate = gate;
    for (int u = 0; u < ntargets; u++) {
        const int64_t mask = targets[u];
        QJ_REQUIRE(mask > 0 && (mask & (mask - 1)) == 0, "targets must be single-bit masks");
        c.tbits[u] = __builtin_ctzll((unsigned long long)mask);
        QJ_REQUIRE(c.tbits[u] < nqubits, "target bit out of range");
    }
    rc = fill_controls(c, qubits, nactive);
    if (rc) return rc;
    return route_dense(h, c);
}

}  // extern "C"
