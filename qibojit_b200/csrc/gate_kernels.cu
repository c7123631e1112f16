// gate_kernels.cu -- register ("direct") gate kernels for sm_100a.
//
// One thread owns one 2^k-tuple of amplitudes (times V amplitudes side by side when the
// lowest index bit is not touched by the gate, so every access is a 16-byte vector).  The
// group index is expanded to an amplitude index by inserting zero bits at the sorted
// control/target positions -- the bit-twiddling form of the reference's
// `multicontrol_index` (gates.py:6-12) -- and control bits are OR-ed in as a mask instead of
// the reference's add-and-subtract addressing.  Lanes of a warp walk consecutive free bits, so
// loads and stores are coalesced whenever the gate leaves the low bits alone; the tile kernel
// (tile_kernels.cu) covers the low-bit and k >= 3 cases through shared memory.
//
// The gate matrix travels in the kernel parameters and is consumed as constant-bank
// operands: no device allocation, copy or synchronisation per gate (the reference
// synchronises the stream after every launch, gpu.py:1006,1035,1075).
//
// Semantics follow gates.py: dense 1/2/k-target (gates.py:16-38, 118-193, 266-424),
// X/Y/Z/Z^p (gates.py:42-114), SWAP (gates.py:197-216), fSim (gates.py:220-254).

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace qj {

namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------ dense k-target
template <typename T, int V, int K, int U>
__global__ void __launch_bounds__(kThreads)
k_dense_direct(typename VecOf<T, V>::type *__restrict__ st, const __grid_constant__ GateGeom geo,
               const __grid_constant__ CMat<T, (1 << K)> mat) {
    constexpr int NE = 1 << K;
    const int64_t ngroups = int64_t(1) << geo.nfree;
    const int64_t g0 = (int64_t(blockIdx.x) * U) * kThreads + threadIdx.x;

    Amp<T, V> x[U][NE];
    int64_t base[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int64_t g = g0 + int64_t(u) * kThreads;
        base[u] = -1;
        if (g < ngroups) {
            base[u] = expand_index(g, geo) | geo.cmask;
#pragma unroll
            for (int e = 0; e < NE; e++) x[u][e] = ld_amp(st + base[u] + geo.off[e]);
        }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        if (base[u] < 0) continue;
#pragma unroll
        for (int i = 0; i < NE; i++) {
            Amp<T, V> acc;
#pragma unroll
            for (int l = 0; l < V; l++) { acc.re[l] = T(0); acc.im[l] = T(0); }
#pragma unroll
            for (int j = 0; j < NE; j++) {
                const T gr = mat.v[2 * (i * NE + j)], gi = mat.v[2 * (i * NE + j) + 1];
#pragma unroll
                for (int l = 0; l < V; l++) {
                    acc.re[l] = fma(gr, x[u][j].re[l], acc.re[l]);
                    acc.re[l] = fma(-gi, x[u][j].im[l], acc.re[l]);
                    acc.im[l] = fma(gr, x[u][j].im[l], acc.im[l]);
                    acc.im[l] = fma(gi, x[u][j].re[l], acc.im[l]);
                }
            }
            st_amp(st + base[u] + geo.off[i], acc);
        }
    }
}

// ------------------------------------------------------------------ X / Y / SWAP
// two amplitudes at base+off[0], base+off[1] are exchanged (Y: with factors -i / +i)
template <typename T, int V, int U, int OP>
__global__ void __launch_bounds__(kThreads)
k_perm2(typename VecOf<T, V>::type *__restrict__ st, const __grid_constant__ GateGeom geo) {
    const int64_t ngroups = int64_t(1) << geo.nfree;
    const int64_t g0 = (int64_t(blockIdx.x) * U) * kThreads + threadIdx.x;
    Amp<T, V> a[U], b[U];
    int64_t base[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int64_t g = g0 + int64_t(u) * kThreads;
        base[u] = -1;
        if (g < ngroups) {
            base[u] = expand_index(g, geo) | geo.cmask;
            a[u] = ld_amp(st + base[u] + geo.off[0]);
            b[u] = ld_amp(st + base[u] + geo.off[1]);
        }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        if (base[u] < 0) continue;
        if (OP == OP_Y) {
            Amp<T, V> na, nb;
#pragma unroll
            for (int l = 0; l < V; l++) {
                na.re[l] = b[u].im[l];  na.im[l] = -b[u].re[l];   // -i * s2
                nb.re[l] = -a[u].im[l]; nb.im[l] = a[u].re[l];    // +i * s1
            }
            st_amp(st + base[u] + geo.off[0], na);
            st_amp(st + base[u] + geo.off[1], nb);
        } else {
            st_amp(st + base[u] + geo.off[0], b[u]);
            st_amp(st + base[u] + geo.off[1], a[u]);
        }
    }
}

// ------------------------------------------------------------------ Z / Z^p (diagonal)
// only the amplitude with every control and the target bit set is touched
template <typename T, int V, int U, bool NEGATE>
__global__ void __launch_bounds__(kThreads)
k_diag1(typename VecOf<T, V>::type *__restrict__ st, const __grid_constant__ GateGeom geo, T pr,
        T pi) {
    const int64_t ngroups = int64_t(1) << geo.nfree;
    const int64_t g0 = (int64_t(blockIdx.x) * U) * kThreads + threadIdx.x;
    Amp<T, V> a[U];
    int64_t base[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int64_t g = g0 + int64_t(u) * kThreads;
        base[u] = -1;
        if (g < ngroups) {
            base[u] = (expand_index(g, geo) | geo.cmask) + geo.off[0];
            a[u] = ld_amp(st + base[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        if (base[u] < 0) continue;
        Amp<T, V> r;
#pragma unroll
        for (int l = 0; l < V; l++) {
            if (NEGATE) {
                r.re[l] = -a[u].re[l]; r.im[l] = -a[u].im[l];
            } else {
                r.re[l] = fma(pr, a[u].re[l], -pi * a[u].im[l]);
                r.im[l] = fma(pr, a[u].im[l], pi * a[u].re[l]);
            }
        }
        st_amp(st + base[u], r);
    }
}

// ------------------------------------------------------------------ fSim
template <typename T, int V, int U>
__global__ void __launch_bounds__(kThreads)
k_fsim(typename VecOf<T, V>::type *__restrict__ st, const __grid_constant__ GateGeom geo,
       const __grid_constant__ CMat<T, 2> g4, T pr, T pi) {
    const int64_t ngroups = int64_t(1) << geo.nfree;
    const int64_t g0 = (int64_t(blockIdx.x) * U) * kThreads + threadIdx.x;
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int64_t g = g0 + int64_t(u) * kThreads;
        if (g >= ngroups) continue;
        const int64_t base = expand_index(g, geo) | geo.cmask;
        Amp<T, V> s1 = ld_amp(st + base + geo.off[0]);
        Amp<T, V> s2 = ld_amp(st + base + geo.off[1]);
        Amp<T, V> s3 = ld_amp(st + base + geo.off[2]);
        Amp<T, V> r1, r2, r3;
#pragma unroll
        for (int l = 0; l < V; l++) {
            r1.re[l] = g4.v[0] * s1.re[l] - g4.v[1] * s1.im[l] + g4.v[2] * s2.re[l] - g4.v[3] * s2.im[l];
            r1.im[l] = g4.v[0] * s1.im[l] + g4.v[1] * s1.re[l] + g4.v[2] * s2.im[l] + g4.v[3] * s2.re[l];
            r2.re[l] = g4.v[4] * s1.re[l] - g4.v[5] * s1.im[l] + g4.v[6] * s2.re[l] - g4.v[7] * s2.im[l];
            r2.im[l] = g4.v[4] * s1.im[l] + g4.v[5] * s1.re[l] + g4.v[6] * s2.im[l] + g4.v[7] * s2.re[l];
            r3.re[l] = pr * s3.re[l] - pi * s3.im[l];
            r3.im[l] = pr * s3.im[l] + pi * s3.re[l];
        }
        st_amp(st + base + geo.off[0], r1);
        st_amp(st + base + geo.off[1], r2);
        st_amp(st + base + geo.off[2], r3);
    }
}

// ------------------------------------------------------------------ generic k > 5
// A block stages GP tuples in shared memory; thread t produces output row(s) of one tuple.
// The matrix is read TRANSPOSED from device memory (coalesced across rows).
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_dense_generic(Cx<T> *__restrict__ st, const __grid_constant__ GateGeom geo,
                const Cx<T> *__restrict__ matT, int K, int64_t ngroups, int groups_per_block) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cx<T> *xs = reinterpret_cast<Cx<T> *>(smem_raw);
    __shared__ int64_t s_tmask[QJ_MAX_TARGETS];
    if (threadIdx.x < K) s_tmask[threadIdx.x] = geo.off[threadIdx.x];
    __syncthreads();
    const int NE = 1 << K;
    const int total = groups_per_block * NE;
    for (int64_t batch = blockIdx.x; batch * groups_per_block < ngroups; batch += gridDim.x) {
        for (int idx = threadIdx.x; idx < total; idx += kThreads) {
            const int grp = idx >> K, e = idx & (NE - 1);
            const int64_t g = batch * groups_per_block + grp;
            if (g < ngroups) {
                int64_t a = expand_index(g, geo) | geo.cmask;
                for (int u = 0; u < K; u++) if ((e >> u) & 1) a += s_tmask[u];
                xs[idx] = st[a];
            }
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < total; idx += kThreads) {
            const int grp = idx >> K, i = idx & (NE - 1);
            const int64_t g = batch * groups_per_block + grp;
            if (g < ngroups) {
                T ar = T(0), ai = T(0);
                const Cx<T> *x = xs + (grp << K);
                for (int j = 0; j < NE; j++) {
                    const Cx<T> m = matT[size_t(j) * NE + i];
                    const Cx<T> v = x[j];
                    ar = fma(m.re, v.re, ar); ar = fma(-m.im, v.im, ar);
                    ai = fma(m.re, v.im, ai); ai = fma(m.im, v.re, ai);
                }
                int64_t a = expand_index(g, geo) | geo.cmask;
                for (int u = 0; u < K; u++) if ((i >> u) & 1) a += s_tmask[u];
                Cx<T> r; r.re = ar; r.im = ai;
                st[a] = r;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ host side
struct Plan {
    GateGeom geo;
    int v;  // log2 of the amplitudes per access
};

// Build the index geometry.  `extra_fixed` lists target bits that the kernel treats like
// controls (diagonal ops: the touched amplitude has the target bit set).
int make_plan(const GateCall &c, int nelem_bits, const int *elem_bits, const int *fixed_one,
              int nfixed, bool allow_vec2, Plan *out) {
    Plan &p = *out;
    memset(&p, 0, sizeof(p));
    std::vector<int> active;
    for (int i = 0; i < nelem_bits; i++) active.push_back(elem_bits[i]);
    for (int i = 0; i < nfixed; i++) active.push_back(fixed_one[i]);
    std::sort(active.begin(), active.end());
    for (size_t i = 0; i < active.size(); i++) {
        if (active[i] < 0 || active[i] >= c.nqubits)
            return fail(QJ_ERR_INVALID, "qubit index out of range");
        if (i && active[i] == active[i - 1])
            return fail(QJ_ERR_INVALID, "duplicate qubit in gate");
    }
    if ((int)active.size() > kMaxPos) return fail(QJ_ERR_INVALID, "too many active qubits");
    const bool bit0_active = !active.empty() && active[0] == 0;
    p.v = (allow_vec2 && c.dtype == QJ_C64 && !bit0_active && c.nqubits >= 1 + (int)active.size()) ? 1 : 0;
    p.geo.npos = (int)active.size();
    for (int i = 0; i < p.geo.npos; i++) p.geo.pos[i] = active[i] - p.v;
    p.geo.nfree = c.nqubits - p.v - p.geo.npos;
    for (int i = 0; i < nfixed; i++) p.geo.cmask |= int64_t(1) << (fixed_one[i] - p.v);
    return QJ_OK;
}

template <typename F>
int launch_checked(qj_handle *h, F &&f) {
    f();
    h->launches++;
    QJ_CUDA_OK(cudaGetLastError());
    return QJ_OK;
}

inline unsigned grid_for(int nfree, int U) {
    const int64_t ngroups = int64_t(1) << nfree;
    const int64_t per_block = int64_t(kThreads) * U;
    return (unsigned)((ngroups + per_block - 1) / per_block);
}

template <typename T, int V, int K>
int launch_dense_k(qj_handle *h, const GateCall &c, const Plan &p) {
    constexpr int NE = 1 << K;
    constexpr int U = (K == 1) ? 4 : (K == 2 ? 2 : 1);
    CMat<T, NE> mat;
    memcpy(mat.v, c.gate, sizeof(mat.v));
    using Vec = typename VecOf<T, V>::type;
    return launch_checked(h, [&] {
        k_dense_direct<T, V, K, U><<<grid_for(p.geo.nfree, U), kThreads, 0, h->stream>>>(
            reinterpret_cast<Vec *>(c.state), p.geo, mat);
    });
}

template <typename T, int V>
int launch_dense_tv(qj_handle *h, const GateCall &c, const Plan &p) {
    switch (c.ntargets) {
        case 1: return launch_dense_k<T, V, 1>(h, c, p);
        case 2: return launch_dense_k<T, V, 2>(h, c, p);
        case 3: return launch_dense_k<T, V, 3>(h, c, p);
        case 4: return launch_dense_k<T, V, 4>(h, c, p);
        case 5: return launch_dense_k<T, V, 5>(h, c, p);
    }
    return fail(QJ_ERR_INVALID, "direct kernel supports 1..5 targets");
}

}  // namespace

int launch_dense_direct(qj_handle *h, const GateCall &c) {
    Plan p;
    // complex64, k = 5: the two-amplitude vector form needs ~170 registers and runs at half the
    // speed of the scalar form (measured, profiles/sweep_c64): keep one amplitude per access
    int rc = make_plan(c, c.ntargets, c.tbits, c.cbits, c.ncontrols, c.ntargets < 5, &p);
    if (rc) return rc;
    for (int e = 0; e < (1 << c.ntargets); e++) {
        int64_t o = 0;
        for (int u = 0; u < c.ntargets; u++)
            if ((e >> u) & 1) o += int64_t(1) << (c.tbits[u] - p.v);
        p.geo.off[e] = o;
    }
    if (c.dtype == QJ_C128) return launch_dense_tv<double, 1>(h, c, p);
    if (p.v == 1) return launch_dense_tv<float, 2>(h, c, p);
    return launch_dense_tv<float, 1>(h, c, p);
}

namespace {

template <typename T, int V>
int launch_special_tv(qj_handle *h, const GateCall &c, const Plan &p, int op) {
    using Vec = typename VecOf<T, V>::type;
    Vec *st = reinterpret_cast<Vec *>(c.state);
    const T *g = reinterpret_cast<const T *>(c.gate);
    constexpr int U = 4;
    const unsigned grid = grid_for(p.geo.nfree, op == OP_FSIM ? 2 : U);
    switch (op) {
        case OP_X:
        case OP_SWAP:
            return launch_checked(h, [&] {
                k_perm2<T, V, U, OP_X><<<grid, kThreads, 0, h->stream>>>(st, p.geo);
            });
        case OP_Y:
            return launch_checked(h, [&] {
                k_perm2<T, V, U, OP_Y><<<grid, kThreads, 0, h->stream>>>(st, p.geo);
            });
        case OP_Z:
            return launch_checked(h, [&] {
                k_diag1<T, V, U, true><<<grid, kThreads, 0, h->stream>>>(st, p.geo, T(0), T(0));
            });
        case OP_PHASE:
        case OP_ZPOW:
            return launch_checked(h, [&] {
                k_diag1<T, V, U, false><<<grid, kThreads, 0, h->stream>>>(st, p.geo, g[0], g[1]);
            });
        case OP_FSIM: {
            CMat<T, 2> m;
            memcpy(m.v, g, sizeof(m.v));
            return launch_checked(h, [&] {
                k_fsim<T, V, 2><<<grid, kThreads, 0, h->stream>>>(st, p.geo, m, g[8], g[9]);
            });
        }
    }
    return fail(QJ_ERR_INVALID, "unknown special op");
}

}  // namespace

int launch_special(qj_handle *h, const GateCall &c, int op) {
    Plan p;
    int rc;
    std::vector<int> fixed(c.cbits, c.cbits + c.ncontrols);
    if (op == OP_Z || op == OP_ZPOW || op == OP_PHASE) {
        // diagonal: target bit behaves like one more control; the single touched element
        // sits at offset 0 from the (controls | target) base.  OP_PHASE has no qubits at all.
        if (op != OP_PHASE) fixed.push_back(c.tbits[0]);
        rc = make_plan(c, 0, nullptr, fixed.data(), (int)fixed.size(), true, &p);
        if (rc) return rc;
        p.geo.off[0] = 0;
    } else {
        rc = make_plan(c, c.ntargets, c.tbits, fixed.data(), (int)fixed.size(), true, &p);
        if (rc) return rc;
        if (op == OP_X || op == OP_Y) {
            p.geo.off[0] = 0;
            p.geo.off[1] = int64_t(1) << (c.tbits[0] - p.v);
        } else {  // SWAP, FSIM: elements |01>, |10>, |11> of the pair
            p.geo.off[0] = int64_t(1) << (c.tbits[0] - p.v);
            p.geo.off[1] = int64_t(1) << (c.tbits[1] - p.v);
            p.geo.off[2] = p.geo.off[0] + p.geo.off[1];
        }
    }
    if (c.dtype == QJ_C128) return launch_special_tv<double, 1>(h, c, p, op);
    if (p.v == 1) return launch_special_tv<float, 2>(h, c, p, op);
    return launch_special_tv<float, 1>(h, c, p, op);
}

namespace {
template <typename T>
int launch_generic_t(qj_handle *h, const GateCall &c, const Plan &p) {
    const int K = c.ntargets, NE = 1 << K;
    // transpose on the host while staging so device reads are coalesced over rows
    std::vector<T> tr(size_t(2) * NE * NE);
    const T *src = reinterpret_cast<const T *>(c.gate);
    for (int i = 0; i < NE; i++)
        for (int j = 0; j < NE; j++) {
            tr[2 * (size_t(j) * NE + i)] = src[2 * (size_t(i) * NE + j)];
            tr[2 * (size_t(j) * NE + i) + 1] = src[2 * (size_t(i) * NE + j) + 1];
        }
    void *dmat = nullptr;
    int slot = -1;
    int rc = stage_gate_matrix(h, tr.data(), tr.size() * sizeof(T), &dmat, &slot);
    if (rc) return rc;
    const int64_t ngroups = int64_t(1) << p.geo.nfree;
    int gpb = std::max(1, kThreads / NE);
    if (gpb > ngroups) gpb = (int)ngroups;
    const size_t smem = size_t(gpb) * NE * sizeof(Cx<T>);
    int64_t nbatches = (ngroups + gpb - 1) / gpb;
    unsigned grid = (unsigned)std::min<int64_t>(nbatches, int64_t(h->sm_count) * 8);
    rc = launch_checked(h, [&] {
        k_dense_generic<T><<<grid, kThreads, smem, h->stream>>>(
            reinterpret_cast<Cx<T> *>(c.state), p.geo, reinterpret_cast<const Cx<T> *>(dmat), K,
            ngroups, gpb);
    });
    gate_slot_release(h, slot);
    return rc;
}
}  // namespace

int launch_dense_generic(qj_handle *h, const GateCall &c) {
    if (c.ntargets > QJ_MAX_TARGETS) return fail(QJ_ERR_INVALID, "too many target qubits");
    Plan p;
    int rc = make_plan(c, c.ntargets, c.tbits, c.cbits, c.ncontrols, false, &p);
    if (rc) return rc;
    for (int u = 0; u < c.ntargets; u++) p.geo.off[u] = int64_t(1) << c.tbits[u];  // masks
    if (c.dtype == QJ_C128) return launch_generic_t<double>(h, c, p);
    return launch_generic_t<float>(h, c, p);
}

}  // namespace qj
