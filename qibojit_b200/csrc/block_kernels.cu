// block_kernels.cu -- multi-gate passes ("tile programs") for sm_100a.
//
// The per-gate kernels (gate_kernels.cu) already run at the HBM roofline, so the time of a
// circuit is the number of passes over the state.  This kernel removes passes: a CTA stages a
// tile of 2^T amplitudes in shared memory -- the low r index bits (a contiguous run, so global
// traffic is made of whole 2^r-amplitude segments) plus T-r arbitrary higher index bits -- and
// applies a whole PROGRAM of gates to it before writing it back:
//
//   * dense 1- and 2-target gates whose targets are tile-local bits; controls may be local or
//     "outer" (a tile-constant predicate on the tile's base index);
//   * diagonal phase tables over up to 12 arbitrary index bits (local bits become bit fields of
//     the local index, outer bits a tile-constant table offset): diagonal gates need no
//     locality at all, which is what lets a QFT level run inside one pass.
//
// Inside the tile the program is executed in ROUNDS: each thread gathers one group of N = 8
// (complex128) or 16 (complex64) amplitudes -- J = 3 / 4 "register" bits chosen by the host --
// from shared memory, applies every micro-op of the round to the registers, and scatters them
// back, so shared memory is read and written once per round, not once per gate.  Runs of gates
// that live entirely on the register bits are multiplied together on the host and applied as
// ONE N x N mat-vec (straight-line FMA code, no per-gate dispatch).  Shared memory uses the
// 16-byte-chunk XOR swizzle (the layout TMA calls SWIZZLE_128B), which keeps the gathers
// conflict-free for any register bits.  One persistent 512-thread CTA per SM double-buffers its
// tiles: the asynchronous copies (LDGSTS) of tile i+1 are in flight while tile i is computed.
//
// Arithmetic contract per micro-op: gates.py:16-38 (one target), gates.py:118-193 (two
// targets), gates.py:82-114 (diagonals, here pre-multiplied into tables on the host).

#include <algorithm>
#include <complex>
#include <vector>

#include "common.cuh"

struct qj_program {
    int dtype = 0;
    int nqubits = 0;
    struct Launch {
        // geometry
        int T, r, nh;
        int hibit[16];
        int npos;
        int pos[QJ_MAX_QUBITS];
        int64_t ntiles;
        // device ranges
        int first_round, nrounds;
        int first_mop, nmops;
    };
    std::vector<Launch> launches;
    void *d_rounds = nullptr;
    void *d_mops = nullptr;
    void *d_outers = nullptr;
    void *d_data = nullptr;
    int64_t total_mops = 0, total_rounds = 0;
};

namespace qj {
namespace {

constexpr int kBlkThreads = 512;      // one CTA per SM: 512 groups of a 64 KiB tile, one per thread
constexpr int kRegBits = 4;           // register bits per round (upper bound; complex128 uses 3)
template <typename T>
struct RegBits {                      // 32 data registers per thread either way
    static constexpr int J = (sizeof(T) == 8) ? 3 : 4;
    static constexpr int N = 1 << J;
};
inline int reg_bits(int dtype) { return dtype == QJ_C128 ? 3 : 4; }
constexpr int kMaxHiBits = 8;
constexpr int kMaxFields = 6;
constexpr int kMaxMopsPerLaunch = 512;

inline int max_tile_bits(int dtype) { return dtype == QJ_C128 ? 12 : 13; }  // 64 KiB tiles

enum {
    MOP_DENSE1 = 1,       // 2x2 on one register slot, matrix inline
    MOP_DENSE1_REAL = 2,  // ... with a real matrix (H, RY): half the arithmetic
    MOP_DENSE2 = 3,       // 4x4 on two register slots
    MOP_PERM1 = 4,        // X: exchange the pair (CNOT / Toffoli through the masks)
    MOP_PERM2 = 5,        // SWAP: exchange |01> and |10>
    MOP_DIAG = 6,         // phase table
    MOP_DENSEJ = 7,       // N x N on all register slots: a run of gates multiplied on the host
    MOP_DENSEJ_REAL = 8,
};

struct Round {
    int32_t rbit[kRegBits];  // local positions of the register bits, ascending
    int32_t first;           // first micro-op (relative to the launch)
    int32_t count;
    int32_t pad[2];
};

// Everything that is uniform over the threads is resolved on the host when a round is closed:
// control bits on register slots become a bitmask over the group elements (`emask`; dense ops
// with such controls are folded into an N x N matrix instead), control bits elsewhere a mask on
// the thread's base index (`tmask`), and a diagonal table's index splits into base fields (per
// thread, once per op) | per-element constants (`eidx`).
struct __align__(16) MicroOp {
    int32_t kind;
    int32_t ab;            // register slots: a | b << 4; a < b for two-target ops
    uint32_t tmask;        // non-register local bits that must be 1 (thread-level predicate)
    int32_t data_off;      // complex elements (dense2 / N x N matrix, diag table)
    uint32_t emask;        // perm / diag: bit e = group element e satisfies the register-slot controls
    int32_t nfields;       // diag: bit fields of the thread's base index
    uint32_t field[kMaxFields];  // src | len << 8 | dst << 16
    union {
        uint16_t eidx[16];       // diag: table-index contribution of group element e
        unsigned char mat[64];   // one-target gates: the 2x2 matrix, inline (arrives with the op prefetch)
    } u;
};
static_assert(sizeof(MicroOp) == 112, "MicroOp is read as 16-byte vectors");

struct OuterDesc {
    uint64_t ocmask;       // bits of the tile base that must be 1
    int32_t nbits;         // diag: outer table bits
    uint8_t src[QJ_MAX_DIAG_BITS], dst[QJ_MAX_DIAG_BITS];
    int32_t pad;
};

struct PassGeom {
    int T, r, nh;          // r counted in amplitudes
    int hibit[kMaxHiBits];
    int npos;
    int pos[QJ_MAX_QUBITS];
    int64_t ntiles;
    int nrounds, nmops;
};

template <typename T>
__device__ __forceinline__ int swz_amp(int l) {
    if (sizeof(T) == 8) return l ^ ((l >> 3) & 7);        // complex128: one amplitude per 16-byte chunk
    return l ^ (((l >> 4) & 7) << 1);                       // complex64: two amplitudes per chunk
}

template <typename T>
__device__ __forceinline__ void cmul_acc(T &ar, T &ai, T gr, T gi, T xr, T xi) {
    ar = fma(gr, xr, ar);
    ar = fma(-gi, xi, ar);
    ai = fma(gr, xi, ai);
    ai = fma(gi, xr, ai);
}

// local offset of group element e given the register-bit strides
__device__ __forceinline__ int elem_off(int e, const int (&o)[kRegBits]) {
    return ((e & 1) ? o[0] : 0) | ((e & 2) ? o[1] : 0) | ((e & 4) ? o[2] : 0) | ((e & 8) ? o[3] : 0);
}

// ---- micro-ops on the N register amplitudes of a group ---------------------------------------
template <typename T, int A, bool REAL>
__device__ __forceinline__ void mop_dense1(Cx<T> (&x)[RegBits<T>::N], const Cx<T> (&m)[4]) {
    constexpr int N = RegBits<T>::N;
    if constexpr (A < RegBits<T>::J) {
#pragma unroll
        for (int p = 0; p < N / 2; p++) {
            const int e0 = ((p >> A) << (A + 1)) | (p & ((1 << A) - 1));
            const int e1 = e0 | (1 << A);
            const Cx<T> s0 = x[e0 % N], s1 = x[e1 % N];
            Cx<T> y0, y1;
            if (REAL) {
                y0.re = fma(m[0].re, s0.re, m[1].re * s1.re);
                y0.im = fma(m[0].re, s0.im, m[1].re * s1.im);
                y1.re = fma(m[2].re, s0.re, m[3].re * s1.re);
                y1.im = fma(m[2].re, s0.im, m[3].re * s1.im);
            } else {
                y0.re = T(0); y0.im = T(0); y1.re = T(0); y1.im = T(0);
                cmul_acc(y0.re, y0.im, m[0].re, m[0].im, s0.re, s0.im);
                cmul_acc(y0.re, y0.im, m[1].re, m[1].im, s1.re, s1.im);
                cmul_acc(y1.re, y1.im, m[2].re, m[2].im, s0.re, s0.im);
                cmul_acc(y1.re, y1.im, m[3].re, m[3].im, s1.re, s1.im);
            }
            x[e0 % N] = y0; x[e1 % N] = y1;
        }
    }
}

// expand a quad index p to the element with zeros at slots A < B
template <int A, int B>
__device__ __forceinline__ constexpr int quad_base(int p) {
    int e = ((p >> A) << (A + 1)) | (p & ((1 << A) - 1));
    e = ((e >> B) << (B + 1)) | (e & ((1 << B) - 1));
    return e;
}

// two-target gate: matrix index bit 0 <-> slot A, bit 1 <-> slot B (A < B)
template <typename T, int A, int B>
__device__ __forceinline__ void mop_dense2(Cx<T> (&x)[RegBits<T>::N], const Cx<T> *__restrict__ m) {
    constexpr int N = RegBits<T>::N;
    if constexpr (B < RegBits<T>::J) {
        Cx<T> g[16];
#pragma unroll
        for (int i = 0; i < 16; i++) g[i] = m[i];
#pragma unroll
        for (int p = 0; p < N / 4; p++) {
            const int e0 = quad_base<A, B>(p);
            Cx<T> y[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                T ar = T(0), ai = T(0);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const Cx<T> sj = x[(e0 | ((j & 1) << A) | ((j >> 1) << B)) % N];
                    cmul_acc(ar, ai, g[i * 4 + j].re, g[i * 4 + j].im, sj.re, sj.im);
                }
                y[i].re = ar; y[i].im = ai;
            }
#pragma unroll
            for (int i = 0; i < 4; i++) x[(e0 | ((i & 1) << A) | ((i >> 1) << B)) % N] = y[i];
        }
    }
}

template <typename T, int A>
__device__ __forceinline__ void mop_perm1(Cx<T> (&x)[RegBits<T>::N], uint32_t emask) {
    constexpr int N = RegBits<T>::N;
    if constexpr (A < RegBits<T>::J) {
#pragma unroll
        for (int p = 0; p < N / 2; p++) {
            const int e0 = ((p >> A) << (A + 1)) | (p & ((1 << A) - 1));
            const int e1 = e0 | (1 << A);
            if (!((emask >> e0) & 1u)) continue;
            const Cx<T> t = x[e0 % N]; x[e0 % N] = x[e1 % N]; x[e1 % N] = t;
        }
    }
}

template <typename T, int A, int B>
__device__ __forceinline__ void mop_perm2(Cx<T> (&x)[RegBits<T>::N], uint32_t emask) {
    constexpr int N = RegBits<T>::N;
    if constexpr (B < RegBits<T>::J) {
#pragma unroll
        for (int p = 0; p < N / 4; p++) {
            const int e0 = quad_base<A, B>(p);
            if (!((emask >> e0) & 1u)) continue;
            const int ea = (e0 | (1 << A)) % N, eb = (e0 | (1 << B)) % N;
            const Cx<T> t = x[ea]; x[ea] = x[eb]; x[eb] = t;
        }
    }
}

// N x N mat-vec on the whole group: the matrix is uniform over the threads (broadcast loads)
template <typename T, bool REAL>
__device__ __forceinline__ void mop_dense_full(Cx<T> (&x)[RegBits<T>::N], const Cx<T> *__restrict__ m) {
    constexpr int N = RegBits<T>::N;
    Cx<T> y[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        T ar = T(0), ai = T(0);
        if (REAL) {  // row i = N reals packed in N/2 complex slots
#pragma unroll
            for (int j = 0; j < N; j += 2) {
                const Cx<T> c = m[(i * N + j) / 2];
                ar = fma(c.re, x[j].re, ar); ai = fma(c.re, x[j].im, ai);
                ar = fma(c.im, x[j + 1].re, ar); ai = fma(c.im, x[j + 1].im, ai);
            }
        } else {
#pragma unroll
            for (int j = 0; j < N; j++) {
                const Cx<T> c = m[i * N + j];
                cmul_acc(ar, ai, c.re, c.im, x[j].re, x[j].im);
            }
        }
        y[i].re = ar; y[i].im = ai;
    }
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = y[i];
}

// a micro-op as it sits in registers: fetched one op ahead of its use
struct OpRegs {
    int4 q0, q1, q3, q4;   // quads 2 (diag fields 2..5) and 5, 6 (second half of a complex128 2x2) on demand
    int32_t oi;
};

template <bool REAL>
__device__ __forceinline__ void unpack_mat(const OpRegs &r, const int4 *src, Cx<double> (&m)[4]) {
    if (REAL) {  // four doubles, compact
        m[0].re = __hiloint2double(r.q3.y, r.q3.x); m[1].re = __hiloint2double(r.q3.w, r.q3.z);
        m[2].re = __hiloint2double(r.q4.y, r.q4.x); m[3].re = __hiloint2double(r.q4.w, r.q4.z);
        m[0].im = m[1].im = m[2].im = m[3].im = 0.0;
    } else {
        const int4 q5 = __ldg(src + 5), q6 = __ldg(src + 6);
        m[0].re = __hiloint2double(r.q3.y, r.q3.x); m[0].im = __hiloint2double(r.q3.w, r.q3.z);
        m[1].re = __hiloint2double(r.q4.y, r.q4.x); m[1].im = __hiloint2double(r.q4.w, r.q4.z);
        m[2].re = __hiloint2double(q5.y, q5.x); m[2].im = __hiloint2double(q5.w, q5.z);
        m[3].re = __hiloint2double(q6.y, q6.x); m[3].im = __hiloint2double(q6.w, q6.z);
    }
}
template <bool REAL>
__device__ __forceinline__ void unpack_mat(const OpRegs &r, const int4 *, Cx<float> (&m)[4]) {
    m[0].re = __int_as_float(r.q3.x); m[0].im = __int_as_float(r.q3.y);
    m[1].re = __int_as_float(r.q3.z); m[1].im = __int_as_float(r.q3.w);
    m[2].re = __int_as_float(r.q4.x); m[2].im = __int_as_float(r.q4.y);
    m[3].re = __int_as_float(r.q4.z); m[3].im = __int_as_float(r.q4.w);
}

template <typename T>
__device__ __forceinline__ void exec_op(const OpRegs &r, const int4 *src, Cx<T> (&x)[RegBits<T>::N], int base,
                                         const Cx<T> *__restrict__ data) {
    constexpr int N = RegBits<T>::N;
    if (r.oi < 0) return;                                  // outer control not satisfied by this tile
    const int kind = r.q0.x, ab = r.q0.y;
    const uint32_t tmask = uint32_t(r.q0.z), emask = uint32_t(r.q1.x);
    if ((uint32_t(base) & tmask) != tmask) return;          // local control outside the register slots
    const Cx<T> *d = data + r.q0.w;
    // the register slots are run-time values of the program; the micro-ops need them as
    // compile-time register indices
    switch (kind) {
        case MOP_DENSEJ: mop_dense_full<T, false>(x, d); break;
        case MOP_DENSEJ_REAL: mop_dense_full<T, true>(x, d); break;
        case MOP_DENSE1: {
            Cx<T> m[4];
            unpack_mat<false>(r, src, m);
            switch (ab) {
                case 0: mop_dense1<T, 0, false>(x, m); break;
                case 1: mop_dense1<T, 1, false>(x, m); break;
                case 2: mop_dense1<T, 2, false>(x, m); break;
                default: mop_dense1<T, 3, false>(x, m); break;
            }
        } break;
        case MOP_DENSE1_REAL: {
            Cx<T> m[4];
            unpack_mat<true>(r, src, m);
            switch (ab) {
                case 0: mop_dense1<T, 0, true>(x, m); break;
                case 1: mop_dense1<T, 1, true>(x, m); break;
                case 2: mop_dense1<T, 2, true>(x, m); break;
                default: mop_dense1<T, 3, true>(x, m); break;
            }
        } break;
        case MOP_DENSE2:
            switch (ab) {  // a | b << 4, a < b
                case 0x10: mop_dense2<T, 0, 1>(x, d); break;
                case 0x20: mop_dense2<T, 0, 2>(x, d); break;
                case 0x21: mop_dense2<T, 1, 2>(x, d); break;
                case 0x30: mop_dense2<T, 0, 3>(x, d); break;
                case 0x31: mop_dense2<T, 1, 3>(x, d); break;
                default: mop_dense2<T, 2, 3>(x, d); break;
            }
            break;
        case MOP_PERM1:
            switch (ab) {
                case 0: mop_perm1<T, 0>(x, emask); break;
                case 1: mop_perm1<T, 1>(x, emask); break;
                case 2: mop_perm1<T, 2>(x, emask); break;
                default: mop_perm1<T, 3>(x, emask); break;
            }
            break;
        case MOP_PERM2:
            switch (ab) {
                case 0x10: mop_perm2<T, 0, 1>(x, emask); break;
                case 0x20: mop_perm2<T, 0, 2>(x, emask); break;
                case 0x21: mop_perm2<T, 1, 2>(x, emask); break;
                case 0x30: mop_perm2<T, 0, 3>(x, emask); break;
                case 0x31: mop_perm2<T, 1, 3>(x, emask); break;
                default: mop_perm2<T, 2, 3>(x, emask); break;
            }
            break;
        default: {  // MOP_DIAG: table index = outer part | fields of the base | element part
            const int nf = r.q1.y;
            int idxb = r.oi;
            auto field = [&](uint32_t fl) { return ((base >> (fl & 255)) & ((1 << ((fl >> 8) & 255)) - 1)) << (fl >> 16); };
            if (nf > 0) idxb |= field(uint32_t(r.q1.z));
            if (nf > 1) idxb |= field(uint32_t(r.q1.w));
            if (nf > 2) {  // rare: tables whose bits are scattered over the tile
                const int4 q2 = __ldg(src + 2);
                idxb |= field(uint32_t(q2.x));
                if (nf > 3) idxb |= field(uint32_t(q2.y));
                if (nf > 4) idxb |= field(uint32_t(q2.z));
                if (nf > 5) idxb |= field(uint32_t(q2.w));
            }
            const uint32_t ew[8] = {uint32_t(r.q3.x), uint32_t(r.q3.y), uint32_t(r.q3.z), uint32_t(r.q3.w),
                                    uint32_t(r.q4.x), uint32_t(r.q4.y), uint32_t(r.q4.z), uint32_t(r.q4.w)};
#pragma unroll
            for (int e4 = 0; e4 < N; e4 += 4) {  // four gathers in flight before the first multiply
                Cx<T> ph[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int e = e4 + k;
                    const int idx = idxb | int((e & 1) ? (ew[e >> 1] >> 16) : (ew[e >> 1] & 0xffffu));
                    if ((emask >> e) & 1u) ph[k] = d[idx];
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int e = e4 + k;
                    if (!((emask >> e) & 1u)) continue;
                    const Cx<T> v = x[e];
                    Cx<T> y;
                    y.re = fma(ph[k].re, v.re, -ph[k].im * v.im);
                    y.im = fma(ph[k].re, v.im, ph[k].im * v.re);
                    x[e] = y;
                }
            }
        }
    }
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                     static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
                 "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <typename T>
__global__ void __launch_bounds__(kBlkThreads, 1)
k_tile_program(Cx<T> *__restrict__ state, const __grid_constant__ PassGeom pg,
               const Round *__restrict__ rounds, const MicroOp *__restrict__ mops,
               const OuterDesc *__restrict__ outers, const Cx<T> *__restrict__ data) {
    constexpr int J = RegBits<T>::J;
    constexpr int N = RegBits<T>::N;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int64_t s_runoff[1 << kMaxHiBits];   // in 16-byte vectors
    __shared__ int32_t s_outer[kMaxMopsPerLaunch];  // -1: op inactive for this tile, else outer table index

    constexpr int VS = (sizeof(T) == 8) ? 0 : 1;    // log2 amplitudes per 16-byte vector
    const int tid = threadIdx.x;
    const int nvec = 1 << (pg.T - VS);
    const int rv = pg.r - VS;                        // run bits in vectors
    const int rvmask = (1 << rv) - 1;
    float4 *gvec = reinterpret_cast<float4 *>(state);
    float4 *const sbuf0 = reinterpret_cast<float4 *>(smem_raw);

    for (int run = tid; run < (1 << pg.nh); run += kBlkThreads) {
        int64_t off = 0;
        for (int b = 0; b < pg.nh; b++) off |= int64_t((run >> b) & 1) << (pg.hibit[b] - VS);
        s_runoff[run] = off;
    }
    __syncthreads();

    auto tile_base_vec = [&](int64_t tile_id) {  // insert zeros at the high local bits
        int64_t tb = tile_id;
#pragma unroll 1
        for (int j = 0; j < pg.npos; j++) {
            const int p = pg.pos[j];
            tb = ((tb >> p) << (p + 1)) | (tb & ((int64_t(1) << p) - 1));
        }
        return (tb << pg.r) >> VS;
    };
    auto prefetch = [&](int64_t tile_id, float4 *dst) {
        const int64_t bv = tile_base_vec(tile_id);
        for (int lv = tid; lv < nvec; lv += kBlkThreads)
            cp_async16(dst + (lv ^ ((lv >> 3) & 7)), gvec + bv + s_runoff[lv >> rv] + (lv & rvmask));
    };

    const int ngroups = 1 << (pg.T - J);
    const bool live = tid < ngroups;
    int64_t tile_id = blockIdx.x;
    if (tile_id < pg.ntiles) prefetch(tile_id, sbuf0);

    for (int it = 0; tile_id < pg.ntiles; tile_id += gridDim.x, it++) {
        float4 *svec = sbuf0 + ((it & 1) ? nvec : 0);
        Cx<T> *tile = reinterpret_cast<Cx<T> *>(svec);
        const int64_t base_vec = tile_base_vec(tile_id);
        const int64_t base_amp = base_vec << VS;
        // per-op tile-constant data: outer control predicate and outer table index (the rounds of
        // the previous tile, the only readers of s_outer, ended with a barrier)
        for (int m = tid; m < pg.nmops; m += kBlkThreads) {
            const OuterDesc *od = outers + m;
            const uint64_t ocmask = od->ocmask;
            int32_t v = 0;
            if ((uint64_t(base_amp) & ocmask) != ocmask) {
                v = -1;
            } else {
                const int nb = od->nbits;
                for (int b = 0; b < nb; b++) v |= int32_t((base_amp >> od->src[b]) & 1) << od->dst[b];
            }
            s_outer[m] = v;
        }
        cp_async_wait_all();
        __syncthreads();  // tile `it` has landed; everybody finished storing tile `it - 1`
        if (tile_id + gridDim.x < pg.ntiles) prefetch(tile_id + gridDim.x, sbuf0 + ((it & 1) ? 0 : nvec));

        // ---- rounds
#pragma unroll 1
        for (int rd = 0; rd < pg.nrounds; rd++) {
            const Round R = rounds[rd];
            if (live) {
                int o[kRegBits];
#pragma unroll
                for (int j = 0; j < kRegBits; j++) o[j] = (j < J) ? (1 << R.rbit[j]) : 0;
                int base = tid;
#pragma unroll
                for (int j = 0; j < J; j++) {
                    const int p = R.rbit[j];
                    base = ((base >> p) << (p + 1)) | (base & ((1 << p) - 1));
                }
                Cx<T> x[N];
#pragma unroll
                for (int e = 0; e < N; e++) x[e] = tile[swz_amp<T>(base | elem_off(e, o))];
                // ops are fetched one ahead so that the header / matrix latency of op m+1 hides
                // behind the arithmetic of op m
                auto fetch = [&](int m, OpRegs &r) {
                    const int4 *p = reinterpret_cast<const int4 *>(mops + m);
                    r.q0 = __ldg(p); r.q1 = __ldg(p + 1); r.q3 = __ldg(p + 3); r.q4 = __ldg(p + 4);
                    r.oi = s_outer[m];
                };
                const int mend = R.first + R.count;
                OpRegs nxt;
                fetch(R.first, nxt);
#pragma unroll 1
                for (int m = R.first; m < mend; m++) {
                    const OpRegs cur = nxt;
                    if (m + 1 < mend) fetch(m + 1, nxt);
                    exec_op<T>(cur, reinterpret_cast<const int4 *>(mops + m), x, base, data);
                }
#pragma unroll
                for (int e = 0; e < N; e++) tile[swz_amp<T>(base | elem_off(e, o))] = x[e];
            }
            __syncthreads();
        }

        // ---- store the tile (stores are fire-and-forget; the next trip's barrier orders the buffer reuse)
        constexpr int UNR = 8;
        for (int v0 = 0; v0 < nvec; v0 += kBlkThreads * UNR) {
#pragma unroll
            for (int u = 0; u < UNR; u++) {
                const int lv = v0 + u * kBlkThreads + tid;
                if (lv < nvec) gvec[base_vec + s_runoff[lv >> rv] + (lv & rvmask)] = svec[lv ^ ((lv >> 3) & 7)];
            }
        }
    }
    cp_async_wait_all();
}

// ------------------------------------------------------------------ host-side lowering
typedef std::complex<double> cd;

struct Pending {
    int kind;              // MOP_DENSE1 / MOP_DENSE2 / MOP_DIAG as described by the caller
    int nt;
    int tpos[2];           // local positions of dense targets (matrix-index bit j <-> tpos[j])
    uint32_t lcmask;       // local control bits
    uint64_t ocmask;       // outer control bits
    int64_t data_off;
    int ndiag;             // diag: table bits
    int dpos[QJ_MAX_DIAG_BITS];   // local position of table bit j, -1 when outer
    int nfields;           // diag: bit fields of the local index
    uint8_t fsrc[kMaxFields], flen[kMaxFields], fdst[kMaxFields];
    OuterDesc od;
};

template <typename T>
cd load_cx(const unsigned char *p, int64_t i) {
    const T *t = reinterpret_cast<const T *>(p);
    return cd(double(t[2 * i]), double(t[2 * i + 1]));
}
cd load_any(const std::vector<unsigned char> &hdata, int dtype, int64_t off) {
    return dtype == QJ_C128 ? load_cx<double>(hdata.data(), off) : load_cx<float>(hdata.data(), off);
}

}  // namespace
}  // namespace qj

using namespace qj;

extern "C" int qj_program_create(qj_handle *h, int dtype, int nqubits, const qj_pass_desc *passes,
                                 int npasses, const qj_op_desc *ops, int64_t nops, const void *data,
                                 int64_t ndata, qj_program **out) {
    QJ_REQUIRE(h && out && (passes || npasses == 0), "null argument");
    QJ_REQUIRE(dtype == QJ_C64 || dtype == QJ_C128, "dtype must be QJ_C64 or QJ_C128");
    QJ_REQUIRE(nqubits >= kRegBits && nqubits <= QJ_MAX_QUBITS, "tile programs need 4 <= nqubits <= QJ_MAX_QUBITS");
    QJ_REQUIRE(npasses >= 0 && nops >= 0 && ndata >= 0, "negative count");
    const size_t esz = (dtype == QJ_C128) ? 16 : 8;
    std::vector<unsigned char> hdata(size_t(ndata) * esz);
    if (ndata) memcpy(hdata.data(), data, hdata.size());

    const int J = reg_bits(dtype);
    const int N = 1 << J;
    std::vector<Round> rounds;
    std::vector<MicroOp> mops;
    std::vector<OuterDesc> outers;
    auto *prog = new qj_program();
    prog->dtype = dtype;
    prog->nqubits = nqubits;
    auto bail = [&](const std::string &msg) {
        delete prog;
        return fail(QJ_ERR_INVALID, msg);
    };
    auto store_cx = [&](int64_t off, cd v) {
        if (dtype == QJ_C128) {
            double *t = reinterpret_cast<double *>(hdata.data());
            t[2 * off] = v.real(); t[2 * off + 1] = v.imag();
        } else {
            float *t = reinterpret_cast<float *>(hdata.data());
            t[2 * off] = float(v.real()); t[2 * off + 1] = float(v.imag());
        }
    };

    for (int pi = 0; pi < npasses; pi++) {
        const qj_pass_desc &pd = passes[pi];
        const int T = pd.nlocal;
        if (T < kRegBits || T > max_tile_bits(dtype) || T > nqubits)
            return bail("pass: need 4 <= nlocal <= min(12 (complex128) / 13 (complex64), nqubits)");
        int lpos[QJ_MAX_QUBITS];
        for (int b = 0; b < QJ_MAX_QUBITS; b++) lpos[b] = -1;
        for (int i = 0; i < T; i++) {
            const int b = pd.local_bits[i];
            if (b < 0 || b >= nqubits) return bail("pass: local bit out of range");
            if (i && b <= pd.local_bits[i - 1]) return bail("pass: local bits must be strictly ascending");
            lpos[b] = i;
        }
        int r = 0;
        while (r < T && pd.local_bits[r] == r) r++;
        if (dtype == QJ_C64 && r < 1) return bail("pass: complex64 tiles must contain index bit 0");
        if (T - r > kMaxHiBits) return bail("pass: too many local bits above the contiguous run");
        if (pd.first_op < 0 || pd.nops < 0 || pd.first_op + pd.nops > nops) return bail("pass: op range out of bounds");

        qj_program::Launch geo;
        memset(&geo, 0, sizeof(geo));
        geo.T = T; geo.r = r; geo.nh = T - r;
        for (int i = r; i < T; i++) {
            geo.hibit[i - r] = pd.local_bits[i];
            geo.pos[i - r] = pd.local_bits[i] - r;
        }
        geo.npos = T - r;
        geo.ntiles = int64_t(1) << (nqubits - T);

        // ---- micro-ops and rounds
        std::vector<Pending> pend;
        std::vector<int> regs;  // local positions claimed by the open round
        int launch_first_round = (int)rounds.size(), launch_first_mop = (int)mops.size();
        auto close_launch = [&]() {
            qj_program::Launch L = geo;
            L.first_round = launch_first_round;
            L.nrounds = (int)rounds.size() - launch_first_round;
            L.first_mop = launch_first_mop;
            L.nmops = (int)mops.size() - launch_first_mop;
            if (L.nrounds > 0) prog->launches.push_back(L);
            launch_first_round = (int)rounds.size();
            launch_first_mop = (int)mops.size();
        };
        auto close_round = [&]() {
            if (pend.empty()) return;
            if ((int)mops.size() - launch_first_mop + (int)pend.size() > kMaxMopsPerLaunch) close_launch();
            std::sort(regs.begin(), regs.end());
            // pad with the highest free local positions so lanes walk the low bits
            for (int p = T - 1; p >= 0 && (int)regs.size() < J; p--)
                if (std::find(regs.begin(), regs.end(), p) == regs.end()) regs.push_back(p);
            std::sort(regs.begin(), regs.end());
            Round R;
            memset(&R, 0, sizeof(R));
            for (int j = 0; j < J; j++) R.rbit[j] = regs[j];
            R.first = (int)mops.size() - launch_first_mop;
            uint32_t regmask = 0;
            for (int j = 0; j < J; j++) regmask |= 1u << regs[j];
            auto slot_of = [&](int p) { return int(std::find(regs.begin(), regs.end(), p) - regs.begin()); };
            auto off_of = [&](int e) {
                uint32_t o = 0;
                for (int j = 0; j < J; j++) if ((e >> j) & 1) o |= 1u << regs[j];
                return o;
            };
            auto elem_ok = [&](const Pending &pe, int e) {  // register-slot controls satisfied by element e
                const uint32_t rc = pe.lcmask & regmask;
                return (off_of(e) & rc) == rc;
            };
            // does the op live entirely on the register slots (no thread- or tile-level predicate)?
            auto internal = [&](const Pending &pe) {
                if (pe.ocmask || (pe.lcmask & ~regmask)) return false;
                if (pe.kind != MOP_DIAG) return true;
                for (int j = 0; j < pe.ndiag; j++)
                    if (pe.dpos[j] < 0 || !((regmask >> pe.dpos[j]) & 1)) return false;
                return true;
            };
            // host application of an op to one N-vector over the register slots (column of a run's matrix)
            auto apply_host = [&](const Pending &pe, std::vector<cd> &v) {
                if (pe.kind == MOP_DIAG) {
                    for (int e = 0; e < N; e++) {
                        if (!elem_ok(pe, e)) continue;
                        int idx = 0;
                        for (int j = 0; j < pe.ndiag; j++) idx |= ((e >> slot_of(pe.dpos[j])) & 1) << j;
                        v[e] *= load_any(hdata, dtype, pe.data_off + idx);
                    }
                    return;
                }
                const int k = pe.nt, dim = 1 << k;
                int sl[2] = {slot_of(pe.tpos[0]), k == 2 ? slot_of(pe.tpos[1]) : 0};
                int tm = (1 << sl[0]) | (k == 2 ? (1 << sl[1]) : 0);
                for (int e = 0; e < N; e++) {
                    if ((e & tm) || !elem_ok(pe, e)) continue;
                    cd in[4], outv[4];
                    int idx[4];
                    for (int i = 0; i < dim; i++) {
                        idx[i] = e | ((i & 1) << sl[0]) | (k == 2 ? ((i >> 1) << sl[1]) : 0);
                        in[i] = v[idx[i]];
                    }
                    for (int i = 0; i < dim; i++) {
                        cd acc = 0;
                        for (int j = 0; j < dim; j++) acc += load_any(hdata, dtype, pe.data_off + i * dim + j) * in[j];
                        outv[i] = acc;
                    }
                    for (int i = 0; i < dim; i++) v[idx[i]] = outv[i];
                }
            };
            auto base_mop = [&](const Pending &pe) {
                MicroOp mo;
                memset(&mo, 0, sizeof(mo));
                mo.tmask = pe.lcmask & ~regmask;
                mo.data_off = (int32_t)pe.data_off;
                for (int e = 0; e < N; e++)
                    if (elem_ok(pe, e)) mo.emask |= 1u << e;
                return mo;
            };
            // a run of ops multiplied into one N x N matrix (appended to the data array)
            auto emit_full = [&](const std::vector<const Pending *> &run) {
                std::vector<cd> M(size_t(N) * N, cd(0));
                for (int c = 0; c < N; c++) {
                    std::vector<cd> v(N, cd(0));
                    v[c] = 1;
                    for (const Pending *pe : run) {
                        if (internal(*pe)) { apply_host(*pe, v); continue; }
                        Pending q = *pe;  // single op with outer predicates: only its slot part is folded
                        q.lcmask &= regmask;
                        apply_host(q, v);
                    }
                    for (int i = 0; i < N; i++) M[size_t(i) * N + c] = v[i];
                }
                bool real = true;
                for (const cd &z : M) real = real && z.imag() == 0.0;
                MicroOp mo = base_mop(*run[0]);
                if (run.size() > 1) mo.tmask = 0;
                mo.emask = 0xffffffffu;
                mo.kind = real ? MOP_DENSEJ_REAL : MOP_DENSEJ;
                const size_t slots = real ? size_t(N) * N / 2 : size_t(N) * N;
                const int64_t off = int64_t(hdata.size() / esz);
                hdata.resize(hdata.size() + slots * esz);
                if (real) {
                    for (size_t i = 0; i < slots; i++) store_cx(off + int64_t(i), cd(M[2 * i].real(), M[2 * i + 1].real()));
                } else {
                    for (size_t i = 0; i < slots; i++) store_cx(off + int64_t(i), M[i]);
                }
                mo.data_off = (int32_t)off;
                mops.push_back(mo);
                outers.push_back(run[0]->od);
            };
            auto emit_single = [&](const Pending &pe) {
                MicroOp mo = base_mop(pe);
                const bool slot_controls = (pe.lcmask & regmask) != 0;
                if (pe.kind == MOP_DIAG) {
                    mo.kind = MOP_DIAG;
                    mo.nfields = pe.nfields;
                    for (int f = 0; f < pe.nfields; f++)
                        mo.field[f] = uint32_t(pe.fsrc[f]) | (uint32_t(pe.flen[f]) << 8) | (uint32_t(pe.fdst[f]) << 16);
                    for (int e = 0; e < N; e++) {
                        const uint32_t l = off_of(e);
                        uint32_t idx = 0;
                        for (int f = 0; f < pe.nfields; f++)
                            idx |= ((l >> pe.fsrc[f]) & ((1u << pe.flen[f]) - 1)) << pe.fdst[f];
                        mo.u.eidx[e] = (uint16_t)idx;
                    }
                } else if (pe.kind == MOP_DENSE1) {
                    const cd m00 = load_any(hdata, dtype, pe.data_off), m01 = load_any(hdata, dtype, pe.data_off + 1);
                    const cd m10 = load_any(hdata, dtype, pe.data_off + 2), m11 = load_any(hdata, dtype, pe.data_off + 3);
                    const bool real = m00.imag() == 0 && m01.imag() == 0 && m10.imag() == 0 && m11.imag() == 0;
                    const bool is_x = real && m00 == 0.0 && m01 == 1.0 && m10 == 1.0 && m11 == 0.0;
                    mo.ab = slot_of(pe.tpos[0]);
                    if (is_x) {
                        mo.kind = MOP_PERM1;
                    } else if (slot_controls) {  // no masked dense variant: fold the control into an N x N matrix
                        std::vector<const Pending *> one(1, &pe);
                        emit_full(one);
                        return;
                    } else {
                        mo.kind = real ? MOP_DENSE1_REAL : MOP_DENSE1;
                        const unsigned char *src = hdata.data() + size_t(pe.data_off) * esz;
                        if (dtype == QJ_C128 && real) {  // four real parts, compact
                            for (int i = 0; i < 4; i++) memcpy(mo.u.mat + 8 * i, src + 16 * i, 8);
                        } else {
                            memcpy(mo.u.mat, src, 4 * esz);
                        }
                    }
                } else {  // MOP_DENSE2
                    static const double swap_m[16] = {1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1};
                    bool is_swap = true;
                    for (int i = 0; i < 16; i++) is_swap = is_swap && load_any(hdata, dtype, pe.data_off + i) == cd(swap_m[i]);
                    int a = slot_of(pe.tpos[0]), b = slot_of(pe.tpos[1]);
                    if (is_swap) {
                        mo.kind = MOP_PERM2;
                        if (a > b) std::swap(a, b);
                    } else if (slot_controls) {
                        std::vector<const Pending *> one(1, &pe);
                        emit_full(one);
                        return;
                    } else {
                        mo.kind = MOP_DENSE2;
                        if (a > b) {  // canonical slot order: exchange the matrix-index bits in the private copy
                            cd tmp[16];
                            auto sw = [](int i) { return ((i & 1) << 1) | (i >> 1); };
                            for (int i = 0; i < 4; i++)
                                for (int j = 0; j < 4; j++) tmp[sw(i) * 4 + sw(j)] = load_any(hdata, dtype, pe.data_off + i * 4 + j);
                            for (int i = 0; i < 16; i++) store_cx(pe.data_off + i, tmp[i]);
                            std::swap(a, b);
                        }
                    }
                    mo.ab = a | (b << 4);
                }
                mops.push_back(mo);
                outers.push_back(pe.od);
            };
            // consecutive internal ops form a run; two or more are multiplied together
            std::vector<const Pending *> run;
            auto flush_run = [&]() {
                if (run.size() >= 2) emit_full(run);
                else if (run.size() == 1) emit_single(*run[0]);
                run.clear();
            };
            for (const Pending &pe : pend) {
                if (internal(pe)) { run.push_back(&pe); continue; }
                flush_run();
                emit_single(pe);
            }
            flush_run();
            R.count = (int)mops.size() - launch_first_mop - R.first;
            rounds.push_back(R);
            pend.clear();
            regs.clear();
        };

        for (int64_t oi = pd.first_op; oi < pd.first_op + pd.nops; oi++) {
            const qj_op_desc &od = ops[oi];
            Pending pe;
            memset(&pe, 0, sizeof(pe));
            pe.data_off = od.data_offset;
            if (od.ncontrols < 0 || od.ncontrols > QJ_MAX_QUBITS) return bail("op: bad control count");
            uint64_t seen = 0;
            for (int c = 0; c < od.ncontrols; c++) {
                const int b = od.controls[c];
                if (b < 0 || b >= nqubits) return bail("op: control bit out of range");
                if ((seen >> b) & 1) return bail("op: duplicate qubit");
                seen |= uint64_t(1) << b;
                if (lpos[b] >= 0) pe.lcmask |= 1u << lpos[b];
                else pe.ocmask |= uint64_t(1) << b;
            }
            pe.od.ocmask = pe.ocmask;
            if (od.kind == QJ_OPK_DENSE1 || od.kind == QJ_OPK_DENSE2) {
                const int nt = (od.kind == QJ_OPK_DENSE1) ? 1 : 2;
                if (od.ntargets != nt) return bail("op: dense target count mismatch");
                const int64_t need = (nt == 1) ? 4 : 16;
                if (od.data_offset < 0 || od.data_offset + need > ndata) return bail("op: matrix outside the data array");
                pe.kind = (nt == 1) ? MOP_DENSE1 : MOP_DENSE2;
                pe.nt = nt;
                for (int t = 0; t < nt; t++) {
                    const int b = od.targets[t];
                    if (b < 0 || b >= nqubits) return bail("op: target bit out of range");
                    if ((seen >> b) & 1) return bail("op: duplicate qubit");
                    seen |= uint64_t(1) << b;
                    if (lpos[b] < 0) return bail("op: dense target is not a local bit of its pass");
                    pe.tpos[t] = lpos[b];
                }
                std::vector<int> merged = regs;
                for (int t = 0; t < nt; t++)
                    if (std::find(merged.begin(), merged.end(), pe.tpos[t]) == merged.end()) merged.push_back(pe.tpos[t]);
                if ((int)merged.size() > J) {
                    close_round();
                    merged.clear();
                    for (int t = 0; t < nt; t++) merged.push_back(pe.tpos[t]);
                }
                regs = merged;
                pend.push_back(pe);
            } else if (od.kind == QJ_OPK_DIAG) {
                const int nb = od.ntargets;
                if (nb < 0 || nb > QJ_MAX_DIAG_BITS) return bail("op: diagonal table over too many bits");
                if (od.data_offset < 0 || od.data_offset + (int64_t(1) << nb) > ndata) return bail("op: table outside the data array");
                pe.kind = MOP_DIAG;
                pe.ndiag = nb;
                // table bit j <- index bit targets[j]; consecutive local positions become one field
                int nf = 0;
                for (int j = 0; j < nb; j++) {
                    const int b = od.targets[j];
                    if (b < 0 || b >= nqubits) return bail("op: table bit out of range");
                    if ((seen >> b) & 1) return bail("op: duplicate qubit");
                    seen |= uint64_t(1) << b;
                    pe.dpos[j] = lpos[b];
                    if (lpos[b] >= 0) {
                        const int p = lpos[b];
                        if (nf > 0 && pe.fsrc[nf - 1] + pe.flen[nf - 1] == p && pe.fdst[nf - 1] + pe.flen[nf - 1] == j) {
                            pe.flen[nf - 1]++;
                        } else {
                            if (nf == kMaxFields) return bail("op: diagonal table needs too many bit fields; order its bits by position");
                            pe.fsrc[nf] = (uint8_t)p; pe.flen[nf] = 1; pe.fdst[nf] = (uint8_t)j;
                            nf++;
                        }
                    } else {
                        pe.od.src[pe.od.nbits] = (uint8_t)b;
                        pe.od.dst[pe.od.nbits] = (uint8_t)j;
                        pe.od.nbits++;
                    }
                }
                pe.nfields = nf;
                pend.push_back(pe);
            } else {
                return bail("op: unknown kind");
            }
            if ((int)pend.size() >= kMaxMopsPerLaunch / 2) close_round();
        }
        close_round();
        close_launch();
    }

    prog->total_mops = (int64_t)mops.size();
    prog->total_rounds = (int64_t)rounds.size();
    cudaSetDevice(h->device);
    auto upload = [&](void **dst, const void *src, size_t bytes) -> cudaError_t {
        if (bytes == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc(dst, bytes);
        if (e != cudaSuccess) return e;
        return cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, h->stream);
    };
    cudaError_t e = upload(&prog->d_rounds, rounds.data(), rounds.size() * sizeof(Round));
    if (e == cudaSuccess) e = upload(&prog->d_mops, mops.data(), mops.size() * sizeof(MicroOp));
    if (e == cudaSuccess) e = upload(&prog->d_outers, outers.data(), outers.size() * sizeof(OuterDesc));
    if (e == cudaSuccess) e = upload(&prog->d_data, hdata.data(), hdata.size());
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);  // host vectors die with this frame
    if (e != cudaSuccess) {
        qj_program_destroy(h, prog);
        return fail(QJ_ERR_CUDA, std::string("program upload: ") + cudaGetErrorString(e));
    }
    *out = prog;
    return QJ_OK;
}

extern "C" int qj_program_destroy(qj_handle *h, qj_program *p) {
    if (!p) return QJ_OK;
    if (h) {
        cudaSetDevice(h->device);
        cudaStreamSynchronize(h->stream);
    }
    cudaFree(p->d_rounds);
    cudaFree(p->d_mops);
    cudaFree(p->d_outers);
    cudaFree(p->d_data);
    delete p;
    return QJ_OK;
}

extern "C" int qj_program_stats(const qj_program *p, int64_t *nlaunches, int64_t *nrounds, int64_t *nmops) {
    QJ_REQUIRE(p != nullptr, "null program");
    if (nlaunches) *nlaunches = (int64_t)p->launches.size();
    if (nrounds) *nrounds = p->total_rounds;
    if (nmops) *nmops = p->total_mops;
    return QJ_OK;
}

namespace {
template <typename T>
int launch_pass(qj_handle *h, const qj_program *p, void *state, const qj_program::Launch &L) {
    static bool configured = false;
    if (!configured) {
        QJ_CUDA_OK(cudaFuncSetAttribute(k_tile_program<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 << 10));
        configured = true;
    }
    PassGeom pg;
    memset(&pg, 0, sizeof(pg));
    pg.T = L.T; pg.r = L.r; pg.nh = L.nh;
    for (int i = 0; i < L.nh; i++) pg.hibit[i] = L.hibit[i];
    pg.npos = L.npos;
    for (int i = 0; i < L.npos; i++) pg.pos[i] = L.pos[i];
    pg.ntiles = L.ntiles;
    pg.nrounds = L.nrounds;
    pg.nmops = L.nmops;
    const size_t smem = 2 * (size_t(1) << L.T) * sizeof(Cx<T>);   // two tile buffers
    // small tiles: several CTAs per SM; full 64 KiB tiles: one persistent CTA per SM
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (size_t(200) << 10) / (smem + 8192)));
    const unsigned grid = (unsigned)std::min<int64_t>(L.ntiles, int64_t(h->sm_count) * per_sm);
    k_tile_program<T><<<grid, kBlkThreads, smem, h->stream>>>(
        reinterpret_cast<Cx<T> *>(state), pg, reinterpret_cast<const Round *>(p->d_rounds) + L.first_round,
        reinterpret_cast<const MicroOp *>(p->d_mops) + L.first_mop,
        reinterpret_cast<const OuterDesc *>(p->d_outers) + L.first_mop,
        reinterpret_cast<const Cx<T> *>(p->d_data));
    h->launches++;
    QJ_CUDA_OK(cudaGetLastError());
    return QJ_OK;
}

int run_launches(qj_handle *h, const qj_program *p, void *state, int first, int count) {
    for (int li = first; li < first + count; li++) {
        const qj_program::Launch &L = p->launches[li];
        const int rc = (p->dtype == QJ_C128) ? launch_pass<double>(h, p, state, L) : launch_pass<float>(h, p, state, L);
        if (rc) return rc;
    }
    return QJ_OK;
}
}  // namespace

extern "C" int qj_program_run(qj_handle *h, const qj_program *p, void *state) {
    QJ_REQUIRE(h && p && state, "null argument");
    QJ_REQUIRE((reinterpret_cast<uintptr_t>(state) & 15) == 0, "state must be 16-byte aligned");
    return run_launches(h, p, state, 0, (int)p->launches.size());
}

extern "C" int qj_program_run_launch(qj_handle *h, const qj_program *p, void *state, int launch) {
    QJ_REQUIRE(h && p && state, "null argument");
    QJ_REQUIRE(launch >= 0 && launch < (int)p->launches.size(), "launch index out of range");
    return run_launches(h, p, state, launch, 1);
}
