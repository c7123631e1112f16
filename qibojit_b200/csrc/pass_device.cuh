// Device side of the multi-gate tile passes: launch structures, op interpreter and the k_pass
// kernel template.  Included by pass_kernels.cu (host encoder, program objects, the complex128
// instantiation) and pass_kernels_f32.cu (the complex64 instantiation): the two instantiations are
// separate translation units so that they compile in parallel (ptxas spends minutes on each).
#pragma once

#include <cstdint>

#include "common.cuh"

namespace qj {
namespace {

constexpr int kThreads = 256;
constexpr int kVecRegBits = 4;          // a thread holds 2^4 vectors per round
constexpr int kMaxTileVecBits = 12;     // 64 KiB tiles
constexpr int kMinTileBits = 6;
constexpr int kMaxHiBits = 8;
// The program image of a launch travels as a KERNEL PARAMETER (constant bank): op headers and
// gate matrices are read with warp-uniform constant loads (LDCU) into uniform registers and feed
// the FP instructions as uniform operands -- no shared-memory traffic, no vector registers and no
// unpacking for them.  (Kernel parameters may total 32764 bytes.)
constexpr int kMaxBlobUnits = 1984;              // 31 KiB
struct ProgParam {
    uint4 u[kMaxBlobUnits];
};
constexpr int kMaxOuter = 512;

template <typename T>
struct Lay;
template <>
struct Lay<double> {
    static constexpr int J = 4, N = 16, VS = 0;
};
template <>
struct Lay<float> {
    static constexpr int J = 5, N = 32, VS = 1;
};

// dispatch codes
enum {
    // one-target gates on a SET of register slots (h0.y >> 16 = slot mask, one matrix per slot in
    // ascending slot order): consecutive gates of one kind on different slots cost one dispatch
    C_GROUP1C = 0,    // complex 2x2 (8 scalars each)
    C_GROUP1R = 1,    // real 2x2 (4 scalars)
    C_GROUP1X = 2,    // real diagonal, imaginary off-diagonal: RX, Y, sqrt-X up to a phase (4 scalars)
    C_PERM1 = 3,      // + slot (X)
    C_DENSE2 = 8,     // + pair index (a < b): b (b - 1) / 2 + a
    C_PERM2 = 18,     // + pair index (SWAP)
    C_PHASE = 28,     // product of per-thread table look-ups, one complex multiply of the selected elements
    C_DIAGN = 29,     // general: one look-up per element
    C_DENSE2R = 30,   // + pair index: REAL 4x4 (16 scalars): half the multiply-adds of the complex one
    C_DIAGF = 40,     // fused diagonal: per-thread, per-element factors G[tid][unit] (x per-tile factors)
    C_DIAGC = 41,     // constant diagonal: one complex constant per selected unit, in the payload
    C_DIAGS = 42,     // slot-factorised diagonal: one factor per register slot and thread (x per-tile factor)
    C_GROUP1H = 43,   // unnormalised Hadamard butterflies (s0 + s1, s0 - s1) on a set of slots: adds only; the
                      // scale 2^(-1/2) of each is carried by the round's last Hadamard-like gate
    C_DIAGCS = 44,    // + slot: constant diagonal over ALL elements with that register bit set (no mask tests)
};

// element selection of a C_PHASE op (which of a thread's register amplitudes the phase multiplies)
enum {
    SEL_ALL = 0,      // every element
    SEL_SLOT = 1,     // + slot: elements whose register bit `slot` is 1
    SEL_PAIR = 6,     // + pair index: elements whose register bits a and b are both 1
    SEL_MASK = 16,    // arbitrary element mask
};

struct PassGeom {
    int T, r, nh;          // tile bits, run bits (amplitudes), high local bits
    int hibit[kMaxHiBits];
    int npos;
    int pos[QJ_MAX_QUBITS];
    int64_t ntiles;
    int blob_units;        // program image, 16-byte units
    int prefix_units;      // its leading part (header, rounds, outers, H and F entries): copied to shared memory
    int nH;                // per-tile phase factors (outer-only parts of the phase groups)
    int nF;                // fused diagonals with an outer part (2^J per-tile, per-element factors each)
    int nFS;               // their slices in total (one table look-up per slice and tile)
    // tile IO of the one-thread-per-16-vectors launch: vector u * nthreads + tid of the tile lives at
    // global vector  tile base + thread part + io_goff[u]  and at shared vector  swz(tid) ^ io_soff[u]
    // (the swizzle is XOR-linear, the run index splits into a thread and a per-iteration part):
    // both per-iteration parts are launch constants, read as constant-bank operands
    int64_t io_goff[16];
    uint32_t io_soff[16];
    int io_fast;           // the launch geometry satisfies the conditions above
    int nrounds_smem;      // rounds whose per-thread constants are staged in shared memory
    int zero_input;        // the input is |0...0>: the launch does not read the state (state preparation fused in)
    int64_t tile_begin, tile_end;   // the tiles this launch processes (all of them unless a caller pipelines sub-blocks)
    int64_t out_off;                // where the tiles are stored, in 16-byte vectors from the state pointer (0: in place)
};

// ---- program image (16-byte units) -------------------------------------------------------------
// unit 0            : {nrounds, nouter, off_rounds, off_outer}
// unit 1            : {nH, off_H, nF, off_F}
// F region          : nF directory units {first slice unit (relative to off_F), nslices, 0, 0}, then the
//                     slices {element mask, table, oslot, 0}: per-tile factor of element e of fused
//                     diagonal f = product over its slices whose mask holds e of table[outer index]
// H entries, 4 units: {ntab, 0, 0, 0} {table, oslot, table, oslot} x 3   (product of <= 6 outer-indexed look-ups)
// rounds, 3 units   : {first_unit, nops, vd0 | vd1 << 16, vd2 | vd3 << 16}
//                     {td[0..7] as uint16}  {tpos[0..7] as uint8, 0, 0}
// outers, 2 units   : {ocmask lo, ocmask hi, nbits, src[0..3]} {src[4..11], dst packed 4 bit x 12 ...}
// ops               : 2-unit header + payload
struct HostOuter {
    uint64_t ocmask = 0;
    int nbits = 0;
    uint8_t src[12] = {0}, dst[12] = {0};
};

__host__ __device__ __forceinline__ uint32_t swz_vec(uint32_t v) {
    return v ^ (((v >> 3) ^ (v >> 6) ^ (v >> 9)) & 7u);
}

template <typename T>
__device__ __forceinline__ void cmul_acc(T &ar, T &ai, T gr, T gi, T xr, T xi) {
    ar = fma(gr, xr, ar);
    ar = fma(-gi, xi, ar);
    ai = fma(gr, xi, ai);
    ai = fma(gi, xr, ai);
}

__device__ __forceinline__ constexpr int insert0(int p, int a) { return ((p >> a) << (a + 1)) | (p & ((1 << a) - 1)); }

// payload readers: NU 16-byte units of the program image (constant bank, warp-uniform index)
template <int NU>
__device__ __forceinline__ void load_units(const ProgParam &pp, int p, double (&d)[2 * NU]) {
#pragma unroll
    for (int i = 0; i < NU; i++) {
        const uint4 q = pp.u[p + i];
        d[2 * i] = __hiloint2double(int(q.y), int(q.x));
        d[2 * i + 1] = __hiloint2double(int(q.w), int(q.z));
    }
}
template <int NU>
__device__ __forceinline__ void load_units(const ProgParam &pp, int p, float (&d)[4 * NU]) {
#pragma unroll
    for (int i = 0; i < NU; i++) {
        const uint4 q = pp.u[p + i];
        d[4 * i] = __uint_as_float(q.x); d[4 * i + 1] = __uint_as_float(q.y);
        d[4 * i + 2] = __uint_as_float(q.z); d[4 * i + 3] = __uint_as_float(q.w);
    }
}

// ---- ops on the register amplitudes ---------------------------------------------------------------
// ---- packed FP32x2 arithmetic (sm_100a FFMA2 / FMUL2): a complex64 amplitude is one 64-bit
// operand.  ptxas folds the packing below into operand modifiers (scalar broadcast, swapped
// halves, per-half negation), so a complex multiply-accumulate is two instructions, no moves.
typedef unsigned long long u64;
__device__ __forceinline__ u64 f2_fma(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 f2_mul(u64 a, u64 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 f2_pack(float lo, float hi) {
    u64 d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ u64 f2_of(const Cx<float> &c) { return f2_pack(c.re, c.im); }
__device__ __forceinline__ u64 f2_swapped(const Cx<float> &c) { return f2_pack(c.im, c.re); }
__device__ __forceinline__ u64 f2_splat(float v) { return f2_pack(v, v); }
// a pair that must be built once and kept (not rematerialised at every use)
__device__ __forceinline__ u64 f2_pack_once(float lo, float hi) {
    u64 d;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void f2_store(Cx<float> &c, u64 v) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c.re), "=f"(c.im) : "l"(v));
}
// A complex64 matrix element g travels as the two pairs S = (gr, gr), N = (-gi, gi) (complex64
// payloads store them ready-made, 16 bytes per element): g * s = S * s + N * swapped(s).
__device__ __forceinline__ u64 f2_cmac(u64 acc, u64 S, u64 Nn, const Cx<float> &s) {
    acc = f2_fma(S, f2_of(s), acc);
    return f2_fma(Nn, f2_swapped(s), acc);
}
__device__ __forceinline__ u64 f2_cmul(u64 S, u64 Nn, const Cx<float> &s) {
    return f2_fma(Nn, f2_swapped(s), f2_mul(S, f2_of(s)));
}

// one-target gate on register slot A.  KIND 0: complex 2x2 (8 scalars), 1: real (4 scalars),
// 2: real diagonal + imaginary off-diagonal, payload {g00, Im g01, Im g10, g11}.
// The unmasked path (no register-slot controls) is straight-line code updating x in place.
template <typename T, int A, int KIND>
__device__ __forceinline__ void dense1_pair(Cx<T> &s0, Cx<T> &s1, const T *m) {
    if constexpr (sizeof(T) == 4) {
        u64 y0, y1;
        if (KIND == 1) {
            y0 = f2_fma(f2_splat(m[0]), f2_of(s0), f2_mul(f2_splat(m[1]), f2_of(s1)));
            y1 = f2_fma(f2_splat(m[3]), f2_of(s1), f2_mul(f2_splat(m[2]), f2_of(s0)));
        } else if (KIND == 2) {   // payload {a, d, -b, b, -c, c, 0, 0}: i b s = (-b, b) * swapped(s)
            y0 = f2_fma(f2_splat(m[0]), f2_of(s0), f2_mul(f2_pack(m[2], m[3]), f2_swapped(s1)));
            y1 = f2_fma(f2_splat(m[1]), f2_of(s1), f2_mul(f2_pack(m[4], m[5]), f2_swapped(s0)));
        } else {                  // payload {gr, gr, -gi, gi} per element
            y0 = f2_cmac(f2_cmul(f2_pack(m[0], m[1]), f2_pack(m[2], m[3]), s0), f2_pack(m[4], m[5]), f2_pack(m[6], m[7]), s1);
            y1 = f2_cmac(f2_cmul(f2_pack(m[8], m[9]), f2_pack(m[10], m[11]), s0), f2_pack(m[12], m[13]), f2_pack(m[14], m[15]), s1);
        }
        f2_store(s0, y0);
        f2_store(s1, y1);
        return;
    }
    if (KIND == 1) {
        const T t0 = m[1] * s1.re, t1 = m[1] * s1.im, u0 = m[2] * s0.re, u1 = m[2] * s0.im;
        s0.re = fma(m[0], s0.re, t0); s0.im = fma(m[0], s0.im, t1);
        s1.re = fma(m[3], s1.re, u0); s1.im = fma(m[3], s1.im, u1);
    } else if (KIND == 2) {
        const T t0 = m[1] * s1.im, t1 = m[1] * s1.re, u0 = m[2] * s0.im, u1 = m[2] * s0.re;
        s0.re = fma(m[0], s0.re, -t0); s0.im = fma(m[0], s0.im, t1);
        s1.re = fma(m[3], s1.re, -u0); s1.im = fma(m[3], s1.im, u1);
    } else {
        T ar = m[0] * s0.re, ai = m[0] * s0.im;
        ar = fma(-m[1], s0.im, ar); ai = fma(m[1], s0.re, ai);
        cmul_acc(ar, ai, m[2], m[3], s1.re, s1.im);
        T br = m[4] * s0.re, bi = m[4] * s0.im;
        br = fma(-m[5], s0.im, br); bi = fma(m[5], s0.re, bi);
        cmul_acc(br, bi, m[6], m[7], s1.re, s1.im);
        s0.re = ar; s0.im = ai; s1.re = br; s1.im = bi;
    }
}

template <typename T, int A, int KIND>
__device__ __forceinline__ void op_dense1(Cx<T> (&x)[Lay<T>::N], const ProgParam &pp, int pay, uint32_t emask) {
    constexpr int N = Lay<T>::N;
    if constexpr (A < Lay<T>::J) {
        // scalars in the payload (a whole number of units); complex64 complex elements are 4 floats
        constexpr int NS = (sizeof(T) == 4) ? (KIND == 0 ? 16 : KIND == 2 ? 8 : 4) : (KIND == 0 ? 8 : 4);
        T m[NS];
        load_units<NS * sizeof(T) / 16>(pp, pay, m);
        if (emask == (N == 32 ? 0xffffffffu : 0xffffu)) {
#pragma unroll
            for (int p = 0; p < N / 2; p++) {
                const int e0 = insert0(p, A), e1 = e0 | (1 << A);
                dense1_pair<T, A, KIND>(x[e0], x[e1], m);
            }
        } else {
#pragma unroll
            for (int p = 0; p < N / 2; p++) {
                const int e0 = insert0(p, A), e1 = e0 | (1 << A);
                if (!((emask >> e0) & 1u)) continue;
                dense1_pair<T, A, KIND>(x[e0], x[e1], m);
            }
        }
    }
}

template <typename T, int KIND>
__device__ __forceinline__ void op_group1(Cx<T> (&x)[Lay<T>::N], const ProgParam &pp, int pay, uint32_t slots, uint32_t emask) {
    constexpr int NU = (sizeof(T) == 4) ? (KIND == 0 ? 4 : KIND == 2 ? 2 : 1) : (KIND == 0 ? 4 : 2);   // units per matrix
    if (slots & 1u) { op_dense1<T, 0, KIND>(x, pp, pay, emask); pay += NU; }
    if (slots & 2u) { op_dense1<T, 1, KIND>(x, pp, pay, emask); pay += NU; }
    if (slots & 4u) { op_dense1<T, 2, KIND>(x, pp, pay, emask); pay += NU; }
    if (slots & 8u) { op_dense1<T, 3, KIND>(x, pp, pay, emask); pay += NU; }
    if (slots & 16u) { op_dense1<T, 4, KIND>(x, pp, pay, emask); }
}

// two-target gate: matrix-index bit 0 <-> slot A, bit 1 <-> slot B (A < B)
template <typename T, int A, int B>
__device__ __forceinline__ void dense2_group(Cx<T> (&x)[Lay<T>::N], const ProgParam &pp, int pay, int e0) {
    constexpr int RU = 8 * sizeof(T) / 16;           // units per matrix row (4 complex)
    Cx<T> s[4], y[4];
#pragma unroll
    for (int j = 0; j < 4; j++) s[j] = x[e0 | ((j & 1) << A) | ((j >> 1) << B)];
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float g[16];                                  // row i: {gr, gr, -gi, gi} x 4
            load_units<4>(pp, pay + i * 4, g);
            u64 acc = f2_cmul(f2_pack(g[0], g[1]), f2_pack(g[2], g[3]), s[0]);
#pragma unroll
            for (int j = 1; j < 4; j++)
                acc = f2_cmac(acc, f2_pack(g[4 * j], g[4 * j + 1]), f2_pack(g[4 * j + 2], g[4 * j + 3]), s[j]);
            f2_store(y[i], acc);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) x[e0 | ((i & 1) << A) | ((i >> 1) << B)] = y[i];
        return;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        T g[8];
        load_units<RU>(pp, pay + i * RU, g);
        T ar = g[0] * s[0].re, ai = g[0] * s[0].im;
        ar = fma(-g[1], s[0].im, ar); ai = fma(g[1], s[0].re, ai);
#pragma unroll
        for (int j = 1; j < 4; j++) cmul_acc(ar, ai, g[2 * j], g[2 * j + 1], s[j].re, s[j].im);
        y[i].re = ar; y[i].im = ai;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) x[e0 | ((i & 1) << A) | ((i >> 1) << B)] = y[i];
}

template <typename T, int A, int B>
__device__ __forceinline__ void op_dense2(Cx<T> (&x)[Lay<T>::N], const ProgParam &pp, int pay, uint32_t emask) {
    constexpr int N = Lay<T>::N;
    if constexpr (B < Lay<T>::J) {
        if (emask == (N == 32 ? 0xffffffffu : 0xffffu)) {
#pragma unroll
            for (int p = 0; p < N / 4; p++) dense2_group<T, A, B>(x, pp, pay, insert0(insert0(p, A), B));
        } else {
#pragma unroll
            for (int p = 0; p < N / 4; p++) {
                const int e0 = insert0(insert0(p, A), B);
                if (!((emask >> e0) & 1u)) continue;
                dense2_group<T, A, B>(x, pp, pay, e0);
            }
        }
    }
}

// two-target gate with a REAL matrix (RY RY CZ RY RY on a pair, products of Hadamards and CZ, ...):
// re and im parts transform separately, 8 multiply-adds per amplitude instead of 16
template <typename T, int A, int B>
__device__ __forceinline__ void dense2r_group(Cx<T> (&x)[Lay<T>::N], const ProgParam &pp, int pay, int e0) {
    // "diagonal last": acc_i = sum_{j != i} g_ij s_j for every row first (all inputs still alive),
    // then s_i = g_ii s_i + acc_i in place -- every result lands in the register its input came
    // from, so the op loop needs no register shuffling at its back edge
    constexpr int E0 = 0, E1 = 1 << A, E2 = 1 << B, E3 = (1 << A) | (1 << B);
    const int idx[4] = {e0 | E0, e0 | E1, e0 | E2, e0 | E3};
    if constexpr (sizeof(T) == 4) {
        u64 acc[4];
        float gd[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float g[4];                                   // row i
            load_units<1>(pp, pay + i, g);
            gd[i] = g[i];
            bool first = true;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (j == i) continue;
                acc[i] = first ? f2_mul(f2_splat(g[j]), f2_of(x[idx[j]])) : f2_fma(f2_splat(g[j]), f2_of(x[idx[j]]), acc[i]);
                first = false;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) f2_store(x[idx[i]], f2_fma(f2_splat(gd[i]), f2_of(x[idx[i]]), acc[i]));
    } else {
        T ar[4], ai[4], gd[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            T g[4];
            load_units<2>(pp, pay + 2 * i, g);
            gd[i] = g[i];
            bool first = true;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (j == i) continue;
                if (first) { ar[i] = g[j] * x[idx[j]].re; ai[i] = g[j] * x[idx[j]].im; }
                else { ar[i] = fma(g[j], x[idx[j]].re, ar[i]); ai[i] = fma(g[j], x[idx[j]].im, ai[i]); }
                first = false;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            x[idx[i]].re = fma(gd[i], x[idx[i]].re, ar[i]);
            x[idx[i]].im = fma(gd[i], x[idx[i]].im, ai[i]);
        }
    }
}

template <typename T, int A, int B>
__device__ __forceinline__ void op_dense2r(Cx<T> (&x)[Lay<T>::N], const ProgParam &pp, int pay, uint32_t emask) {
    constexpr int N = Lay<T>::N;
    if constexpr (B < Lay<T>::J) {
        if (emask == (N == 32 ? 0xffffffffu : 0xffffu)) {
#pragma unroll
            for (int p = 0; p < N / 4; p++) dense2r_group<T, A, B>(x, pp, pay, insert0(insert0(p, A), B));
        } else {
#pragma unroll
            for (int p = 0; p < N / 4; p++) {
                const int e0 = insert0(insert0(p, A), B);
                if (!((emask >> e0) & 1u)) continue;
                dense2r_group<T, A, B>(x, pp, pay, e0);
            }
        }
    }
}

template <typename T, int A>
__device__ __forceinline__ void op_perm1(Cx<T> (&x)[Lay<T>::N], uint32_t emask) {
    constexpr int N = Lay<T>::N;
    if constexpr (A < Lay<T>::J) {
#pragma unroll
        for (int p = 0; p < N / 2; p++) {
            const int e0 = insert0(p, A), e1 = e0 | (1 << A);
            if (!((emask >> e0) & 1u)) continue;
            const Cx<T> t = x[e0]; x[e0] = x[e1]; x[e1] = t;
        }
    }
}

template <typename T, int A, int B>
__device__ __forceinline__ void op_perm2(Cx<T> (&x)[Lay<T>::N], uint32_t emask) {
    constexpr int N = Lay<T>::N;
    if constexpr (B < Lay<T>::J) {
#pragma unroll
        for (int p = 0; p < N / 4; p++) {
            const int e0 = insert0(insert0(p, A), B);
            if (!((emask >> e0) & 1u)) continue;
            const int ea = e0 | (1 << A), eb = e0 | (1 << B);
            const Cx<T> t = x[ea]; x[ea] = x[eb]; x[eb] = t;
        }
    }
}

template <typename T>
__device__ __forceinline__ void cmul_inplace(Cx<T> &v, T pr, T pi) {
    if constexpr (sizeof(T) == 4) {
        f2_store(v, f2_cmul(f2_splat(pr), f2_pack(-pi, pi), v));
        return;
    }
    const T t = pi * v.im, u = pi * v.re;
    v.re = fma(pr, v.re, -t);
    v.im = fma(pr, v.im, u);
}

template <typename T>
__device__ __forceinline__ void mul_masked(Cx<T> (&x)[Lay<T>::N], uint32_t emask, T pr, T pi) {
#pragma unroll
    for (int e = 0; e < Lay<T>::N; e++) {
        if (!((emask >> e) & 1u)) continue;
        cmul_inplace<T>(x[e], pr, pi);
    }
}

__device__ __forceinline__ double flip(double v, uint32_t s) {
    return __hiloint2double(int(uint32_t(__double2hiint(v)) ^ s), __double2loint(v));
}
__device__ __forceinline__ float flip(float v, uint32_t s) { return __uint_as_float(__float_as_uint(v) ^ s); }

template <typename T>
__device__ __forceinline__ void flip_masked(Cx<T> (&x)[Lay<T>::N], uint32_t emask, uint32_t s) {
#pragma unroll
    for (int e = 0; e < Lay<T>::N; e++) {
        if (!((emask >> e) & 1u)) continue;
        x[e].re = flip(x[e].re, s);
        x[e].im = flip(x[e].im, s);
    }
}

// phase (SIGN = false) or sign flip (SIGN = true) of the elements with every bit of the static
// mask BITS set: straight-line code
template <typename T, int BITS, bool SIGN>
__device__ __forceinline__ void phase_static(Cx<T> (&x)[Lay<T>::N], T pr, T pi, uint32_t sg) {
    if constexpr (BITS < Lay<T>::N) {
        if constexpr (sizeof(T) == 4 && !SIGN) {
            const u64 S = f2_pack_once(pr, pr), Nn = f2_pack_once(-pi, pi);
#pragma unroll
            for (int e = 0; e < Lay<T>::N; e++) {
                if ((e & BITS) != BITS) continue;
                f2_store(x[e], f2_cmul(S, Nn, x[e]));
            }
        } else {
#pragma unroll
            for (int e = 0; e < Lay<T>::N; e++) {
                if ((e & BITS) != BITS) continue;
                if (SIGN) { x[e].re = flip(x[e].re, sg); x[e].im = flip(x[e].im, sg); }
                else cmul_inplace<T>(x[e], pr, pi);
            }
        }
    }
}

template <typename T, bool SIGN>
__device__ __forceinline__ void phase_apply(Cx<T> (&x)[Lay<T>::N], uint32_t sel, uint32_t emask, T pr, T pi,
                                            uint32_t sg) {
    switch (sel) {
#define QJ_PH(CODE, BITS) case CODE: phase_static<T, BITS, SIGN>(x, pr, pi, sg); break;
        QJ_PH(SEL_ALL, 0)
        QJ_PH(SEL_SLOT + 0, 1) QJ_PH(SEL_SLOT + 1, 2) QJ_PH(SEL_SLOT + 2, 4) QJ_PH(SEL_SLOT + 3, 8)
        QJ_PH(SEL_SLOT + 4, 16)
        QJ_PH(SEL_PAIR + 0, 3) QJ_PH(SEL_PAIR + 1, 5) QJ_PH(SEL_PAIR + 2, 6) QJ_PH(SEL_PAIR + 3, 9)
        QJ_PH(SEL_PAIR + 4, 10) QJ_PH(SEL_PAIR + 5, 12) QJ_PH(SEL_PAIR + 6, 17) QJ_PH(SEL_PAIR + 7, 18)
        QJ_PH(SEL_PAIR + 8, 20) QJ_PH(SEL_PAIR + 9, 24)
#undef QJ_PH
        default:
            if (SIGN) flip_masked<T>(x, emask, sg);
            else mul_masked<T>(x, emask, pr, pi);
    }
}

__device__ __forceinline__ uint32_t sign_of(double v) { return uint32_t(__double2hiint(v)) & 0x80000000u; }
__device__ __forceinline__ uint32_t sign_of(float v) { return __float_as_uint(v) & 0x80000000u; }

// phase-table / factor loads: read-only path, kept in L1 against the streaming tile traffic
__device__ __forceinline__ Cx<double> ldg_cx(const Cx<double> *p) {
    Cx<double> c;
    asm("ld.global.nc.L1::evict_last.v2.f64 {%0, %1}, [%2];" : "=d"(c.re), "=d"(c.im) : "l"(p));
    return c;
}
__device__ __forceinline__ Cx<float> ldg_cx(const Cx<float> *p) {
    Cx<float> c;
    asm("ld.global.nc.L1::evict_last.v2.f32 {%0, %1}, [%2];" : "=f"(c.re), "=f"(c.im) : "l"(p));
    return c;
}

__device__ __forceinline__ float4 ldg_f4(const void *p) {
    float4 c;
    asm("ld.global.nc.L1::evict_last.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w) : "l"(p));
    return c;
}

template <typename T>
__device__ __forceinline__ Cx<T> cx_mul(const Cx<T> &a, const Cx<T> &b) {
    Cx<T> o;
    o.re = fma(a.re, b.re, -(a.im * b.im));
    o.im = fma(a.re, b.im, a.im * b.re);
    return o;
}

// Fused diagonal: every diagonal gate waiting at this point of the round whose other bits are all
// inside the tile (-> G, one host-made factor per 16-byte unit and thread, laid out [thread][unit]:
// a thread's 16 units are 256 contiguous bytes, every load has an immediate offset; L1/L2 resident)
// or all outside it (-> F, one factor per element and tile, made at tile start) in ONE op:
// x[e] *= G[tid][unit(e)] * F[e].  `um`: the units that hold a non-trivial factor (FULL: all).
template <typename T, bool HASF, bool FULL>
__device__ __forceinline__ void op_diagf(Cx<T> (&x)[Lay<T>::N], const Cx<T> *__restrict__ gp, uint32_t um, bool has_g,
                                         const Cx<T> *fp) {
    constexpr int UPE = sizeof(T) == 8 ? 1 : 2;    // elements per unit
#pragma unroll
    for (int c = 0; c < 16; c += 8) {
        Cx<T> z[8 * UPE];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const bool on = FULL || ((um >> (c + k)) & 1u);
            if constexpr (sizeof(T) == 8) {
                z[k].re = T(1); z[k].im = T(0);
                if (has_g && on) z[k] = ldg_cx(gp + (c + k));
            } else {
                float4 q = make_float4(1.f, 0.f, 1.f, 0.f);
                if (has_g && on) q = ldg_f4(gp + 2 * (c + k));
                z[2 * k].re = q.x; z[2 * k].im = q.y; z[2 * k + 1].re = q.z; z[2 * k + 1].im = q.w;
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (!FULL && !((um >> (c + k)) & 1u)) continue;
#pragma unroll
            for (int j = 0; j < UPE; j++) {
                const int e = (c + k) * UPE + j;
                Cx<T> w = z[k * UPE + j];
                if (HASF) w = cx_mul<T>(w, fp[e]);
                cmul_inplace<T>(x[e], w.re, w.im);
            }
        }
    }
}

// Slot-factorised diagonal: the phase groups waiting at one point of a round that each multiply
// "the elements whose register bit s is 1" (the ladders of controlled phases behind the H gates of
// a QFT round) in ONE op: the per-thread factors of all slots are 64 contiguous bytes per thread
// (all loads issued together, tables a few KiB: L1 resident), the per-tile factors come from the
// fused-diagonal machinery (entry 2^s of the op's per-element array), then x[e] *= Q_s for every
// element with bit s set.  Same arithmetic as one phase group per slot, one dispatch and one
// memory latency instead of J.
template <typename T, bool HASF>
__device__ __forceinline__ void op_diags(Cx<T> (&x)[Lay<T>::N], const Cx<T> *__restrict__ gp, uint32_t sm, bool has_g,
                                         const Cx<T> *fp) {
    constexpr int J = Lay<T>::J, N = Lay<T>::N;
    Cx<T> q[J];
#pragma unroll
    for (int s = 0; s < J; s++) {
        q[s].re = T(1); q[s].im = T(0);
        if (has_g && ((sm >> s) & 1u)) q[s] = ldg_cx(gp + s);
    }
    if (HASF) {
#pragma unroll
        for (int s = 0; s < J; s++)
            if ((sm >> s) & 1u) q[s] = cx_mul<T>(q[s], fp[1 << s]);
    }
#pragma unroll
    for (int s = 0; s < J; s++) {
        if (!((sm >> s) & 1u)) continue;
#pragma unroll
        for (int e = 0; e < N; e++)
            if ((e >> s) & 1) cmul_inplace<T>(x[e], q[s].re, q[s].im);
    }
}

// Unnormalised Hadamard on register slot A: (s0, s1) <- (s0 + s1, s0 - s1), adds only.
template <typename T, int A>
__device__ __forceinline__ void op_hadamard(Cx<T> (&x)[Lay<T>::N]) {
    constexpr int N = Lay<T>::N;
    if constexpr (A < Lay<T>::J) {
#pragma unroll
        for (int p = 0; p < N / 2; p++) {
            const int e0 = insert0(p, A), e1 = e0 | (1 << A);
            const Cx<T> a = x[e0], b = x[e1];
            x[e0].re = a.re + b.re; x[e0].im = a.im + b.im;
            x[e1].re = a.re - b.re; x[e1].im = a.im - b.im;
        }
    }
}

// Constant diagonal over the N / 2 elements whose register bit A is 1, one constant each (1 for the
// elements the gates leave alone): straight-line code, no mask tests.
template <typename T, int A>
__device__ __forceinline__ void op_diagc_slot(Cx<T> (&x)[Lay<T>::N], const ProgParam &pp, int pay) {
    constexpr int N = Lay<T>::N;
    if constexpr (A < Lay<T>::J) {
#pragma unroll
        for (int p = 0; p < N / 2; p++) {
            const int e = insert0(p, A) | (1 << A);
            if constexpr (sizeof(T) == 8) {
                const uint4 q = pp.u[pay + p];
                cmul_inplace<T>(x[e], __hiloint2double(int(q.y), int(q.x)), __hiloint2double(int(q.w), int(q.z)));
            } else {
                const uint4 q = pp.u[pay + (p >> 1)];
                if (p & 1) cmul_inplace<T>(x[e], __uint_as_float(q.z), __uint_as_float(q.w));
                else cmul_inplace<T>(x[e], __uint_as_float(q.x), __uint_as_float(q.y));
            }
        }
    }
}

// Constant diagonal: factors that depend on the register bits only (the phases between the H
// gates of a QFT round, CZ / CU1 inside a register block): one complex constant per selected
// 16-byte unit, read from the program image as uniform operands -- no memory traffic at all.
template <typename T>
__device__ __forceinline__ void op_diagc(Cx<T> (&x)[Lay<T>::N], const ProgParam &pp, int pay, uint32_t um) {
    constexpr int UPE = sizeof(T) == 8 ? 1 : 2;
#pragma unroll
    for (int u = 0; u < 16; u++) {
        if (!((um >> u) & 1u)) continue;
        const uint4 q = pp.u[pay];
        pay++;
        if constexpr (sizeof(T) == 8) {
            cmul_inplace<T>(x[u], __hiloint2double(int(q.y), int(q.x)), __hiloint2double(int(q.w), int(q.z)));
        } else {
            cmul_inplace<T>(x[2 * u], __uint_as_float(q.x), __uint_as_float(q.y));
            cmul_inplace<T>(x[2 * u + 1], __uint_as_float(q.z), __uint_as_float(q.w));
        }
    }
}

__device__ __forceinline__ int field_of(uint32_t base, uint32_t fl) {
    return int(((base >> (fl & 255u)) & ((1u << ((fl >> 8) & 255u)) - 1u)) << (fl >> 16));
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                     static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
                 "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// tile constant of outer slot m: -1 when its outer controls are not satisfied by this tile, else
// the outer part of the table index
__device__ __forceinline__ int32_t outer_value(const uint4 *outers, int m, int64_t base_amp) {
    const uint4 o0 = outers[2 * m], o1 = outers[2 * m + 1];
    const uint64_t ocmask = uint64_t(o0.x) | (uint64_t(o0.y) << 32);
    if ((uint64_t(base_amp) & ocmask) != ocmask) return -1;
    int32_t v = 0;
    const int nb = int(o0.z);
    const uint32_t srcw[3] = {o0.w, o1.x, o1.y};
    const uint32_t dstw[2] = {o1.z, o1.w};
    for (int b = 0; b < nb; b++) {
        const int src = (srcw[b >> 2] >> ((b & 3) * 8)) & 255;
        const int dst = (dstw[b >> 3] >> ((b & 7) * 4)) & 15;
        v |= int32_t((base_amp >> src) & 1) << dst;
    }
    return v;
}

#define QJ_SLOT_CASES(CODE, CALL)                  \
    case CODE + 0: { CALL(0); } break;             \
    case CODE + 1: { CALL(1); } break;             \
    case CODE + 2: { CALL(2); } break;             \
    case CODE + 3: { CALL(3); } break;             \
    case CODE + 4: { CALL(4); } break;
#define QJ_PAIR_CASES(CODE, CALL)                  \
    case CODE + 0: { CALL(0, 1); } break;          \
    case CODE + 1: { CALL(0, 2); } break;          \
    case CODE + 2: { CALL(1, 2); } break;          \
    case CODE + 3: { CALL(0, 3); } break;          \
    case CODE + 4: { CALL(1, 3); } break;          \
    case CODE + 5: { CALL(2, 3); } break;          \
    case CODE + 6: { CALL(0, 4); } break;          \
    case CODE + 7: { CALL(1, 4); } break;          \
    case CODE + 8: { CALL(2, 4); } break;          \
    case CODE + 9: { CALL(3, 4); } break;

template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
k_pass(Cx<T> *__restrict__ state, const __grid_constant__ PassGeom pg, const Cx<T> *__restrict__ tables,
       const __grid_constant__ ProgParam pp) {
    constexpr int N = Lay<T>::N;
    constexpr int VS = Lay<T>::VS;
    extern __shared__ __align__(128) unsigned char smem_raw[];

    const int tid = threadIdx.x;
    const int Tv = pg.T - VS;                     // tile bits in vectors
    const int nvec = 1 << Tv;
    const int rv = pg.r - VS;                     // run bits in vectors
    const int rvmask = (1 << rv) - 1;
    uint4 *const tilev = reinterpret_cast<uint4 *>(smem_raw);
    uint4 *const prog = tilev + nvec;
    uint4 *const s_H = prog + pg.prefix_units;                                      // one 16-byte slot per factor
    Cx<T> *const s_F = reinterpret_cast<Cx<T> *>(s_H + pg.nH);                      // N factors per fused diagonal
    uint4 *const s_FS = reinterpret_cast<uint4 *>(s_F + pg.nF * N);                 // one 16-byte slot per slice look-up
    uint2 *const s_rconst = reinterpret_cast<uint2 *>(s_FS + pg.nFS);               // [round][thread] {S, base}
    int64_t *const s_runoff = reinterpret_cast<int64_t *>(s_rconst + pg.nrounds_smem * int(blockDim.x));   // in vectors
    int32_t *const s_outer = reinterpret_cast<int32_t *>(s_runoff + (1 << pg.nh));
    uint4 *const gvec = reinterpret_cast<uint4 *>(state);

    const int nthr = int(blockDim.x);
    // the descriptors that threads index individually (outer slots, H and F entries) live in
    // shared memory; the op stream stays in the constant bank
    for (int i = tid; i < pg.prefix_units; i += nthr) prog[i] = pp.u[i];
    for (int run = tid; run < (1 << pg.nh); run += nthr) {
        int64_t off = 0;
        for (int b = 0; b < pg.nh; b++) off |= int64_t((run >> b) & 1) << (pg.hibit[b] - VS);
        s_runoff[run] = off;
    }
    __syncthreads();
    const uint4 hdr = pp.u[0], hdr1 = pp.u[1];
    const int nrounds = int(hdr.x), nouter = int(hdr.y);
    const int rounds = int(hdr.z);                 // unit index of the round descriptors (constant bank)
    const uint4 *const outers = prog + hdr.w;
    const int nH = int(hdr1.x);
    const uint4 *const hents = prog + hdr1.y;
    const int nF = int(hdr1.z);
    const uint4 *const fents = prog + hdr1.w;

    // a thread's place in the tile is the same in every tile: the shared-memory offset of its
    // first register element (S) and its local position bits (base) are computed once per round
    // for the whole launch
    for (int rd = 0; rd < pg.nrounds_smem; rd++) {
        const uint4 r1 = pp.u[rounds + 3 * rd + 1], r2 = pp.u[rounds + 3 * rd + 2];
        const uint32_t tdw[4] = {r1.x, r1.y, r1.z, r1.w};
        uint32_t S = 0, base = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if ((tid >> k) & 1) {
                S ^= (tdw[k >> 1] >> ((k & 1) * 16)) & 0xffffu;
                base |= 1u << (((k < 4 ? r2.x : r2.y) >> ((k & 3) * 8)) & 255u);
            }
        }
        s_rconst[rd * nthr + tid] = make_uint2(S, base);
    }

    // (the launch uses exactly one thread per 16 vectors of the tile: every thread is live)
    const bool fast_io = pg.io_fast != 0;
    const uint32_t sw_t = swz_vec(uint32_t(tid));
    const int lane_off = tid & rvmask, run_t = tid >> rv;
    // this thread's part of every global vector address of the fast tile IO
    const int64_t io_thr = fast_io ? int64_t(lane_off) + s_runoff[run_t] : 0;

    for (int64_t tile_id = pg.tile_begin + blockIdx.x; tile_id < pg.tile_end; tile_id += gridDim.x) {
        // tile base: insert zeros at the high local bits
        int64_t tb = tile_id;
#pragma unroll 1
        for (int j = 0; j < pg.npos; j++) {
            const int p = pg.pos[j];
            tb = ((tb >> p) << (p + 1)) | (tb & ((int64_t(1) << p) - 1));
        }
        const int64_t base_amp = tb << pg.r;
        const int64_t base_vec = base_amp >> VS;

        // ---- load (asynchronous copies straight into the swizzled tile)
        if (pg.zero_input) {
            // |0...0> input (ops.py:14-18 fused into the first pass): nothing to read -- the tile is zero
            // except amplitude 0 of tile 0, which the thread that owns vector 0 sets after its own fill
            const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
            for (int lv = tid; lv < nvec; lv += nthr) tilev[swz_vec(uint32_t(lv))] = zero;
            if (tile_id == 0 && tid == 0) {
                if constexpr (sizeof(T) == 8) *reinterpret_cast<double2 *>(tilev) = make_double2(1.0, 0.0);
                else *reinterpret_cast<float4 *>(tilev) = make_float4(1.f, 0.f, 0.f, 0.f);
            }
        } else if (fast_io) {
            // vector u * nthr + tid: the swizzle is XOR-linear and nthr a multiple of the run length,
            // so the per-thread and the per-iteration (warp-uniform) parts separate
            const uint4 *const gsrc = gvec + base_vec + io_thr;
#pragma unroll
            for (int u = 0; u < 16; u++) cp_async16(tilev + (sw_t ^ pg.io_soff[u]), gsrc + pg.io_goff[u]);
        } else {
            for (int lv = tid; lv < nvec; lv += nthr)
                cp_async16(tilev + swz_vec(uint32_t(lv)), gvec + base_vec + s_runoff[lv >> rv] + (lv & rvmask));
        }
        // per-op tile constants: outer control predicate and outer part of the table index
        for (int m = tid; m < nouter; m += nthr) s_outer[m] = outer_value(outers, m, base_amp);
        // per-tile phase factors: product of the outer-indexed look-ups of a phase group
        for (int m = tid; m < nH; m += nthr) {
            const uint4 *he = hents + 4 * m;
            const int ntab = int(he[0].x);
            Cx<T> acc;
            acc.re = T(1); acc.im = T(0);
            for (int t = 0; t < ntab; t++) {
                const uint4 u = he[1 + (t >> 1)];
                const uint32_t table = (t & 1) ? u.z : u.x, osl = (t & 1) ? u.w : u.y;
                const int32_t v = outer_value(outers, int(osl), base_amp);
                if (v < 0) continue;
                const Cx<T> z = ldg_cx(tables + table + v);
                const T nr = fma(acc.re, z.re, -(acc.im * z.im));
                acc.im = fma(acc.re, z.im, acc.im * z.re);
                acc.re = nr;
            }
            *reinterpret_cast<Cx<T> *>(s_H + m) = acc;
        }
        // per-tile factors of the fused diagonals, stage 1: ONE table look-up per thread (all the
        // slices of all fused diagonals in parallel: the latency of a single global load)
        for (int m = tid; m < pg.nFS; m += nthr) {
            const uint4 u = fents[nF + m];
            const int32_t v = outer_value(outers, int(u.z), base_amp);
            Cx<T> z;
            z.re = T(1); z.im = T(0);
            if (v >= 0) z = ldg_cx(tables + u.y + v);
            *reinterpret_cast<Cx<T> *>(s_FS + m) = z;
        }
        cp_async_wait_all();
        __syncthreads();
        // stage 2: per element, the product of the slices that cover it (two independent chains)
        if (nF) {
            for (int m = tid; m < nF * N; m += nthr) {
                const int f = m / N, e = m % N;
                const uint4 dir = fents[f];
                Cx<T> acc0, acc1;
                acc0.re = T(1); acc0.im = T(0); acc1 = acc0;
                for (uint32_t sl = 0; sl < dir.y; sl += 2) {
                    if ((fents[dir.x + sl].x >> e) & 1u)
                        acc0 = cx_mul<T>(acc0, *reinterpret_cast<const Cx<T> *>(s_FS + (dir.x - nF + sl)));
                    if (sl + 1 < dir.y && ((fents[dir.x + sl + 1].x >> e) & 1u))
                        acc1 = cx_mul<T>(acc1, *reinterpret_cast<const Cx<T> *>(s_FS + (dir.x - nF + sl + 1)));
                }
                s_F[m] = cx_mul<T>(acc0, acc1);
            }
            __syncthreads();
        }

        // ---- rounds
#pragma unroll 1
        for (int rd = 0; rd < nrounds; rd++) {
            const uint4 r0 = pp.u[rounds + 3 * rd];
            {
                uint32_t vd[4] = {r0.z & 0xffffu, r0.z >> 16, r0.w & 0xffffu, r0.w >> 16};
                const uint2 rc = s_rconst[rd * nthr + tid];
                uint32_t S = rc.x;
                const uint32_t base = rc.y;
                Cx<T> x[N];
#pragma unroll
                for (int v = 0; v < 16; v++) {
                    const uint32_t off = S ^ ((v & 1) ? vd[0] : 0u) ^ ((v & 2) ? vd[1] : 0u) ^
                                         ((v & 4) ? vd[2] : 0u) ^ ((v & 8) ? vd[3] : 0u);
                    if constexpr (sizeof(T) == 8) {
                        const double2 q = *reinterpret_cast<const double2 *>(smem_raw + off);
                        x[v].re = q.x; x[v].im = q.y;
                    } else {
                        const float4 q = *reinterpret_cast<const float4 *>(smem_raw + off);
                        x[2 * v].re = q.x; x[2 * v].im = q.y;
                        x[2 * v + 1].re = q.z; x[2 * v + 1].im = q.w;
                    }
                }

                // The op loop has warp-uniform control flow only: a thread whose predicate fails
                // (control bit outside the registers is 0, outer control not satisfied) runs the
                // same op with an empty element mask / a unit phase instead of branching around it.
                int op = int(r0.x);                      // unit index into the program image
                const int op_end = op + int(r0.y);       // r0.y = units of this round's op stream
#pragma unroll 1
                while (op != op_end) {
                    uint4 h0 = pp.u[op], h1 = pp.u[op + 1];
                    int pay = op + 2;
                    op += int(h0.x >> 16);
                    const uint32_t code = h0.x & 0xffffu;
                    if (code >= uint32_t(C_DIAGF)) {
                        // the fused diagonals and butterflies: no predicates, their own short compare tree
                        switch (code) {
                        case C_DIAGF: {   // h0.y = F index (0xffff: none) | has_g << 16, h0.w = unit mask, h1.x = G
                            const uint32_t fidx = h0.y & 0xffffu;
                            const Cx<T> *const gp = tables + h1.x + tid * (VS ? 32 : 16);
                            const bool full = h0.w == 0xffffu;
                            if (fidx != 0xffffu) {
                                if (full) op_diagf<T, true, true>(x, gp, h0.w, (h0.y >> 16) != 0u, s_F + fidx * N);
                                else op_diagf<T, true, false>(x, gp, h0.w, (h0.y >> 16) != 0u, s_F + fidx * N);
                            } else {
                                if (full) op_diagf<T, false, true>(x, gp, h0.w, true, nullptr);
                                else op_diagf<T, false, false>(x, gp, h0.w, true, nullptr);
                            }
                        } break;
                        case C_DIAGC: op_diagc<T>(x, pp, pay, h0.w); break;
                        case C_GROUP1H: {
                            const uint32_t slots = h0.y >> 16;
                            if (slots & 1u) op_hadamard<T, 0>(x);
                            if (slots & 2u) op_hadamard<T, 1>(x);
                            if (slots & 4u) op_hadamard<T, 2>(x);
                            if (slots & 8u) op_hadamard<T, 3>(x);
                            if (slots & 16u) op_hadamard<T, 4>(x);
                        } break;
#define QJ_DCS(A) op_diagc_slot<T, A>(x, pp, pay)
                        QJ_SLOT_CASES(C_DIAGCS, QJ_DCS)
#undef QJ_DCS
                        case C_DIAGS: {   // h0.y = F index (0xffff: none) | has_g << 16, h0.w = slot mask, h1.x = G
                            const uint32_t fidx = h0.y & 0xffffu;
                            const Cx<T> *const gp = tables + h1.x + tid * (VS ? 8 : 4);
                            if (fidx != 0xffffu) op_diags<T, true>(x, gp, h0.w, (h0.y >> 16) != 0u, s_F + fidx * N);
                            else op_diags<T, false>(x, gp, h0.w, true, nullptr);
                        } break;
                        }
                        continue;
                    }
                    uint32_t emask = h0.w;
                    int oi = 0;
                    if (code != C_PHASE && (h0.z != 0u || (h0.y & 0xffffu) != 0xffffu)) {   // predicated op
                        const uint32_t oslot = h0.y & 0xffffu, tmask = h0.z;
                        bool ok = (base & tmask) == tmask;
                        if (oslot != 0xffffu) {
                            oi = s_outer[oslot];
                            ok = ok && oi >= 0;
                            oi = max(oi, 0);
                        }
                        emask = ok ? emask : 0u;
                    }
                    switch (code) {
#define QJ_P1(A) op_perm1<T, A>(x, emask)
#define QJ_D2(A, B) op_dense2<T, A, B>(x, pp, pay, emask)
#define QJ_P2(A, B) op_perm2<T, A, B>(x, emask)
#define QJ_D2R(A, B) op_dense2r<T, A, B>(x, pp, pay, emask)
                        case C_GROUP1C: op_group1<T, 0>(x, pp, pay, h0.y >> 16, emask); break;
                        case C_GROUP1R: op_group1<T, 1>(x, pp, pay, h0.y >> 16, emask); break;
                        case C_GROUP1X: op_group1<T, 2>(x, pp, pay, h0.y >> 16, emask); break;
                        QJ_SLOT_CASES(C_PERM1, QJ_P1)
                        QJ_PAIR_CASES(C_DENSE2, QJ_D2)
                        QJ_PAIR_CASES(C_PERM2, QJ_P2)
                        QJ_PAIR_CASES(C_DENSE2R, QJ_D2R)
#undef QJ_D2R
#undef QJ_P1
#undef QJ_D2
#undef QJ_P2
                        case C_PHASE: {
                          // A run of consecutive phase groups is handled here without going back to the
                          // dispatcher, and the per-thread factor of the NEXT group is loaded before the
                          // current one is applied (its latency hides behind the multiplies).
                          Cx<T> gcur;
                          gcur.re = T(1); gcur.im = T(0);
                          if (h1.x != 0xffffffffu) gcur = ldg_cx(tables + h1.x + tid);
                          for (;;) {
                            const bool more = op != op_end && (pp.u[op].x & 0xffffu) == uint32_t(C_PHASE);
                            uint4 n0 = h0, n1 = h1;
                            Cx<T> gnext;
                            gnext.re = T(1); gnext.im = T(0);
                            if (more) {
                                n0 = pp.u[op]; n1 = pp.u[op + 1];
                                if (n1.x != 0xffffffffu) gnext = ldg_cx(tables + n1.x + tid);
                            }
                            // h0.y = ntab | sel << 16, h0.z = all-sign flag; descriptors: 2 (<= 5 fields)
                            // or 3 units: {table, nf | oslot << 16, tmask, f0} {f1..f4} {f5..f8}
                            const int ntab = int(h0.y & 0xffffu);
                            const uint32_t sel = h0.y >> 16;
                            const bool allsign = (h0.z & 1u) != 0u;
                            int d = pay;
                            Cx<T> ph;
                            ph.re = T(1); ph.im = T(0);
                            uint32_t sg = 0;
                            bool have = false;
                            constexpr int KB = 4;   // look-ups in flight
                            // h1.x: per-thread factor G[tid] (thread-only part, precomputed on the host:
                            // a thread's tile position is the same in every tile); h1.y: per-tile factor
                            if (h1.x != 0xffffffffu) {
                                if (allsign) sg ^= sign_of(gcur.re);
                                else ph = gcur;
                                have = true;
                            }
                            if (h0.z & 2u) sg ^= h1.w ^ (uint32_t(__popc(base & h1.z) & 1) << 31);   // parity sign: no table
                            if (h1.y != 0xffffffffu) {
                                const Cx<T> z = *reinterpret_cast<const Cx<T> *>(s_H + h1.y);
                                if (allsign) {
                                    sg ^= sign_of(z.re);
                                } else if (!have) {
                                    ph = z;
                                } else {
                                    const T nr = fma(ph.re, z.re, -(ph.im * z.im));
                                    ph.im = fma(ph.re, z.im, ph.im * z.re);
                                    ph.re = nr;
                                }
                                have = true;
                            }
#pragma unroll 1
                            for (int t0 = 0; t0 < ntab; t0 += KB) {
                                Cx<T> z[KB];
#pragma unroll
                                for (int k = 0; k < KB; k++) {
                                    z[k].re = T(1); z[k].im = T(0);
                                    if (t0 + k < ntab) {
                                        const uint4 d0 = pp.u[d], d1 = pp.u[d + 1];
                                        const int nf = int(d0.y & 0xffffu);
                                        const uint32_t osl = d0.y >> 16;
                                        const int dx = d + 2;
                                        d += (nf > 5) ? 3 : 2;
                                        int idx = 0;
                                        bool ok = (base & d0.z) == d0.z;     // tile-local control outside the registers
                                        if (osl != 0xffffu) {
                                            idx = s_outer[osl];
                                            ok = ok && idx >= 0;             // outer control
                                        }
                                        {
                                            if (nf > 0) idx |= field_of(base, d0.w);
                                            if (nf > 1) {
                                                idx |= field_of(base, d1.x);
                                                if (nf > 2) idx |= field_of(base, d1.y);
                                                if (nf > 3) idx |= field_of(base, d1.z);
                                                if (nf > 4) idx |= field_of(base, d1.w);
                                                if (nf > 5) {
                                                    const uint4 d2 = pp.u[dx];
                                                    idx |= field_of(base, d2.x);
                                                    if (nf > 6) idx |= field_of(base, d2.y);
                                                    if (nf > 7) idx |= field_of(base, d2.z);
                                                    if (nf > 8) idx |= field_of(base, d2.w);
                                                }
                                            }
                                            if (ok) z[k] = ldg_cx(tables + d0.x + idx);
                                        }
                                    }
                                }
#pragma unroll
                                for (int k = 0; k < KB; k++) {
                                    if (t0 + k >= ntab) break;
                                    if (allsign) {
                                        sg ^= sign_of(z[k].re);
                                    } else if (!have) {
                                        ph = z[k];
                                        have = true;
                                    } else {
                                        const T nr = fma(ph.re, z[k].re, -(ph.im * z[k].im));
                                        ph.im = fma(ph.re, z[k].im, ph.im * z[k].re);
                                        ph.re = nr;
                                    }
                                }
                            }
                            if (allsign) phase_apply<T, true>(x, sel, emask, T(0), T(0), sg);
                            else phase_apply<T, false>(x, sel, emask, ph.re, ph.im, 0u);
                            if (!more) break;
                            h0 = n0; h1 = n1;
                            pay = op + 2;
                            op += int(n0.x >> 16);
                            emask = n0.w;
                            gcur = gnext;
                          }
                        } break;
                        default: {  // C_DIAGN: table index = outer part | fields of the base | element part
                            const int nf = int(h0.y >> 16);
                            const uint4 f = pp.u[pay], w = pp.u[pay + 1];
                            int idxb = oi;
                            if (nf > 0) idxb |= field_of(base, h1.y);
                            if (nf > 1) idxb |= field_of(base, h1.z);
                            if (nf > 2) idxb |= field_of(base, h1.w);
                            if (nf > 3) idxb |= field_of(base, f.x);
                            if (nf > 4) idxb |= field_of(base, f.y);
                            if (nf > 5) idxb |= field_of(base, f.z);
                            if (nf > 6) idxb |= field_of(base, f.w);
                            const int w0 = int(w.x & 0xffffu), w1 = int(w.x >> 16), w2 = int(w.y & 0xffffu),
                                      w3 = int(w.y >> 16), w4 = int(w.z & 0xffffu);
                            const Cx<T> *tab = tables + h1.x;
#pragma unroll
                            for (int e8 = 0; e8 < N; e8 += 8) {  // eight gathers in flight before the first multiply
                                Cx<T> ph[8];
#pragma unroll
                                for (int k = 0; k < 8; k++) {
                                    const int e = e8 + k;
                                    const int idx = idxb | ((e & 1) ? w0 : 0) | ((e & 2) ? w1 : 0) | ((e & 4) ? w2 : 0) |
                                                    ((e & 8) ? w3 : 0) | ((e & 16) ? w4 : 0);
                                    ph[k].re = T(1); ph[k].im = T(0);
                                    if ((emask >> e) & 1u) ph[k] = ldg_cx(tab + idx);
                                }
#pragma unroll
                                for (int k = 0; k < 8; k++) cmul_inplace<T>(x[e8 + k], ph[k].re, ph[k].im);
                            }
                        }
                    }
                }

                // the 16 scatter addresses are recomputed, not kept live across the op loop
                asm volatile("" : "+r"(S), "+r"(vd[0]), "+r"(vd[1]), "+r"(vd[2]), "+r"(vd[3]));
#pragma unroll
                for (int v = 0; v < 16; v++) {
                    const uint32_t off = S ^ ((v & 1) ? vd[0] : 0u) ^ ((v & 2) ? vd[1] : 0u) ^
                                         ((v & 4) ? vd[2] : 0u) ^ ((v & 8) ? vd[3] : 0u);
                    if constexpr (sizeof(T) == 8) {
                        *reinterpret_cast<double2 *>(smem_raw + off) = make_double2(x[v].re, x[v].im);
                    } else {
                        *reinterpret_cast<float4 *>(smem_raw + off) =
                            make_float4(x[2 * v].re, x[2 * v].im, x[2 * v + 1].re, x[2 * v + 1].im);
                    }
                }
            }
            __syncthreads();
        }

        // ---- store the tile (four vectors in flight per thread)
        if (fast_io) {
            uint4 *const gdst = gvec + pg.out_off + base_vec + io_thr;
#pragma unroll
            for (int u0 = 0; u0 < 16; u0 += 4) {
                uint4 q[4];
#pragma unroll
                for (int k = 0; k < 4; k++) q[k] = tilev[sw_t ^ pg.io_soff[u0 + k]];
#pragma unroll
                for (int k = 0; k < 4; k++) gdst[pg.io_goff[u0 + k]] = q[k];
            }
        } else {
#pragma unroll 1
            for (int v0 = tid; v0 < nvec; v0 += 4 * nthr) {
                uint4 q[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int lv = v0 + u * nthr;
                    if (lv < nvec) q[u] = tilev[swz_vec(uint32_t(lv))];
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int lv = v0 + u * nthr;
                    if (lv < nvec) gvec[pg.out_off + base_vec + s_runoff[lv >> rv] + (lv & rvmask)] = q[u];
                }
            }
        }
        __syncthreads();   // the next tile's asynchronous copies overwrite the buffer
    }
}

}  // namespace

// complex64 launch, defined in pass_kernels_f32.cu (`geom` / `pp`: PassGeom / ProgParam of this header)
int launch_k_pass_f32(int device, unsigned grid, int threads, size_t smem, cudaStream_t stream, void *state,
                      const void *geom, const void *tables, const void *pp);
}  // namespace qj
