// ops_kernels.cu -- state preparation, collapse, probabilities, sampling and shard exchange.
//
// Replaces /root/reference/src/qibojit/custom_operators/ops.py (numba) and the collapse /
// initial-state RawKernels (raw_kernels.py:521-558).  All reductions use warp shuffles and a
// fixed two-stage tree (per-block partials in the handle's scratch, then one block), so the
// norm and the marginals are reproducible from run to run -- the reference's prange
// reduction (ops.py:67-75) is not.

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace qj {
namespace {

constexpr int kThreads = 256;

inline unsigned persistent_grid(const qj_handle *h, int64_t work_items, int per_block) {
    int64_t need = (work_items + per_block - 1) / per_block;
    int64_t cap = int64_t(h->sm_count) * 8;
    return (unsigned)std::max<int64_t>(1, std::min(need, cap));
}

// ------------------------------------------------------------------ initial state
// ops.py:14-18: state[0] = 1, everything else 0.  16-byte stores, grid-stride.
__global__ void __launch_bounds__(kThreads) k_initial_state(float4 *st, int64_t nvec, int is_c128) {
    const int64_t stride = int64_t(gridDim.x) * kThreads;
    for (int64_t i = int64_t(blockIdx.x) * kThreads + threadIdx.x; i < nvec; i += stride) {
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i == 0) {
            if (is_c128) {
                double2 one = make_double2(1.0, 0.0);
                z = *reinterpret_cast<float4 *>(&one);
            } else {
                z.x = 1.f;
            }
        }
        st[i] = z;
    }
}

// ------------------------------------------------------------------ collapse
// pass 1 (ops.py:47-56 / 63-75): zero every amplitude whose measured bits differ from the
// outcome, accumulate |amp|^2 of the survivors.  One thread per amplitude, contiguous, so the
// zeroing stores and the survivor loads are both coalesced; doomed amplitudes are never read.
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_collapse_zero(Cx<T> *__restrict__ st, int64_t n, int64_t mask, int64_t want, int do_norm,
                double *__restrict__ partials) {
    const int64_t stride = int64_t(gridDim.x) * kThreads;
    double acc = 0.0;
    for (int64_t i = int64_t(blockIdx.x) * kThreads + threadIdx.x; i < n; i += stride) {
        if ((i & mask) != want) {
            Cx<T> z; z.re = T(0); z.im = T(0);
            st[i] = z;
        } else if (do_norm) {
            const Cx<T> a = st[i];
            acc += double(a.re) * double(a.re) + double(a.im) * double(a.im);
        }
    }
    if (do_norm) {
        const double s = block_sum<kThreads>(acc);
        if (threadIdx.x == 0) partials[blockIdx.x] = s;
    }
}

// fixed-order sum of the per-block partials -> out[0] = sum, out[1] = sqrt(sum)
__global__ void __launch_bounds__(kThreads) k_finish_sum(const double *partials, int n, double *out) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += kThreads) acc += partials[i];
    const double s = block_sum<kThreads>(acc);
    if (threadIdx.x == 0) { out[0] = s; out[1] = sqrt(s); }
}

// pass 2 (ops.py:76-79): divide the survivors by the norm (read from device memory: no
// host round trip between the passes)
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_collapse_scale(Cx<T> *__restrict__ st, const __grid_constant__ GateGeom geo, const double *norm) {
    const int64_t ngroups = int64_t(1) << geo.nfree;
    const double nrm = norm[1];
    const int64_t stride = int64_t(gridDim.x) * kThreads;
    for (int64_t g = int64_t(blockIdx.x) * kThreads + threadIdx.x; g < ngroups; g += stride) {
        const int64_t i = expand_index(g, geo) | geo.cmask;
        Cx<T> a = st[i];
        a.re = T(double(a.re) / nrm);
        a.im = T(double(a.im) / nrm);
        st[i] = a;
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_norm2(const Cx<T> *__restrict__ st, int64_t n, double *__restrict__ partials) {
    const int64_t stride = int64_t(gridDim.x) * kThreads;
    double acc = 0.0;
    for (int64_t i = int64_t(blockIdx.x) * kThreads + threadIdx.x; i < n; i += stride) {
        const Cx<T> a = st[i];
        acc += double(a.re) * double(a.re) + double(a.im) * double(a.im);
    }
    const double s = block_sum<kThreads>(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// max_i |st[i] - ref|: the distance of the state from a constant vector (the closed form of
// QFT|0..0>, 2^(-n/2) everywhere) without a state-sized temporary; warp-shuffle + block tree
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_max_deviation(const Cx<T> *__restrict__ st, int64_t n, double rr, double ri, double *__restrict__ partials) {
    const int64_t stride = int64_t(gridDim.x) * kThreads;
    double acc = 0.0;
    for (int64_t i = int64_t(blockIdx.x) * kThreads + threadIdx.x; i < n; i += stride) {
        const Cx<T> a = st[i];
        const double dr = double(a.re) - rr, di = double(a.im) - ri;
        acc = fmax(acc, dr * dr + di * di);
    }
    __shared__ double warp_max[kThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc = fmax(acc, __shfl_xor_sync(0xffffffffu, acc, o));
    if ((threadIdx.x & 31) == 0) warp_max[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int w = 0; w < kThreads / 32; w++) m = fmax(m, warp_max[w]);
        partials[blockIdx.x] = m;
    }
}
__global__ void __launch_bounds__(kThreads) k_finish_max(const double *partials, int n, double *out) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += kThreads) acc = fmax(acc, partials[i]);
    __shared__ double warp_max[kThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc = fmax(acc, __shfl_xor_sync(0xffffffffu, acc, o));
    if ((threadIdx.x & 31) == 0) warp_max[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int w = 0; w < kThreads / 32; w++) m = fmax(m, warp_max[w]);
        out[0] = sqrt(m);
    }
}

// ------------------------------------------------------------------ probabilities
// (a) every qubit measured in natural order: |psi|^2 elementwise
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_probs_full(const Cx<T> *__restrict__ st, int64_t n, T *__restrict__ probs) {
    // 16-byte accesses on both sides: a thread turns 32 (complex128) / 32 (complex64) bytes of
    // amplitudes into 16 bytes of probabilities; two independent groups per iteration
    constexpr int V = 16 / sizeof(T);              // probabilities per 16-byte store: 2 (double) / 4 (float)
    const int64_t ngroups = n / V;
    const int64_t stride = int64_t(gridDim.x) * kThreads;
    const float4 *src = reinterpret_cast<const float4 *>(st);
    float4 *dst = reinterpret_cast<float4 *>(probs);
    for (int64_t g = int64_t(blockIdx.x) * kThreads + threadIdx.x; g < ngroups; g += stride) {
        if constexpr (sizeof(T) == 8) {
            const float4 q0 = __ldcs(src + 2 * g), q1 = __ldcs(src + 2 * g + 1);
            const double2 a = *reinterpret_cast<const double2 *>(&q0), b = *reinterpret_cast<const double2 *>(&q1);
            double2 o = make_double2(a.x * a.x + a.y * a.y, b.x * b.x + b.y * b.y);
            __stcs(dst + g, *reinterpret_cast<float4 *>(&o));
        } else {
            const float4 a = __ldcs(src + 2 * g), b = __ldcs(src + 2 * g + 1);
            __stcs(dst + g, make_float4(a.x * a.x + a.y * a.y, a.z * a.z + a.w * a.w, b.x * b.x + b.y * b.y, b.z * b.z + b.w * b.w));
        }
    }
    // (registers of fewer than V amplitudes: one or two qubits)
    for (int64_t i = ngroups * V + int64_t(blockIdx.x) * kThreads + threadIdx.x; i < n; i += stride) {
        const Cx<T> a = st[i];
        probs[i] = a.re * a.re + a.im * a.im;
    }
}

struct ProbGeom {
    int nmeas;
    int bits[QJ_MAX_QUBITS];  // index bit of output bit (nmeas-1-j)
};

__device__ __forceinline__ int64_t gather_bits(int64_t i, const ProbGeom &pg) {
    int64_t o = 0;
#pragma unroll 1
    for (int j = 0; j < pg.nmeas; j++) o |= ((i >> pg.bits[j]) & 1) << (pg.nmeas - 1 - j);
    return o;
}

// (b) few measured qubits (2^t bins fit in shared memory): coalesced streaming read, per-block
// shared-memory histogram (warp-aggregated when the whole warp hits one bin), per-block bins to
// scratch, fixed-order finish.
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_probs_hist(const Cx<T> *__restrict__ st, int64_t n, const __grid_constant__ ProbGeom pg,
             double *__restrict__ block_bins) {
    extern __shared__ double s_bins[];
    const int nb = 1 << pg.nmeas;
    for (int b = threadIdx.x; b < nb; b += kThreads) s_bins[b] = 0.0;
    __syncthreads();
    const int64_t stride = int64_t(gridDim.x) * kThreads;
    const int64_t start = int64_t(blockIdx.x) * kThreads + threadIdx.x;
    // all threads of a warp iterate the same number of times (n is a power of two >= 32 or the
    // guard below masks the tail), so the full-mask shuffles are safe
    for (int64_t i0 = int64_t(blockIdx.x) * kThreads; i0 < n; i0 += stride) {
        const int64_t i = i0 + threadIdx.x;
        double p = 0.0;
        int64_t bin = 0;
        if (i < n) {
            const Cx<T> a = st[i];
            p = double(a.re) * double(a.re) + double(a.im) * double(a.im);
            bin = gather_bits(i, pg);
        }
        const int64_t bin0 = __shfl_sync(0xffffffffu, bin, 0);
        if (__all_sync(0xffffffffu, bin == bin0 || i >= n)) {
            const double s = warp_sum(p);
            if ((threadIdx.x & 31) == 0) atomicAdd(&s_bins[bin0], s);
        } else if (i < n) {
            atomicAdd(&s_bins[bin], p);
        }
    }
    (void)start;
    __syncthreads();
    for (int b = threadIdx.x; b < nb; b += kThreads) block_bins[size_t(blockIdx.x) * nb + b] = s_bins[b];
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_probs_hist_finish(const double *__restrict__ block_bins, int nblocks, int nb, T *__restrict__ probs) {
    for (int b = blockIdx.x * kThreads + threadIdx.x; b < nb; b += gridDim.x * kThreads) {
        double acc = 0.0;
        for (int k = 0; k < nblocks; k++) acc += block_bins[size_t(k) * nb + b];
        probs[b] = T(acc);
    }
}

// (c) many measured qubits: one thread per output bin sums its 2^u unmeasured amplitudes
struct ScatterGeom {
    int nmeas, nun;
    int mbits[QJ_MAX_QUBITS];  // index bit of output bit j (LSB first)
    int ubits[QJ_MAX_QUBITS];  // unmeasured index bits
};
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_probs_gather(const Cx<T> *__restrict__ st, const __grid_constant__ ScatterGeom sg,
               T *__restrict__ probs) {
    const int64_t nout = int64_t(1) << sg.nmeas;
    const int64_t nun = int64_t(1) << sg.nun;
    const int64_t stride = int64_t(gridDim.x) * kThreads;
    for (int64_t o = int64_t(blockIdx.x) * kThreads + threadIdx.x; o < nout; o += stride) {
        int64_t base = 0;
        for (int j = 0; j < sg.nmeas; j++) base |= ((o >> j) & 1) << sg.mbits[j];
        double acc = 0.0;
        for (int64_t r = 0; r < nun; r++) {
            int64_t a = base;
            for (int j = 0; j < sg.nun; j++) a |= ((r >> j) & 1) << sg.ubits[j];
            const Cx<T> v = st[a];
            acc += double(v.re) * double(v.re) + double(v.im) * double(v.im);
        }
        probs[o] = T(acc);
    }
}

// ------------------------------------------------------------------ Metropolis sampler
// ops.py:86-108 with numba's MT19937 streams (numba/_random.c:37-75,
// numba/cpython/randomimpl.py:109-196, 454-523).  One warp per chain: the 624-word state lives
// in shared memory, the twist is done cooperatively, every lane replays the (cheap) chain
// logic redundantly so control flow stays uniform; lane 0 commits the counts.
struct WarpMT {
    uint32_t *mt;
    int idx;
};

__device__ __forceinline__ void mt_twist_warp(uint32_t *mt) {
    const int lane = threadIdx.x & 31;
    // sequential dependencies: new mt[i] needs old mt[i], old mt[i+1] (new mt[0] for i = 623)
    // and mt[(i+397)%624], which is old for i < 227 and new otherwise -> chunks of 32 in order.
    for (int s = 0; s < 623; s += 32) {
        const int i = s + lane;
        uint32_t v = 0;
        const bool on = i < 623;
        if (on) {
            const uint32_t y = (mt[i] & 0x80000000u) | (mt[i + 1] & 0x7fffffffu);
            const int src = (i < 227) ? i + 397 : i - 227;
            v = mt[src] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        __syncwarp();
        if (on) mt[i] = v;
        __syncwarp();
    }
    if (lane == 0) {
        const uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    __syncwarp();
}

__device__ __forceinline__ uint32_t mt_next(WarpMT &s) {
    if (s.idx >= 624) { mt_twist_warp(s.mt); s.idx = 0; }
    uint32_t y = s.mt[s.idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

__device__ __forceinline__ int64_t mt_randint(WarpMT &s, int64_t n) {
    if (n == 1) return 0;
    const int nbits = 64 - __clzll((unsigned long long)(n - 1));
    for (;;) {
        int64_t r;
        if (nbits <= 32) {
            const uint32_t mask = 0xffffffffu >> (32 - nbits);
            r = int64_t(mt_next(s) & mask);
        } else {
            const uint32_t mask = 0xffffffffu >> (64 - nbits);
            const uint64_t high = mt_next(s) & mask;
            const uint64_t low = mt_next(s);
            r = int64_t(low + (high << 32));
        }
        if (r < n) return r;
    }
}

template <typename T>
__global__ void __launch_bounds__(32)
k_metropolis(unsigned long long *__restrict__ freq, const T *__restrict__ probs, int64_t nstates,
             const int64_t *__restrict__ chain_seed, const int64_t *__restrict__ chain_shots,
             const int64_t *__restrict__ start) {
    __shared__ uint32_t s_mt[624];
    const int lane = threadIdx.x;
    const uint32_t seed = (uint32_t)chain_seed[blockIdx.x];
    if (lane == 0) {
        s_mt[0] = seed;
        for (int i = 1; i < 624; i++)
            s_mt[i] = 1812433253u * (s_mt[i - 1] ^ (s_mt[i - 1] >> 30)) + (uint32_t)i;
    }
    __syncwarp();
    WarpMT rng{s_mt, 624};
    int64_t shot = start[0];
    T pshot = probs[shot];
    const int64_t nshots = chain_shots[blockIdx.x];
    for (int64_t it = 0; it < nshots; it++) {
        const int64_t r = mt_randint(rng, nstates);
        const int64_t new_shot = (shot + r) % nstates;
        const T pnew = probs[new_shot];
        const uint32_t a = mt_next(rng) >> 5, b = mt_next(rng) >> 6;
        const double u = (double(b) + double(a) * 67108864.0) / 9007199254740992.0;
        const T ratio = pnew / pshot;  // IEEE division in the probability dtype, as numba does
        if (double(ratio) > u) { shot = new_shot; pshot = pnew; }
        if (lane == 0) atomicAdd(freq + shot, 1ull);
    }
}

// first index of the maximum (np.argmax), two-stage, deterministic
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_argmax_partial(const T *__restrict__ p, int64_t n, double *__restrict__ pv, int64_t *__restrict__ pi) {
    __shared__ double sv[kThreads];
    __shared__ int64_t si[kThreads];
    double best = -1.0;
    int64_t bi = INT64_MAX;
    const int64_t stride = int64_t(gridDim.x) * kThreads;
    for (int64_t i = int64_t(blockIdx.x) * kThreads + threadIdx.x; i < n; i += stride) {
        const double v = double(p[i]);
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
    sv[threadIdx.x] = best; si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = kThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double v = sv[threadIdx.x + o];
            const int64_t j = si[threadIdx.x + o];
            if (v > sv[threadIdx.x] || (v == sv[threadIdx.x] && j < si[threadIdx.x])) {
                sv[threadIdx.x] = v; si[threadIdx.x] = j;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { pv[blockIdx.x] = sv[0]; pi[blockIdx.x] = si[0]; }
}
__global__ void k_argmax_finish(const double *pv, const int64_t *pi, int n, int64_t *out) {
    double best = -1.0;
    int64_t bi = INT64_MAX;
    for (int i = 0; i < n; i++)
        if (pv[i] > best || (pv[i] == best && pi[i] < bi)) { best = pv[i]; bi = pi[i]; }
    out[0] = bi;
}

// ------------------------------------------------------------------ inverse-CDF sampling
// three-kernel inclusive scan of probs -> cdf (double), then searchsorted(side="right")
constexpr int kScanItems = 8;
constexpr int kScanTile = kThreads * kScanItems;

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_scan_tile_sums(const T *__restrict__ p, int64_t n, double *__restrict__ tile_sums) {
    const int64_t base = int64_t(blockIdx.x) * kScanTile + int64_t(threadIdx.x) * kScanItems;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) if (base + k < n) acc += double(p[base + k]);
    const double s = block_sum<kThreads>(acc);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = s;
}
// exclusive scan of tile sums by a single block (sequential over chunks of kThreads)
__global__ void __launch_bounds__(kThreads) k_scan_offsets(double *tile_sums, int64_t ntiles) {
    __shared__ double s[kThreads];
    __shared__ double carry;
    if (threadIdx.x == 0) carry = 0.0;
    __syncthreads();
    for (int64_t c0 = 0; c0 < ntiles; c0 += kThreads) {
        const int64_t i = c0 + threadIdx.x;
        const double v = (i < ntiles) ? tile_sums[i] : 0.0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < kThreads; o <<= 1) {
            const double t = (threadIdx.x >= o) ? s[threadIdx.x - o] : 0.0;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < ntiles) tile_sums[i] = carry + s[threadIdx.x] - v;  // exclusive
        __syncthreads();
        if (threadIdx.x == 0) carry += s[kThreads - 1];
        __syncthreads();
    }
}
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_scan_apply(const T *__restrict__ p, int64_t n, const double *__restrict__ tile_off,
             double *__restrict__ cdf) {
    __shared__ double s[kThreads];
    const int64_t base = int64_t(blockIdx.x) * kScanTile + int64_t(threadIdx.x) * kScanItems;
    double loc[kScanItems];
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        acc += (base + k < n) ? double(p[base + k]) : 0.0;
        loc[k] = acc;
    }
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 1; o < kThreads; o <<= 1) {
        const double t = (threadIdx.x >= o) ? s[threadIdx.x - o] : 0.0;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    const double off = tile_off[blockIdx.x] + s[threadIdx.x] - acc;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) if (base + k < n) cdf[base + k] = off + loc[k];
}
__global__ void __launch_bounds__(kThreads)
k_searchsorted(const double *__restrict__ cdf, int64_t n, const double *__restrict__ u, int64_t nshots,
               int64_t *__restrict__ shots) {
    const double total = cdf[n - 1];
    const int64_t stride = int64_t(gridDim.x) * kThreads;
    for (int64_t s = int64_t(blockIdx.x) * kThreads + threadIdx.x; s < nshots; s += stride) {
        const double x = u[s];
        int64_t lo = 0, hi = n;  // first index with cdf[idx]/total > x
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (cdf[mid] / total > x) hi = mid; else lo = mid + 1;
        }
        shots[s] = lo < n ? lo : n - 1;
    }
}

// ------------------------------------------------------------------ shard exchange
// ops.py:131-137.  In-place swap over a peer mapping: this rank handles the first or second
// half of the exchanged amplitudes so both NVLink directions and both GPUs are busy.
__global__ void __launch_bounds__(kThreads)
k_swap_peer(float4 *__restrict__ local, float4 *__restrict__ peer, int mv, int64_t local_bit,
            int64_t peer_bit, int64_t g_begin, int64_t g_end) {
    const int64_t stride = int64_t(gridDim.x) * kThreads;
    for (int64_t g = g_begin + int64_t(blockIdx.x) * kThreads + threadIdx.x; g < g_end; g += stride) {
        const int64_t lo = g & ((int64_t(1) << mv) - 1);
        const int64_t i = ((g >> mv) << (mv + 1)) | lo;
        const float4 a = local[i | local_bit];
        const float4 b = peer[i | peer_bit];
        local[i | local_bit] = b;
        peer[i | peer_bit] = a;
    }
}
__global__ void __launch_bounds__(kThreads)
k_swap_pack(const float4 *__restrict__ local, float4 *__restrict__ buf, int mv, int64_t bit,
            int64_t g_begin, int64_t count, int unpack) {
    const int64_t stride = int64_t(gridDim.x) * kThreads;
    for (int64_t k = int64_t(blockIdx.x) * kThreads + threadIdx.x; k < count; k += stride) {
        const int64_t g = g_begin + k;
        const int64_t lo = g & ((int64_t(1) << mv) - 1);
        const int64_t i = (((g >> mv) << (mv + 1)) | lo) | bit;
        if (unpack) const_cast<float4 *>(local)[i] = buf[k];
        else buf[k] = local[i];
    }
}

// sub-block of a shard selected by several index bits: group g -> vector index with zeros
// inserted at pos[] (ascending vector-level positions), OR the fixed bit values
struct SubBlock {
    int npos;
    int pos[QJ_MAX_GLOBAL_SWAP];
    int64_t ormask;
};
__global__ void __launch_bounds__(kThreads)
k_swap_pack_bits(const float4 *__restrict__ local, float4 *__restrict__ buf, const __grid_constant__ SubBlock sb,
                 int64_t g_begin, int64_t count, int unpack) {
    const int64_t stride = int64_t(gridDim.x) * kThreads;
    for (int64_t k = int64_t(blockIdx.x) * kThreads + threadIdx.x; k < count; k += stride) {
        int64_t i = g_begin + k;
#pragma unroll 1
        for (int j = 0; j < sb.npos; j++) {
            const int p = sb.pos[j];
            i = ((i >> p) << (p + 1)) | (i & ((int64_t(1) << p) - 1));
        }
        i |= sb.ormask;
        if (unpack) const_cast<float4 *>(local)[i] = buf[k];
        else buf[k] = local[i];
    }
}

// Multi-qubit exchange over peer memory (NVLink): this rank's sub-block `lmask` and the peer's
// sub-block `pmask` trade places, element by element, with no staging buffer, no pack / unpack
// pass and no library call in between: each of the two ranks of a pair runs this kernel on its
// half of the groups, so both link directions carry reads and (posted) writes at once.
// Four independent vector swaps per thread and iteration keep enough remote loads in flight.
__global__ void __launch_bounds__(kThreads)
k_swap_bits_peer(float4 *__restrict__ local, float4 *__restrict__ peer, const __grid_constant__ SubBlock sb,
                 int64_t lmask, int64_t pmask, int64_t g_begin, int64_t g_end) {
    // consecutive lanes take consecutive groups (whole 512-byte requests per warp, local and
    // remote); the U groups of a thread are a whole grid apart
    constexpr int U = 8;
    const int64_t step = int64_t(gridDim.x) * kThreads;
    for (int64_t g0 = g_begin + int64_t(blockIdx.x) * kThreads + threadIdx.x; g0 < g_end; g0 += step * U) {
        int64_t idx[U];
        float4 a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            int64_t i = g0 + u * step;
#pragma unroll 1
            for (int j = 0; j < sb.npos; j++) {
                const int p = sb.pos[j];
                i = ((i >> p) << (p + 1)) | (i & ((int64_t(1) << p) - 1));
            }
            idx[u] = i;
        }
#pragma unroll
        for (int u = 0; u < U; u++)
            if (g0 + u * step < g_end) b[u] = __ldcs(peer + (idx[u] | pmask));
#pragma unroll
        for (int u = 0; u < U; u++)
            if (g0 + u * step < g_end) a[u] = __ldcs(local + (idx[u] | lmask));
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (g0 + u * step < g_end) {
                __stcs(peer + (idx[u] | pmask), a[u]);
                __stcs(local + (idx[u] | lmask), b[u]);
            }
        }
    }
}


// Stream-ordered handshake with peers over mapped flag words (one 32-bit slot per writing rank in
// every rank's flag array).  Thread i publishes epochs[i] into its peer's slot for this rank, then
// waits until that peer's epoch shows up in this rank's own array.  Everything enqueued before the
// kernel on this stream is complete (stream order) and fenced before the flag is published, so a
// peer that has seen the epoch may read this rank's memory; a peer that never shows up trips the
// timeout and the launch fails instead of hanging the device.
struct PeerFlags {
    unsigned *remote_slot[64];
    int src_rank[64];
    unsigned epoch[64];
    int n;
};
__global__ void k_peer_handshake(volatile unsigned *local_flags, const __grid_constant__ PeerFlags pf,
                                 unsigned long long timeout_ns) {
    const int i = threadIdx.x;
    if (i >= pf.n) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pf.remote_slot[i]), "r"(pf.epoch[i]) : "memory");
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned seen;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(local_flags + pf.src_rank[i]) : "memory");
        if (int(seen - pf.epoch[i]) >= 0) break;
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) __trap();
        __nanosleep(200);
    }
    __threadfence_system();
}

template <typename F>
int launch_checked(qj_handle *h, F &&f) {
    f();
    h->launches++;
    QJ_CUDA_OK(cudaGetLastError());
    return QJ_OK;
}

}  // namespace
}  // namespace qj

using namespace qj;

extern "C" int qj_initial_state(qj_handle *h, void *state, int dtype, int nqubits) {
    qj::DeviceGuard device_guard(h);
    QJ_REQUIRE(h && state, "null handle or state");
    QJ_REQUIRE(nqubits >= 0 && nqubits <= QJ_MAX_QUBITS, "nqubits out of range");
    const int64_t bytes = (int64_t(1) << nqubits) * (dtype == QJ_C128 ? 16 : 8);
    if (bytes < 16) {  // single complex64 amplitude
        const float one[2] = {1.f, 0.f};
        QJ_CUDA_OK(cudaMemcpyAsync(state, one, 8, cudaMemcpyHostToDevice, h->stream));
        return QJ_OK;
    }
    const int64_t nvec = bytes / 16;
    return launch_checked(h, [&] {
        k_initial_state<<<persistent_grid(h, nvec, kThreads * 4), kThreads, 0, h->stream>>>(
            reinterpret_cast<float4 *>(state), nvec, dtype == QJ_C128);
    });
}

namespace {
template <typename T>
int collapse_t(qj_handle *h, void *state, int nqubits, const int32_t *qubits, int nt, int64_t result,
               int normalize) {
    const int64_t n = int64_t(1) << nqubits;
    int64_t mask = 0, want = 0;
    std::vector<int> pos;
    for (int j = 0; j < nt; j++) {
        QJ_REQUIRE(qubits[j] >= 0 && qubits[j] < nqubits, "measured qubit out of range");
        QJ_REQUIRE(!((mask >> qubits[j]) & 1), "duplicate measured qubit");
        mask |= int64_t(1) << qubits[j];
        want |= ((result >> j) & 1) << qubits[j];   // ops.py:34-40: bit j of result -> qubits[j]
        pos.push_back(qubits[j]);
    }
    QJ_REQUIRE(result >= 0 && result < (int64_t(1) << nt), "collapse outcome out of range");
    const unsigned grid = persistent_grid(h, n, kThreads * 4);
    QJ_REQUIRE(grid + 2 <= h->scratch_doubles, "scratch too small");
    int rc = launch_checked(h, [&] {
        k_collapse_zero<T><<<grid, kThreads, 0, h->stream>>>(reinterpret_cast<Cx<T> *>(state), n, mask,
                                                             want, normalize, h->scratch + 2);
    });
    if (rc || !normalize) return rc;
    rc = launch_checked(h, [&] {
        k_finish_sum<<<1, kThreads, 0, h->stream>>>(h->scratch + 2, (int)grid, h->scratch);
    });
    if (rc) return rc;
    GateGeom geo;
    memset(&geo, 0, sizeof(geo));
    std::sort(pos.begin(), pos.end());
    geo.npos = nt;
    for (int j = 0; j < nt; j++) geo.pos[j] = pos[j];
    geo.nfree = nqubits - nt;
    geo.cmask = want;
    return launch_checked(h, [&] {
        k_collapse_scale<T><<<persistent_grid(h, int64_t(1) << geo.nfree, kThreads * 4), kThreads, 0,
                              h->stream>>>(reinterpret_cast<Cx<T> *>(state), geo, h->scratch);
    });
}
}  // namespace

extern "C" int qj_collapse_state(qj_handle *h, void *state, int dtype, int nqubits,
                                 const int32_t *qubits, int ntargets, int64_t result, int normalize) {
    qj::DeviceGuard device_guard(h);
    QJ_REQUIRE(h && state && (qubits || ntargets == 0), "null argument");
    QJ_REQUIRE(ntargets >= 0 && ntargets <= nqubits && nqubits <= QJ_MAX_QUBITS, "bad qubit counts");
    if (dtype == QJ_C128) return collapse_t<double>(h, state, nqubits, qubits, ntargets, result, normalize);
    if (dtype == QJ_C64) return collapse_t<float>(h, state, nqubits, qubits, ntargets, result, normalize);
    return fail(QJ_ERR_INVALID, "unknown dtype");
}

extern "C" int qj_norm2(qj_handle *h, const void *state, int dtype, int nqubits, double *out) {
    qj::DeviceGuard device_guard(h);
    QJ_REQUIRE(h && state && out, "null argument");
    const int64_t n = int64_t(1) << nqubits;
    const unsigned grid = persistent_grid(h, n, kThreads * 4);
    int rc = launch_checked(h, [&] {
        if (dtype == QJ_C128)
            k_norm2<double><<<grid, kThreads, 0, h->stream>>>(reinterpret_cast<const Cx<double> *>(state), n, h->scratch + 2);
        else
            k_norm2<float><<<grid, kThreads, 0, h->stream>>>(reinterpret_cast<const Cx<float> *>(state), n, h->scratch + 2);
    });
    if (rc) return rc;
    rc = launch_checked(h, [&] { k_finish_sum<<<1, kThreads, 0, h->stream>>>(h->scratch + 2, (int)grid, h->scratch); });
    if (rc) return rc;
    QJ_CUDA_OK(cudaMemcpyAsync(out, h->scratch, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    QJ_CUDA_OK(cudaStreamSynchronize(h->stream));
    return QJ_OK;
}

extern "C" int qj_max_deviation(qj_handle *h, const void *state, int dtype, int nqubits, double ref_re,
                                double ref_im, double *out) {
    qj::DeviceGuard device_guard(h);
    QJ_REQUIRE(h && state && out, "null argument");
    QJ_REQUIRE(dtype == QJ_C64 || dtype == QJ_C128, "dtype must be QJ_C64 or QJ_C128");
    QJ_REQUIRE(nqubits >= 0 && nqubits <= QJ_MAX_QUBITS, "nqubits out of range");
    const int64_t n = int64_t(1) << nqubits;
    const unsigned grid = persistent_grid(h, n, kThreads * 4);
    int rc = launch_checked(h, [&] {
        if (dtype == QJ_C128)
            k_max_deviation<double><<<grid, kThreads, 0, h->stream>>>(reinterpret_cast<const Cx<double> *>(state), n,
                                                                      ref_re, ref_im, h->scratch + 2);
        else
            k_max_deviation<float><<<grid, kThreads, 0, h->stream>>>(reinterpret_cast<const Cx<float> *>(state), n,
                                                                     ref_re, ref_im, h->scratch + 2);
    });
    if (rc) return rc;
    rc = launch_checked(h, [&] { k_finish_max<<<1, kThreads, 0, h->stream>>>(h->scratch + 2, (int)grid, h->scratch); });
    if (rc) return rc;
    QJ_CUDA_OK(cudaMemcpyAsync(out, h->scratch, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    QJ_CUDA_OK(cudaStreamSynchronize(h->stream));
    return QJ_OK;
}

namespace {
template <typename T>
int probs_t(qj_handle *h, const void *state, int nqubits, const int32_t *bits, int nmeas, void *probs) {
    const int64_t n = int64_t(1) << nqubits;
    const Cx<T> *st = reinterpret_cast<const Cx<T> *>(state);
    T *out = reinterpret_cast<T *>(probs);
    int64_t seen = 0;
    bool natural = (nmeas == nqubits);
    for (int j = 0; j < nmeas; j++) {
        QJ_REQUIRE(bits[j] >= 0 && bits[j] < nqubits, "measured qubit out of range");
        QJ_REQUIRE(!((seen >> bits[j]) & 1), "duplicate measured qubit");
        seen |= int64_t(1) << bits[j];
        if (bits[j] != nqubits - 1 - j) natural = false;
    }
    if (natural) {
        QJ_REQUIRE((reinterpret_cast<uintptr_t>(probs) & 15) == 0 && (reinterpret_cast<uintptr_t>(state) & 15) == 0,
                   "state and probabilities must be 16-byte aligned");
        return launch_checked(h, [&] {
            k_probs_full<T><<<persistent_grid(h, n, kThreads * 4), kThreads, 0, h->stream>>>(st, n, out);
        });
    }
    if (nmeas <= 10) {
        ProbGeom pg;
        memset(&pg, 0, sizeof(pg));
        pg.nmeas = nmeas;
        for (int j = 0; j < nmeas; j++) pg.bits[j] = bits[j];
        const int nb = 1 << nmeas;
        unsigned grid = persistent_grid(h, n, kThreads * 8);
        const size_t cap = h->scratch_doubles / nb;
        if (grid > cap) grid = (unsigned)cap;
        QJ_REQUIRE(grid >= 1, "scratch too small for marginal");
        int rc = launch_checked(h, [&] {
            k_probs_hist<T><<<grid, kThreads, nb * sizeof(double), h->stream>>>(st, n, pg, h->scratch);
        });
        if (rc) return rc;
        return launch_checked(h, [&] {
            k_probs_hist_finish<T><<<(nb + kThreads - 1) / kThreads, kThreads, 0, h->stream>>>(h->scratch, (int)grid, nb, out);
        });
    }
    ScatterGeom sg;
    memset(&sg, 0, sizeof(sg));
    sg.nmeas = nmeas;
    for (int j = 0; j < nmeas; j++) sg.mbits[j] = bits[nmeas - 1 - j];
    for (int b = 0; b < nqubits; b++)
        if (!((seen >> b) & 1)) sg.ubits[sg.nun++] = b;
    return launch_checked(h, [&] {
        k_probs_gather<T><<<persistent_grid(h, int64_t(1) << nmeas, kThreads), kThreads, 0, h->stream>>>(st, sg, out);
    });
}
}  // namespace

extern "C" int qj_calculate_probabilities(qj_handle *h, const void *state, int dtype, int nqubits,
                                          const int32_t *bits, int nmeas, void *probs) {
    qj::DeviceGuard device_guard(h);
    QJ_REQUIRE(h && state && probs && (bits || nmeas == 0), "null argument");
    QJ_REQUIRE(nmeas >= 0 && nmeas <= nqubits && nqubits <= QJ_MAX_QUBITS, "bad qubit counts");
    if (dtype == QJ_C128) return probs_t<double>(h, state, nqubits, bits, nmeas, probs);
    if (dtype == QJ_C64) return probs_t<float>(h, state, nqubits, bits, nmeas, probs);
    return fail(QJ_ERR_INVALID, "unknown dtype");
}

namespace {
// host copy of the top-level generator that hands out chain seeds (ops.py:92-93)
struct HostMT {
    uint32_t mt[624];
    int idx;
    void seed(uint32_t s) {
        mt[0] = s;
        for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    uint32_t next() {
        if (idx >= 624) {
            for (int i = 0; i < 624; i++) {
                uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
                mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
        return y;
    }
    int64_t randint(int64_t n) {
        if (n == 1) return 0;
        int nbits = 64 - __builtin_clzll((uint64_t)(n - 1));
        for (;;) {
            int64_t r;
            if (nbits <= 32) r = int64_t(next() & (0xffffffffu >> (32 - nbits)));
            else {
                uint64_t high = next() & (0xffffffffu >> (64 - nbits));
                uint64_t low = next();
                r = int64_t(low + (high << 32));
            }
            if (r < n) return r;
        }
    }
};
}  // namespace

extern "C" int qj_measure_frequencies(qj_handle *h, int64_t *frequencies, const void *probs,
                                      int real_dtype, int64_t nshots, int nqubits, int64_t seed,
                                      int nthreads) {
    qj::DeviceGuard device_guard(h);
    QJ_REQUIRE(h && frequencies && probs, "null argument");
    QJ_REQUIRE(nthreads >= 1 && nthreads <= 4096, "nthreads out of range");
    QJ_REQUIRE(nshots >= 0 && nqubits >= 0 && nqubits <= QJ_MAX_QUBITS, "bad sampler arguments");
    const int64_t n = int64_t(1) << nqubits;
    // chain bookkeeping (a few integers) is prepared on the host exactly as ops.py:88-93
    std::vector<int64_t> meta(2 * size_t(nthreads));
    for (int t = 0; t < nthreads; t++) meta[nthreads + t] = nshots / nthreads;
    meta[2 * nthreads - 1] += nshots % nthreads;
    HostMT top;
    top.seed((uint32_t)seed);
    for (int t = 0; t < nthreads; t++) meta[t] = top.randint(100000000);
    const size_t need = 2 * size_t(nthreads) + 1 + 2 * 1024;
    QJ_REQUIRE(need <= h->scratch_doubles, "scratch too small for sampler");
    int64_t *d_meta = reinterpret_cast<int64_t *>(h->scratch);
    int64_t *d_start = d_meta + 2 * nthreads;
    double *d_pv = h->scratch + 2 * nthreads + 1;
    int64_t *d_pi = reinterpret_cast<int64_t *>(d_pv + 1024);
    QJ_CUDA_OK(cudaMemcpyAsync(d_meta, meta.data(), meta.size() * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
    unsigned grid = std::min<unsigned>(1024u, persistent_grid(h, n, kThreads * 4));
    int rc = launch_checked(h, [&] {
        if (real_dtype == QJ_C128)
            k_argmax_partial<double><<<grid, kThreads, 0, h->stream>>>(reinterpret_cast<const double *>(probs), n, d_pv, d_pi);
        else
            k_argmax_partial<float><<<grid, kThreads, 0, h->stream>>>(reinterpret_cast<const float *>(probs), n, d_pv, d_pi);
    });
    if (rc) return rc;
    rc = launch_checked(h, [&] { k_argmax_finish<<<1, 1, 0, h->stream>>>(d_pv, d_pi, (int)grid, d_start); });
    if (rc) return rc;
    rc = launch_checked(h, [&] {
        if (real_dtype == QJ_C128)
            k_metropolis<double><<<nthreads, 32, 0, h->stream>>>(reinterpret_cast<unsigned long long *>(frequencies), reinterpret_cast<const double *>(probs), n, d_meta, d_meta + nthreads, d_start);
        else
            k_metropolis<float><<<nthreads, 32, 0, h->stream>>>(reinterpret_cast<unsigned long long *>(frequencies), reinterpret_cast<const float *>(probs), n, d_meta, d_meta + nthreads, d_start);
    });
    if (rc) return rc;
    // meta lives in pageable host memory: make sure the copy has been consumed
    QJ_CUDA_OK(cudaStreamSynchronize(h->stream));
    return QJ_OK;
}

extern "C" int qj_sample_shots(qj_handle *h, const void *probs, int real_dtype, int nqubits,
                               const double *uniforms, int64_t nshots, int64_t *shots,
                               double *cdf_scratch) {
    qj::DeviceGuard device_guard(h);
    QJ_REQUIRE(h && probs && uniforms && shots && cdf_scratch, "null argument");
    const int64_t n = int64_t(1) << nqubits;
    const int64_t ntiles = (n + kScanTile - 1) / kScanTile;
    double *tile_sums = nullptr;
    QJ_CUDA_OK(cudaMallocAsync(&tile_sums, sizeof(double) * size_t(ntiles), h->stream));
    double *d_u = nullptr;
    QJ_CUDA_OK(cudaMallocAsync(&d_u, sizeof(double) * size_t(std::max<int64_t>(nshots, 1)), h->stream));
    QJ_CUDA_OK(cudaMemcpyAsync(d_u, uniforms, sizeof(double) * size_t(nshots), cudaMemcpyHostToDevice, h->stream));
    int rc = launch_checked(h, [&] {
        if (real_dtype == QJ_C128) k_scan_tile_sums<double><<<(unsigned)ntiles, kThreads, 0, h->stream>>>(reinterpret_cast<const double *>(probs), n, tile_sums);
        else k_scan_tile_sums<float><<<(unsigned)ntiles, kThreads, 0, h->stream>>>(reinterpret_cast<const float *>(probs), n, tile_sums);
    });
    if (!rc) rc = launch_checked(h, [&] { k_scan_offsets<<<1, kThreads, 0, h->stream>>>(tile_sums, ntiles); });
    if (!rc) rc = launch_checked(h, [&] {
        if (real_dtype == QJ_C128) k_scan_apply<double><<<(unsigned)ntiles, kThreads, 0, h->stream>>>(reinterpret_cast<const double *>(probs), n, tile_sums, cdf_scratch);
        else k_scan_apply<float><<<(unsigned)ntiles, kThreads, 0, h->stream>>>(reinterpret_cast<const float *>(probs), n, tile_sums, cdf_scratch);
    });
    if (!rc && nshots > 0) rc = launch_checked(h, [&] {
        k_searchsorted<<<persistent_grid(h, nshots, kThreads), kThreads, 0, h->stream>>>(cdf_scratch, n, d_u, nshots, shots);
    });
    cudaFreeAsync(tile_sums, h->stream);
    cudaFreeAsync(d_u, h->stream);
    if (rc) return rc;
    QJ_CUDA_OK(cudaStreamSynchronize(h->stream));
    return QJ_OK;
}

namespace {
int swap_geom(int dtype, int nlocal, int m, int *mv, int64_t *ngroups) {
    // work in 16-byte vectors: complex128 = 1 amplitude, complex64 = 2 amplitudes
    const int v = (dtype == QJ_C128) ? 0 : 1;
    if (m < v) return fail(QJ_ERR_UNSUPPORTED, "swap on index bit 0 of a complex64 shard: choose another local partner bit");
    if (m >= nlocal) return fail(QJ_ERR_INVALID, "local bit out of range");
    *mv = m - v;
    *ngroups = int64_t(1) << (nlocal - v - 1);
    return QJ_OK;
}
}  // namespace

extern "C" int qj_swap_pieces_peer(qj_handle *h, void *local, void *peer, int dtype, int nlocal, int m,
                                   int is_upper) {
    qj::DeviceGuard device_guard(h);
    QJ_REQUIRE(h && local && peer, "null argument");
    int mv; int64_t ngroups;
    int rc = swap_geom(dtype, nlocal, m, &mv, &ngroups);
    if (rc) return rc;
    // piece0[i + tk] <-> piece1[i]: the lower rank owns bit m = 1, the upper rank bit m = 0
    const int64_t local_bit = is_upper ? 0 : (int64_t(1) << mv);
    const int64_t peer_bit = is_upper ? (int64_t(1) << mv) : 0;
    const int64_t half = ngroups / 2;
    const int64_t g_begin = is_upper ? half : 0;
    const int64_t g_end = is_upper ? ngroups : half;
    if (g_end <= g_begin) {  // single group: the lower rank moves it
        if (is_upper) return QJ_OK;
        return launch_checked(h, [&] {
            k_swap_peer<<<1, kThreads, 0, h->stream>>>(reinterpret_cast<float4 *>(local), reinterpret_cast<float4 *>(peer), mv, local_bit, peer_bit, 0, ngroups);
        });
    }
    return launch_checked(h, [&] {
        k_swap_peer<<<persistent_grid(h, g_end - g_begin, kThreads * 4), kThreads, 0, h->stream>>>(
            reinterpret_cast<float4 *>(local), reinterpret_cast<float4 *>(peer), mv, local_bit, peer_bit, g_begin, g_end);
    });
}

namespace {
int swap_pack_impl(qj_handle *h, const void *local, void *buf, int dtype, int nlocal, int m, int is_upper,
                   int64_t begin, int64_t len, int unpack) {
    QJ_REQUIRE(h && local && buf, "null argument");
    int mv; int64_t ngroups;
    int rc = swap_geom(dtype, nlocal, m, &mv, &ngroups);
    if (rc) return rc;
    const int v = (dtype == QJ_C128) ? 0 : 1;
    QJ_REQUIRE(begin >= 0 && len >= 0 && ((begin | len) & ((1 << v) - 1)) == 0, "chunk must be vector aligned");
    const int64_t gb = begin >> v, cnt = len >> v;
    QJ_REQUIRE(gb + cnt <= ngroups, "chunk outside the half shard");
    if (cnt == 0) return QJ_OK;
    const int64_t bit = is_upper ? 0 : (int64_t(1) << mv);
    return launch_checked(h, [&] {
        k_swap_pack<<<persistent_grid(h, cnt, kThreads * 4), kThreads, 0, h->stream>>>(
            reinterpret_cast<const float4 *>(local), reinterpret_cast<float4 *>(buf), mv, bit, gb, cnt, unpack);
    });
}
}  // namespace

extern "C" int qj_swap_pack(qj_handle *h, const void *local, void *buf, int dtype, int nlocal, int m,
                            int is_upper, int64_t chunk_begin, int64_t chunk_len) {
    qj::DeviceGuard device_guard(h);
    return swap_pack_impl(h, local, buf, dtype, nlocal, m, is_upper, chunk_begin, chunk_len, 0);
}
extern "C" int qj_swap_unpack(qj_handle *h, void *local, const void *buf, int dtype, int nlocal, int m,
                              int is_upper, int64_t chunk_begin, int64_t chunk_len) {
    qj::DeviceGuard device_guard(h);
    return swap_pack_impl(h, local, const_cast<void *>(buf), dtype, nlocal, m, is_upper, chunk_begin, chunk_len, 1);
}

namespace {
int swap_pack_bits_impl(qj_handle *h, const void *local, void *buf, int dtype, int nlocal, const int32_t *bits,
                        int nbits, int value, int64_t begin, int64_t len, int unpack) {
    QJ_REQUIRE(h && local && buf && bits, "null argument");
    QJ_REQUIRE(nbits >= 1 && nbits <= QJ_MAX_GLOBAL_SWAP && nbits < nlocal, "bad number of exchange bits");
    QJ_REQUIRE(value >= 0 && value < (1 << nbits), "sub-block value out of range");
    const int v = (dtype == QJ_C128) ? 0 : 1;
    SubBlock sb;
    sb.npos = nbits;
    sb.ormask = 0;
    for (int i = 0; i < nbits; i++) {
        if (bits[i] < v) return fail(QJ_ERR_UNSUPPORTED, "swap on index bit 0 of a complex64 shard: choose another local partner bit");
        QJ_REQUIRE(bits[i] < nlocal && (i == 0 || bits[i] > bits[i - 1]), "exchange bits must be ascending and below nlocal");
        sb.pos[i] = bits[i] - v;
        if ((value >> i) & 1) sb.ormask |= int64_t(1) << (bits[i] - v);
    }
    QJ_REQUIRE(begin >= 0 && len >= 0 && ((begin | len) & ((1 << v) - 1)) == 0, "chunk must be vector aligned");
    const int64_t gb = begin >> v, cnt = len >> v;
    QJ_REQUIRE(gb + cnt <= (int64_t(1) << (nlocal - v - nbits)), "chunk outside the sub-block");
    if (cnt == 0) return QJ_OK;
    return launch_checked(h, [&] {
        k_swap_pack_bits<<<persistent_grid(h, cnt, kThreads * 4), kThreads, 0, h->stream>>>(
            reinterpret_cast<const float4 *>(local), reinterpret_cast<float4 *>(buf), sb, gb, cnt, unpack);
    });
}
}  // namespace

extern "C" int qj_swap_bits_peer(qj_handle *h, void *local, void *peer, int dtype, int nlocal, const int32_t *bits,
                                 int nbits, int local_value, int peer_value, int part, int nparts) {
    qj::DeviceGuard device_guard(h);
    QJ_REQUIRE(h && local && peer && bits, "null argument");
    QJ_REQUIRE(dtype == QJ_C64 || dtype == QJ_C128, "dtype must be QJ_C64 or QJ_C128");
    QJ_REQUIRE(nbits >= 1 && nbits <= QJ_MAX_GLOBAL_SWAP && nbits < nlocal, "bad number of exchange bits");
    QJ_REQUIRE(local_value >= 0 && local_value < (1 << nbits) && peer_value >= 0 && peer_value < (1 << nbits),
               "sub-block value out of range");
    QJ_REQUIRE(nparts >= 1 && part >= 0 && part < nparts, "bad part");
    const int v = (dtype == QJ_C128) ? 0 : 1;
    SubBlock sb;
    sb.npos = nbits;
    sb.ormask = 0;
    int64_t lmask = 0, pmask = 0;
    for (int i = 0; i < nbits; i++) {
        if (bits[i] < v) return fail(QJ_ERR_UNSUPPORTED, "swap on index bit 0 of a complex64 shard: choose another local partner bit");
        QJ_REQUIRE(bits[i] < nlocal && (i == 0 || bits[i] > bits[i - 1]), "exchange bits must be ascending and below nlocal");
        sb.pos[i] = bits[i] - v;
        if ((local_value >> i) & 1) lmask |= int64_t(1) << (bits[i] - v);
        if ((peer_value >> i) & 1) pmask |= int64_t(1) << (bits[i] - v);
    }
    const int64_t ngroups = int64_t(1) << (nlocal - v - nbits);
    const int64_t per = (ngroups + nparts - 1) / nparts;
    const int64_t g_begin = std::min<int64_t>(ngroups, per * part), g_end = std::min<int64_t>(ngroups, per * (part + 1));
    if (g_end <= g_begin) return QJ_OK;
    return launch_checked(h, [&] {
        k_swap_bits_peer<<<persistent_grid(h, g_end - g_begin, kThreads * 8), kThreads, 0, h->stream>>>(
            reinterpret_cast<float4 *>(local), reinterpret_cast<float4 *>(peer), sb, lmask, pmask, g_begin, g_end);
    });
}

// ---- CUDA IPC: a rank exports its shard allocation, its peers map it (one process per GPU)
extern "C" int qj_ipc_export(const void *ptr, void *handle64, int64_t *offset) {
    QJ_REQUIRE(ptr && handle64 && offset, "null argument");
    // the allocation that holds `ptr` (torch's caching allocator hands out pieces of larger blocks)
    typedef int (*range_fn)(unsigned long long *, size_t *, unsigned long long);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    QJ_CUDA_OK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
    QJ_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuMemGetAddressRange is not available");
    unsigned long long base = 0;
    size_t size = 0;
    if (reinterpret_cast<range_fn>(fn)(&base, &size, (unsigned long long)(uintptr_t)ptr) != 0)
        return fail(QJ_ERR_CUDA, "cuMemGetAddressRange failed");
    cudaIpcMemHandle_t hd;
    QJ_CUDA_OK(cudaIpcGetMemHandle(&hd, reinterpret_cast<void *>(uintptr_t(base))));
    static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &hd, 64);
    *offset = int64_t((unsigned long long)(uintptr_t)ptr - base);
    return QJ_OK;
}
extern "C" int qj_ipc_open(const void *handle64, void **base_out) {
    QJ_REQUIRE(handle64 && base_out, "null argument");
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle64, 64);
    QJ_CUDA_OK(cudaIpcOpenMemHandle(base_out, hd, cudaIpcMemLazyEnablePeerAccess));
    return QJ_OK;
}
extern "C" int qj_ipc_close(void *base) {
    if (base) QJ_CUDA_OK(cudaIpcCloseMemHandle(base));
    return QJ_OK;
}

extern "C" int qj_peer_handshake(qj_handle *h, void *local_flags, void *const *remote_slots, const int32_t *src_ranks,
                                 const uint32_t *epochs, int npeers, double timeout_seconds) {
    qj::DeviceGuard device_guard(h);
    QJ_REQUIRE(h && local_flags && remote_slots && src_ranks && epochs, "null argument");
    QJ_REQUIRE(npeers >= 1 && npeers <= 64, "between 1 and 64 peers per handshake");
    PeerFlags pf;
    pf.n = npeers;
    for (int i = 0; i < npeers; i++) {
        QJ_REQUIRE(remote_slots[i] != nullptr && src_ranks[i] >= 0 && src_ranks[i] < 64, "bad peer slot");
        pf.remote_slot[i] = reinterpret_cast<unsigned *>(remote_slots[i]);
        pf.src_rank[i] = src_ranks[i];
        pf.epoch[i] = epochs[i];
    }
    const unsigned long long ns = (unsigned long long)(1e9 * (timeout_seconds > 0 ? timeout_seconds : 30.0));
    return launch_checked(h, [&] {
        k_peer_handshake<<<1, 64, 0, h->stream>>>(reinterpret_cast<volatile unsigned *>(local_flags), pf, ns);
    });
}

extern "C" int qj_copy_async(qj_handle *h, void *dst, const void *src, int64_t bytes) {
    qj::DeviceGuard device_guard(h);
    QJ_REQUIRE(h && dst && src && bytes >= 0, "bad argument");
    if (bytes == 0) return QJ_OK;
    // (copy engine; either side may be a peer allocation mapped with qj_ipc_open)
    QJ_CUDA_OK(cudaMemcpyAsync(dst, src, size_t(bytes), cudaMemcpyDefault, h->stream));
    return QJ_OK;
}

extern "C" int qj_swap_pack_bits(qj_handle *h, const void *local, void *buf, int dtype, int nlocal,
                                 const int32_t *bits, int nbits, int value, int64_t chunk_begin, int64_t chunk_len) {
    qj::DeviceGuard device_guard(h);
    return swap_pack_bits_impl(h, local, buf, dtype, nlocal, bits, nbits, value, chunk_begin, chunk_len, 0);
}
extern "C" int qj_swap_unpack_bits(qj_handle *h, void *local, const void *buf, int dtype, int nlocal,
                                   const int32_t *bits, int nbits, int value, int64_t chunk_begin, int64_t chunk_len) {
    qj::DeviceGuard device_guard(h);
    return swap_pack_bits_impl(h, local, const_cast<void *>(buf), dtype, nlocal, bits, nbits, value, chunk_begin,
                               chunk_len, 1);
}
