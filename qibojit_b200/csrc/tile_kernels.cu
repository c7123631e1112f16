// tile_kernels.cu -- shared-memory staged dense k-target kernel (k <= 5), sm_100a.
//
// The fused-block kernel of the design: one HBM pass per block of gates, independent of
// where the target qubits sit.  A tile is the set of 2^T amplitudes that share all index bits
// except (a) the lowest r bits -- a contiguous run, so global traffic is made of whole
// 2^r-amplitude segments -- and (b) the target bits above the run.  Every run of a tile is
// moved global -> shared by one 1-D bulk asynchronous copy (TMA engine, `cp.async.bulk`,
// completion on an mbarrier) and back by a bulk store, so no LSU instruction touches global
// memory and coalescing no longer depends on the target positions.  Persistent CTAs walk the
// tiles through a ring of kStages buffers: loads run two tiles ahead of the math, stores
// drain one tile behind it.
//
// Inside a tile, threads gather their 2^k tuple from shared memory into registers, multiply by
// the gate matrix (kernel parameters -> constant bank) and scatter in place.  When a tile has
// fewer tuples than threads the output rows of a tuple are split across warps (warp-uniform
// row slices keep the matrix operands uniform).
//
// Arithmetic contract: gates.py:266-424 (`apply_*_qubit_gate_kernel`), gates.py:16-38, 118-193.

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace qj {
namespace {

constexpr int kTileThreads = 128;
constexpr int kStages = 4;

struct TileGeom {
    int T;                 // tile index bits
    int r;                 // run bits (low, contiguous)
    int nh;                // target bits above the run
    int hibit[kMaxDirectTargets];
    int64_t ntiles;
    int npos;              // zero-insert positions of the tile index, relative to bit r
    int pos[kMaxPos];
    int64_t cmask;         // control bits above the run, relative to bit r
    int lcmask;            // control bits inside the run (local index space)
    int ntpos;             // local positions of the targets, ascending
    int tpos[kMaxDirectTargets];
    int loff[1 << kMaxDirectTargets];  // local offset of matrix index e
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar` (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ int64_t tile_base(int64_t tile, const TileGeom &tg) {
    int64_t g = tile;
#pragma unroll 1
    for (int j = 0; j < tg.npos; j++) {
        const int p = tg.pos[j];
        const int64_t lo = g & ((int64_t(1) << p) - 1);
        g = ((g >> p) << (p + 1)) | lo;
    }
    return (g | tg.cmask) << tg.r;
}

// rows [S*RPT, (S+1)*RPT) of the tuple: fully unrolled so every matrix element is an
// immediate constant-bank operand of its FMA
template <typename T, int K, int SL, int S>
__device__ __forceinline__ void tile_rows(const Cx<T> (&x)[1 << K], Cx<T> *__restrict__ tile, int lb,
                                          const TileGeom &tg, const CMat<T, (1 << K)> &mat) {
    constexpr int NE = 1 << K;
    constexpr int RPT = NE >> SL;
#pragma unroll
    for (int ii = 0; ii < RPT; ii++) {
        constexpr int dummy = 0;
        (void)dummy;
        const int i = S * RPT + ii;
        T ar = T(0), ai = T(0);
#pragma unroll
        for (int j = 0; j < NE; j++) {
            const T gr = mat.v[2 * (i * NE + j)], gi = mat.v[2 * (i * NE + j) + 1];
            ar = fma(gr, x[j].re, ar);
            ar = fma(-gi, x[j].im, ar);
            ai = fma(gr, x[j].im, ai);
            ai = fma(gi, x[j].re, ai);
        }
        Cx<T> y; y.re = ar; y.im = ai;
        tile[lb + tg.loff[i]] = y;
    }
}

// SL = log2 of the row slices a tuple is split into (slice index is warp-uniform)
template <typename T, int K, int SL>
__global__ void __launch_bounds__(kTileThreads)
k_dense_tile(Cx<T> *__restrict__ state, const __grid_constant__ TileGeom tg,
             const __grid_constant__ CMat<T, (1 << K)> mat) {
    constexpr int NE = 1 << K;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kStages];

    const int tid = threadIdx.x;
    const int tile_amps = 1 << tg.T;
    const uint32_t run_bytes = uint32_t(sizeof(Cx<T>)) << tg.r;
    const uint32_t tile_bytes = uint32_t(sizeof(Cx<T>)) << tg.T;
    const int nruns = 1 << tg.nh;  // <= 32: lane h of warp 0 moves run h
    Cx<T> *bufs = reinterpret_cast<Cx<T> *>(smem_raw);

    if (tid == 0) {
        for (int s = 0; s < kStages; s++) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // my tiles: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int64_t first = blockIdx.x;
    const int64_t step = gridDim.x;
    const int64_t my_count = (tg.ntiles > first) ? (tg.ntiles - first + step - 1) / step : 0;

    // warp 0 moves the data; the run a lane owns is fixed for the whole kernel
    int64_t my_goff = 0;
    for (int b = 0; b < tg.nh; b++) my_goff |= int64_t((tid >> b) & 1) << tg.hibit[b];
    const size_t my_soff = size_t(tid & 31) << tg.r;
    const bool mover = tid < nruns;

    auto issue_load = [&](int64_t it) {
        const int s = int(it % kStages);
        const int64_t base = tile_base(first + it * step, tg);
        if (tid == 0) mbar_expect_tx(&full_bar[s], tile_bytes);
        __syncwarp();
        if (mover)
            bulk_g2s(bufs + size_t(s) * tile_amps + my_soff, state + base + my_goff, run_bytes, &full_bar[s]);
    };
    auto issue_store = [&](int64_t it) {
        const int s = int(it % kStages);
        const int64_t base = tile_base(first + it * step, tg);
        if (mover) bulk_s2g(state + base + my_goff, bufs + size_t(s) * tile_amps + my_soff, run_bytes);
        bulk_commit();
    };

    if (tid < 32) {
        for (int64_t it = 0; it < kStages - 2 && it < my_count; it++) issue_load(it);
    }

    // compute-phase geometry: 2^(T-K) tuples, each split into 2^SL row slices
    constexpr int kThreadsLog = 7;
    const int ntuples_log = tg.T - K;
    const int rounds = (ntuples_log + SL > kThreadsLog) ? (1 << (ntuples_log + SL - kThreadsLog)) : 1;
    const int slice = (SL == 0) ? 0 : (tid >> (kThreadsLog - SL));
    const int tuple_lane = (SL == 0) ? tid : (tid & ((1 << (kThreadsLog - SL)) - 1));

    for (int64_t it = 0; it < my_count; it++) {
        const int s = int(it % kStages);
        if (tid < 32) {
            // buffer of tile it+kStages-2 was last used by tile it-2: its store must have
            // finished reading shared memory (at most the newest group may be pending)
            const int64_t nxt = it + kStages - 2;
            if (nxt < my_count) {
                bulk_wait_read<1>();
                __syncwarp();
                issue_load(nxt);
            }
        }
        mbar_wait(&full_bar[s], uint32_t((it / kStages) & 1));
        Cx<T> *tile = bufs + size_t(s) * tile_amps;

        for (int rd = 0; rd < rounds; rd++) {
            // local base: insert zeros at the target positions
            int lb = (rd << (kThreadsLog - SL)) | tuple_lane;
#pragma unroll
            for (int j = 0; j < K; j++) {
                const int p = tg.tpos[j];
                lb = ((lb >> p) << (p + 1)) | (lb & ((1 << p) - 1));
            }
            const bool active = (lb & tg.lcmask) == tg.lcmask;
            Cx<T> x[NE];
            if (active) {
#pragma unroll
                for (int e = 0; e < NE; e++) x[e] = tile[lb + tg.loff[e]];
            }
            if (SL > 0) __syncthreads();
            if (active) {
                if (SL == 0) {
                    tile_rows<T, K, 0, 0>(x, tile, lb, tg, mat);
                } else if (SL == 1) {
                    if (slice == 0) tile_rows<T, K, SL, 0>(x, tile, lb, tg, mat);
                    else tile_rows<T, K, SL, (SL >= 1 ? 1 : 0)>(x, tile, lb, tg, mat);
                } else {
                    switch (slice) {
                        case 0: tile_rows<T, K, SL, 0>(x, tile, lb, tg, mat); break;
                        case 1: tile_rows<T, K, SL, (SL >= 2 ? 1 : 0)>(x, tile, lb, tg, mat); break;
                        case 2: tile_rows<T, K, SL, (SL >= 2 ? 2 : 0)>(x, tile, lb, tg, mat); break;
                        default: tile_rows<T, K, SL, (SL >= 2 ? 3 : 0)>(x, tile, lb, tg, mat); break;
                    }
                }
            }
            if (SL > 0 && rounds > 1) __syncthreads();
        }
        fence_async_smem();
        __syncthreads();
        if (tid < 32) issue_store(it);
    }
    if (tid < 32) bulk_wait_all();
}

struct TilePlan {
    TileGeom tg;
    size_t smem;
    unsigned grid;
};

bool plan_tile(const qj_handle *h, const GateCall &c, TilePlan *out) {
    const int K = c.ntargets;
    if (K < 1 || K > kMaxDirectTargets) return false;
    const int amp_log = (c.dtype == QJ_C128) ? 4 : 3;
    int T = std::max(K + 5, 14 - amp_log);  // 16 KiB tiles unless k forces more
    if (T > c.nqubits) T = c.nqubits;
    if (T < K + 5 || T < 8) return false;   // too small a register for this kernel
    if (((size_t(1) << T) << amp_log) * kStages > size_t(200) << 10) return false;
    TileGeom &tg = out->tg;
    memset(&tg, 0, sizeof(tg));
    // grow the run until run + high targets fill the tile
    std::vector<int> t(c.tbits, c.tbits + K);
    std::sort(t.begin(), t.end());
    int r = T;
    for (;;) {
        int nh = 0;
        for (int b : t) nh += (b >= r);
        if (r + nh <= T) break;
        r--;
    }
    // r + nh may be < T when targets fall inside the run: then use the larger run
    int nh = 0;
    for (int b : t) if (b >= r) tg.hibit[nh++] = b;
    while (r + nh < T) {  // widen the run while it does not swallow a high target
        bool clash = false;
        for (int b = 0; b < nh; b++) clash |= (tg.hibit[b] == r);
        if (clash) {  // the run reaches the lowest high target: it becomes part of the run
            for (int b = 0; b + 1 < nh; b++) tg.hibit[b] = tg.hibit[b + 1];
            nh--;
        }
        r++;
    }
    if ((size_t(1) << r) << amp_log < 16) return false;
    tg.T = r + nh;
    tg.r = r;
    tg.nh = nh;
    if (tg.T < K + 5 || tg.T < 8) return false;
    // tile index: every bit >= r that is neither a high target nor a control
    std::vector<int> fixed;
    for (int b = 0; b < nh; b++) fixed.push_back(tg.hibit[b] - r);
    for (int i = 0; i < c.ncontrols; i++) {
        if (c.cbits[i] >= r) {
            fixed.push_back(c.cbits[i] - r);
            tg.cmask |= int64_t(1) << (c.cbits[i] - r);
        } else {
            tg.lcmask |= 1 << c.cbits[i];
        }
    }
    std::sort(fixed.begin(), fixed.end());
    tg.npos = (int)fixed.size();
    for (int i = 0; i < tg.npos; i++) tg.pos[i] = fixed[i];
    const int tile_index_bits = c.nqubits - r - tg.npos;
    if (tile_index_bits < 0) return false;
    tg.ntiles = int64_t(1) << tile_index_bits;
    // local position of every target: inside the run it keeps its bit, above it is r + rank
    auto local_pos = [&](int bit) {
        if (bit < r) return bit;
        for (int b = 0; b < nh; b++) if (tg.hibit[b] == bit) return r + b;
        return -1;
    };
    std::vector<int> lp;
    for (int u = 0; u < K; u++) lp.push_back(local_pos(c.tbits[u]));
    for (int e = 0; e < (1 << K); e++) {
        int o = 0;
        for (int u = 0; u < K; u++) if ((e >> u) & 1) o |= 1 << lp[u];
        tg.loff[e] = o;
    }
    std::sort(lp.begin(), lp.end());
    tg.ntpos = K;
    for (int u = 0; u < K; u++) tg.tpos[u] = lp[u];
    out->smem = ((size_t(1) << tg.T) << amp_log) * kStages;
    out->grid = 0;  // filled by the launcher from the occupancy of the instantiation
    (void)h;
    return true;
}

template <typename T, int K, int SL>
int launch_tile_ks(qj_handle *h, const GateCall &c, const TilePlan &p) {
    constexpr int NE = 1 << K;
    CMat<T, NE> mat;
    memcpy(mat.v, c.gate, sizeof(mat.v));
    static bool configured[kMaxDevices] = {false};   // the attribute is per device, not per process
    static int per_sm_cached[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (!configured[h->device]) {
        QJ_CUDA_OK(cudaFuncSetAttribute(k_dense_tile<T, K, SL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10));
        configured[h->device] = true;
    }
    const int slot = std::min<int>(7, int(p.smem >> 15));
    if (per_sm_cached[slot] == 0) {
        int per_sm = 1;
        QJ_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dense_tile<T, K, SL>, kTileThreads, p.smem));
        per_sm_cached[slot] = std::max(1, per_sm);
    }
    const unsigned grid = (unsigned)std::min<int64_t>(p.tg.ntiles, int64_t(h->sm_count) * per_sm_cached[slot]);
    k_dense_tile<T, K, SL><<<grid, kTileThreads, p.smem, h->stream>>>(reinterpret_cast<Cx<T> *>(c.state), p.tg, mat);
    h->launches++;
    QJ_CUDA_OK(cudaGetLastError());
    return QJ_OK;
}

// slices needed so that 2^(T-K) tuples x 2^SL slices cover the 128 threads of a CTA
template <typename T, int K>
int launch_tile_k(qj_handle *h, const GateCall &c, const TilePlan &p) {
    const int sl = std::max(0, 7 - (p.tg.T - K));
    if (sl > 2 || sl > K) return fail(QJ_ERR_UNSUPPORTED, "tile too small for the tile kernel");
    if (sl == 0) return launch_tile_ks<T, K, 0>(h, c, p);
    if constexpr (K >= 1) { if (sl == 1) return launch_tile_ks<T, K, 1>(h, c, p); }
    if constexpr (K >= 2) { if (sl == 2) return launch_tile_ks<T, K, 2>(h, c, p); }
    return fail(QJ_ERR_UNSUPPORTED, "tile kernel slice configuration");
}

template <typename T>
int launch_tile_t(qj_handle *h, const GateCall &c, const TilePlan &p) {
    switch (c.ntargets) {
        case 1: return launch_tile_k<T, 1>(h, c, p);
        case 2: return launch_tile_k<T, 2>(h, c, p);
        case 3: return launch_tile_k<T, 3>(h, c, p);
        case 4: return launch_tile_k<T, 4>(h, c, p);
        case 5: return launch_tile_k<T, 5>(h, c, p);
    }
    return fail(QJ_ERR_INVALID, "tile kernel supports 1..5 targets");
}

}  // namespace

bool tile_kernel_applies(const qj_handle *h, const GateCall &c) {
    TilePlan p;
    return plan_tile(h, c, &p);
}

int launch_dense_tile(qj_handle *h, const GateCall &c) {
    TilePlan p;
    if (!plan_tile(h, c, &p)) return fail(QJ_ERR_UNSUPPORTED, "tile kernel does not apply to this gate");
    if (c.dtype == QJ_C128) return launch_tile_t<double>(h, c, p);
    return launch_tile_t<float>(h, c, p);
}

}  // namespace qj
