// pass_kernels.cu -- multi-gate passes ("tile programs") for sm_100a.
//
// The per-gate kernels (gate_kernels.cu) already run at the HBM roofline, so the time of a
// circuit is the number of passes over the state.  This kernel removes passes: a CTA stages a
// tile of 2^T amplitudes in shared memory -- the low r index bits (a contiguous run, so global
// traffic is made of whole 2^r-amplitude segments) plus T-r arbitrary higher index bits -- and
// applies a whole PROGRAM of gates to it before writing it back.
//
// Execution model (the schedule itself is made on the host, qibojit_b200/planner.py):
//   * the program of a pass is a list of ROUNDS.  In a round every thread owns 16 16-byte
//     vectors of the tile = the 2^J amplitudes spanned by J "register" bits (complex128: J = 4;
//     complex64: J = 5, index bit 0 -- the second amplitude of a vector -- always being one of
//     them).  It gathers them from shared memory, applies every op of the round in registers,
//     and scatters them back: shared memory is read and written once per round, not per gate.
//   * ops: one-target gates on a SET of register slots (one dispatch for up to J gates of a
//     kind: complex / real / real-diagonal + imaginary-off-diagonal), dense 4x4 on two slots,
//     X / SWAP as register renaming, and phase groups.  Controls may be register slots
//     (-> element mask), other tile bits (-> thread predicate) or bits outside the tile
//     (-> tile predicate); a failed predicate zeroes the element mask, so the op loop has
//     warp-uniform control flow only.
//   * phase groups: the diagonal tables of a run that multiply the same elements are one op,
//     phase = G[tid] * H[tile] * (mixed per-thread look-ups).  G is the product of the tables
//     over tile-local bits only, precomputed on the host per thread (a thread's tile position
//     is the same in every tile); H is the product of the tables over outer bits only, computed
//     once per tile.  +-1 factors are sign flips.  Consecutive groups run inside one dispatch.
//   * complex64 arithmetic is packed FP32x2 (FFMA2 / FMUL2): an amplitude is one 64-bit operand.
//   * the whole program (rounds, op headers, gate matrices) is copied to shared memory once
//     per CTA and read with warp-uniform (broadcast) loads; tables and G stay in global memory
//     (L1, evict_last).
//   * shared-memory layout: 16-byte vectors, XOR swizzle of the low three vector-index bits with
//     the three-bit groups above them; the host orders the thread bits of every round so that
//     the eight lanes of a quarter warp hit eight different bank groups whatever the register
//     bits are.  Tile IO (cp.async in, LDS + STG out) splits its address arithmetic into a
//     per-thread and a warp-uniform part (the swizzle is XOR-linear).
//   * one thread per 16 vectors of the tile, 128 registers: complex128 runs four 128-thread
//     CTAs with 32 KiB tiles per SM, complex64 two 256-thread CTAs with 64 KiB tiles; while one
//     CTA waits for its tile or drains its stores the others compute.
//
// Arithmetic contract per op: gates.py:16-38 (one target), gates.py:118-193 (two targets),
// gates.py:82-114 (diagonals, pre-multiplied into tables on the host).

#include <algorithm>
#include <complex>
#include <vector>

#include "pass_device.cuh"

namespace qj {
namespace {

// ------------------------------------------------------------------ host-side encoding
typedef std::complex<double> cd;

template <typename T>
cd load_cx(const unsigned char *p, int64_t i) {
    const T *t = reinterpret_cast<const T *>(p);
    return cd(double(t[2 * i]), double(t[2 * i + 1]));
}

struct Unit {
    uint32_t w[4];
};

struct Encoder {
    int dtype;
    size_t esz;
    const unsigned char *hdata;
    int64_t ndata;
    std::vector<unsigned char> tables;   // device phase tables, state dtype

    cd data_at(int64_t off) const {
        return dtype == QJ_C128 ? load_cx<double>(hdata, off) : load_cx<float>(hdata, off);
    }
    uint32_t push_table(const std::vector<cd> &t) {
        const uint32_t off = uint32_t(tables.size() / esz);
        const size_t at = tables.size();
        tables.resize(at + t.size() * esz);
        for (size_t i = 0; i < t.size(); i++) {
            if (dtype == QJ_C128) {
                double *d = reinterpret_cast<double *>(tables.data() + at) + 2 * i;
                d[0] = t[i].real(); d[1] = t[i].imag();
            } else {
                float *d = reinterpret_cast<float *>(tables.data() + at) + 2 * i;
                d[0] = float(t[i].real()); d[1] = float(t[i].imag());
            }
        }
        return off;
    }
    // complex matrix elements: complex128 (re, im); complex64 the ready-made FP32x2 operand pairs
    // (re, re, -im, im), 16 bytes per element either way
    void push_complex(std::vector<Unit> &out, const std::vector<cd> &z) const {
        std::vector<double> sc;
        for (const cd &g : z) {
            if (dtype == QJ_C128) {
                sc.push_back(g.real()); sc.push_back(g.imag());
            } else {
                sc.push_back(g.real()); sc.push_back(g.real()); sc.push_back(-g.imag()); sc.push_back(g.imag());
            }
        }
        push_scalars(out, sc);
    }
    // scalars (state precision) appended to a unit stream
    void push_scalars(std::vector<Unit> &out, const std::vector<double> &s) const {
        std::vector<unsigned char> raw;
        if (dtype == QJ_C128) {
            raw.resize(s.size() * 8);
            memcpy(raw.data(), s.data(), raw.size());
        } else {
            raw.resize(s.size() * 4);
            for (size_t i = 0; i < s.size(); i++) {
                const float f = float(s[i]);
                memcpy(raw.data() + 4 * i, &f, 4);
            }
        }
        raw.resize((raw.size() + 15) / 16 * 16, 0);
        for (size_t i = 0; i < raw.size(); i += 16) {
            Unit u;
            memcpy(u.w, raw.data() + i, 16);
            out.push_back(u);
        }
    }
};

inline int pair_index(int a, int b) { return b * (b - 1) / 2 + a; }

}  // namespace
}  // namespace qj

struct qj_program {
    int dtype = 0;
    int nqubits = 0;
    struct Launch {
        qj::PassGeom geom;
        int64_t blob_off = 0;   // units into h_blob
        size_t smem = 0;
        int nrounds = 0, nops = 0;
    };
    std::vector<Launch> launches;
    std::vector<qj::ProgParam> images;   // one kernel-parameter image per launch (host memory)
    void *d_tables = nullptr;
    int64_t total_mops = 0, total_rounds = 0;
};

using namespace qj;

// host image of an encoded program (qj_program_encode: inspection / CPU tests of the encoder)
struct qj_program_image {
    int dtype = 0, nqubits = 0;
    std::vector<Unit> blob;
    std::vector<unsigned char> tables;
    std::vector<qj_program::Launch> launches;
};

// Encodes the passes; uploads them into a qj_program (h, out) or, when `image` is given, hands the
// host image back without touching a device.
static int program_build(qj_handle *h, int dtype, int nqubits, const qj_pass_desc *passes,
                         int npasses, const qj_round_desc *rounds_in, int64_t nrounds_in,
                         const qj_op_desc *ops, int64_t nops, const void *data, int64_t ndata,
                         qj_program **out, qj_program_image *image) {
    QJ_REQUIRE(((h && out) || image) && (passes || npasses == 0), "null argument");
    QJ_REQUIRE(dtype == QJ_C64 || dtype == QJ_C128, "dtype must be QJ_C64 or QJ_C128");
    QJ_REQUIRE(nqubits >= kMinTileBits && nqubits <= QJ_MAX_QUBITS, "tile programs need 6 <= nqubits <= QJ_MAX_QUBITS");
    QJ_REQUIRE(npasses >= 0 && nrounds_in >= 0 && nops >= 0 && ndata >= 0, "negative count");
    QJ_REQUIRE((rounds_in || nrounds_in == 0) && (ops || nops == 0) && (data || ndata == 0), "null argument");

    const int VS = (dtype == QJ_C128) ? 0 : 1;
    const int J = (dtype == QJ_C128) ? 4 : 5;
    const int N = 1 << J;
    Encoder enc;
    enc.dtype = dtype;
    enc.esz = (dtype == QJ_C128) ? 16 : 8;
    enc.hdata = static_cast<const unsigned char *>(data);
    enc.ndata = ndata;

    // fused diagonals (C_DIAGF) replace the phase groups waiting at one point of a round when there
    // are at least this many of them (0: never); QJ_DIAGF_MIN overrides it for experiments
    int fuse_min = 2;
    if (const char *ev = getenv("QJ_DIAGF_MIN")) fuse_min = atoi(ev);
    const bool generic_only = fuse_min >= 100;      // 100 + n: threshold n, never the slot-factorised form (tests)
    if (generic_only) fuse_min -= 100;
    const bool hadamard_on = !generic_only && fuse_min > 0;   // (the same switch turns the butterfly form off)

    auto *prog = new qj_program();
    prog->dtype = dtype;
    prog->nqubits = nqubits;
    std::vector<Unit> blob_all;
    auto bail = [&](const std::string &msg) {
        delete prog;
        return fail(QJ_ERR_INVALID, msg);
    };

    for (int pi = 0; pi < npasses; pi++) {
        const qj_pass_desc &pd = passes[pi];
        const int T = pd.nlocal;
        if (T < kMinTileBits || T - VS > kMaxTileVecBits || T > nqubits)
            return bail("pass: need 6 <= nlocal <= min(12 (complex128) / 13 (complex64), nqubits)");
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
        if (r < 1) return bail("pass: the tile must contain index bit 0");
        // the tile IO moves one contiguous run per group of threads: a run longer than a 16th of
        // the tile is split (its upper bits are addressed like arbitrary high bits)
        r = std::max(1, std::min(r, T - 4));
        if (T - r > kMaxHiBits) return bail("pass: too many local bits above the contiguous run");
        if (pd.first_round < 0 || pd.nrounds < 0 || pd.first_round + pd.nrounds > nrounds_in)
            return bail("pass: round range out of bounds");

        PassGeom geo;
        memset(&geo, 0, sizeof(geo));
        geo.T = T; geo.r = r; geo.nh = T - r;
        for (int i = r; i < T; i++) {
            geo.hibit[i - r] = pd.local_bits[i];
            geo.pos[i - r] = pd.local_bits[i] - r;
        }
        geo.npos = T - r;
        geo.ntiles = int64_t(1) << (nqubits - T);
        const int Tv = T - VS;

        // one launch = the rounds whose image fits the shared-memory budget
        std::vector<Unit> round_units, outer_units, op_units, h_units;
        std::vector<Unit> f_dir, f_slices;   // fused diagonals with an outer part: directory + slices
        int launch_rounds = 0, launch_ops = 0;
        auto close_launch = [&]() {
            if (launch_rounds == 0) return;
            std::vector<Unit> img;
            Unit hdr, hdr1;
            memset(&hdr1, 0, sizeof(hdr1));
            const uint32_t off_rounds = 2, off_outer = off_rounds + uint32_t(round_units.size());
            const uint32_t off_h = off_outer + uint32_t(outer_units.size());
            const uint32_t off_f = off_h + uint32_t(h_units.size());
            const uint32_t off_ops = off_f + uint32_t(f_dir.size() + f_slices.size());
            hdr.w[0] = uint32_t(launch_rounds); hdr.w[1] = uint32_t(outer_units.size() / 2);
            hdr.w[2] = off_rounds; hdr.w[3] = off_outer;
            hdr1.w[0] = uint32_t(h_units.size() / 4); hdr1.w[1] = off_h;
            hdr1.w[2] = uint32_t(f_dir.size()); hdr1.w[3] = off_f;
            for (Unit &d : f_dir) d.w[0] += uint32_t(f_dir.size());   // slices follow the directory
            img.push_back(hdr);
            img.push_back(hdr1);
            for (size_t i = 0; i < round_units.size(); i += 3) round_units[i].w[0] += off_ops;   // first_unit
            img.insert(img.end(), round_units.begin(), round_units.end());
            img.insert(img.end(), outer_units.begin(), outer_units.end());
            img.insert(img.end(), h_units.begin(), h_units.end());
            img.insert(img.end(), f_dir.begin(), f_dir.end());
            img.insert(img.end(), f_slices.begin(), f_slices.end());
            img.insert(img.end(), op_units.begin(), op_units.end());
            qj_program::Launch L;
            L.geom = geo;
            L.geom.blob_units = int(img.size());
            L.geom.prefix_units = int(off_ops);
            L.blob_off = int64_t(blob_all.size());
            L.nrounds = launch_rounds;
            L.nops = launch_ops;
            L.geom.nH = int(h_units.size() / 4);
            L.geom.nF = int(f_dir.size());
            L.geom.nFS = int(f_slices.size());
            L.geom.nrounds_smem = launch_rounds;
            {   // constants of the fast tile IO (one thread per 16 vectors)
                const int rv = geo.r - VS;
                const int nvec = 1 << Tv;
                const int nthr = std::max(1, std::min(kThreads, nvec >> kVecRegBits));
                L.geom.io_fast = (nthr >= 8 && nthr >= (1 << rv) && nvec == 16 * nthr) ? 1 : 0;
                for (int u = 0; u < 16; u++) {
                    L.geom.io_soff[u] = swz_vec(uint32_t(u * nthr));
                    int64_t off = 0;
                    if (L.geom.io_fast) {
                        const int run = (u * nthr) >> rv;     // the per-iteration part of the run index
                        for (int b2 = 0; b2 < geo.nh; b2++) off |= int64_t((run >> b2) & 1) << (geo.hibit[b2] - VS);
                    }
                    L.geom.io_goff[u] = off;
                }
            }
            L.smem = (size_t(1) << Tv) * 16 + size_t(off_ops) * 16 + (h_units.size() / 4) * 16 + f_dir.size() * 256 + f_slices.size() * 16 +
                     size_t(launch_rounds) * size_t(std::max(1, std::min(kThreads, (1 << Tv) >> kVecRegBits))) * 8 +
                     (size_t(8) << geo.nh) + (outer_units.size() / 2) * 4 + 16;
            prog->launches.push_back(L);
            blob_all.insert(blob_all.end(), img.begin(), img.end());
            round_units.clear(); outer_units.clear(); op_units.clear(); h_units.clear();
            f_dir.clear(); f_slices.clear();
            launch_rounds = 0; launch_ops = 0;
        };

        int h_left = 0;               // Hadamard-like gates of this pass still to come
        double round_scale = 1.0;     // product of the scales the butterflies so far left out
        for (int64_t ri = pd.first_round; ri < pd.first_round + pd.nrounds; ri++) {
            const qj_round_desc &rdesc = rounds_in[ri];
            if (rdesc.nreg != J) return bail("round: need 4 (complex128) / 5 (complex64) register bits");
            if (rdesc.first_op < 0 || rdesc.nops < 0 || rdesc.first_op + rdesc.nops > nops)
                return bail("round: op range out of bounds");
            // register slots: slot j <-> local position rp[j] (ascending)
            int rp[8];
            int slot_of_pos[QJ_MAX_LOCAL_BITS];
            for (int i = 0; i < QJ_MAX_LOCAL_BITS; i++) slot_of_pos[i] = -1;
            for (int j = 0; j < J; j++) {
                const int b = rdesc.reg_bits[j];
                if (b < 0 || b >= nqubits || lpos[b] < 0) return bail("round: register bit is not a local bit of its pass");
                if (j && b <= rdesc.reg_bits[j - 1]) return bail("round: register bits must be strictly ascending");
                rp[j] = lpos[b];
                slot_of_pos[rp[j]] = j;
            }
            if (VS && rp[0] != 0) return bail("round: complex64 rounds must keep index bit 0 in registers");
            // vector positions of the register slots and of the thread bits
            bool is_regvec[kMaxTileVecBits + 1] = {false};
            uint32_t vd[4];
            for (int j = VS; j < J; j++) {
                const int q = rp[j] - VS;
                is_regvec[q] = true;
                vd[j - VS] = swz_vec(1u << q) << 4;
            }
            std::vector<int> tq;      // thread bit k <-> vector position tq[k]
            {
                bool used_res[3] = {false, false, false};
                std::vector<int> freeq;
                for (int q = 0; q < Tv; q++) if (!is_regvec[q]) freeq.push_back(q);
                std::vector<bool> picked(freeq.size(), false);
                for (size_t i = 0; i < freeq.size() && tq.size() < 3; i++) {
                    if (!used_res[freeq[i] % 3]) {
                        used_res[freeq[i] % 3] = true;
                        picked[i] = true;
                        tq.push_back(freeq[i]);
                    }
                }
                for (size_t i = 0; i < freeq.size(); i++) if (!picked[i]) tq.push_back(freeq[i]);
            }
            uint32_t regmask = 0;     // local positions held in registers
            for (int j = 0; j < J; j++) regmask |= 1u << rp[j];

            // would this round overflow the image?  (conservative: close before encoding)
            std::vector<Unit> r_ops;
            std::vector<Unit> r_outer;
            int r_nops = 0;
            const int outer_base = int(outer_units.size() / 2);

            auto elem_bits = [&](int e) {   // local positions set in element e
                uint32_t o = 0;
                for (int j = 0; j < J; j++) if ((e >> j) & 1) o |= 1u << rp[j];
                return o;
            };
            auto add_outer = [&](const HostOuter &ho) -> int {
                if (ho.ocmask == 0 && ho.nbits == 0) return 0xffff;
                Unit u0, u1;
                memset(&u0, 0, sizeof(u0)); memset(&u1, 0, sizeof(u1));
                u0.w[0] = uint32_t(ho.ocmask); u0.w[1] = uint32_t(ho.ocmask >> 32); u0.w[2] = uint32_t(ho.nbits);
                uint32_t srcw[3] = {0, 0, 0}, dstw[2] = {0, 0};
                for (int b = 0; b < ho.nbits; b++) {
                    srcw[b >> 2] |= uint32_t(ho.src[b]) << ((b & 3) * 8);
                    dstw[b >> 3] |= uint32_t(ho.dst[b]) << ((b & 7) * 4);
                }
                u0.w[3] = srcw[0]; u1.w[0] = srcw[1]; u1.w[1] = srcw[2]; u1.w[2] = dstw[0]; u1.w[3] = dstw[1];
                r_outer.push_back(u0); r_outer.push_back(u1);
                return outer_base + int(r_outer.size() / 2) - 1;
            };
            auto push_op = [&](uint32_t code, int oslot, int nfields, uint32_t tmask, uint32_t emask, uint32_t table,
                               const uint32_t *fields, const std::vector<Unit> &payload) {
                Unit h0, h1;
                h0.w[0] = code | (uint32_t(2 + payload.size()) << 16);
                h0.w[1] = uint32_t(oslot) | (uint32_t(nfields) << 16);
                h0.w[2] = tmask; h0.w[3] = emask;
                h1.w[0] = table;
                h1.w[1] = fields ? fields[0] : 0; h1.w[2] = fields ? fields[1] : 0; h1.w[3] = fields ? fields[2] : 0;
                r_ops.push_back(h0); r_ops.push_back(h1);
                r_ops.insert(r_ops.end(), payload.begin(), payload.end());
                r_nops++;
            };

            // diagonal slices wait here and are emitted as phase groups: the slices of a run of
            // consecutive diagonal ops commute, so those that multiply the same elements are merged
            // into ONE op (thread-level product of the look-ups, one multiply per element)
            struct PendingSlice {
                uint32_t emask, tmask, table;
                int oslot, nf;
                uint32_t fields[9];
                bool sign;
                int cls;                 // 0: thread-only (folded into G), 1: outer-only (H), 2: mixed (per-thread look-up)
                std::vector<cd> host;    // class 0: the slice's table, kept on the host
            };
            std::vector<PendingSlice> pending;
            std::vector<Unit> r_H;          // per-tile factor entries of this round (4 units each)
            const int h_base = int(h_units.size() / 4);
            std::vector<Unit> r_Fdir, r_Fsl;   // fused diagonals of this round with an outer part
            const int f_base = int(f_dir.size());
            const int nthr_round = 1 << int(tq.size());
            auto host_field = [](uint32_t base, uint32_t fl) -> uint32_t {
                return ((base >> (fl & 255u)) & ((1u << ((fl >> 8) & 255u)) - 1u)) << (fl >> 16);
            };
            std::vector<size_t> group_run;  // r_ops indices of the trailing run of plain one-target groups
            auto select_code = [&](uint32_t emask) -> uint32_t {
                const uint32_t full = (N == 32) ? 0xffffffffu : 0xffffu;
                if (emask == full) return SEL_ALL;
                for (int a = 0; a < J; a++) {
                    uint32_t m = 0;
                    for (int e = 0; e < N; e++) if ((e >> a) & 1) m |= 1u << e;
                    if (m == emask) return SEL_SLOT + uint32_t(a);
                }
                for (int b2 = 1; b2 < J; b2++)
                    for (int a = 0; a < b2; a++) {
                        uint32_t m = 0;
                        for (int e = 0; e < N; e++) if (((e >> a) & 1) && ((e >> b2) & 1)) m |= 1u << e;
                        if (m == emask) return SEL_PAIR + uint32_t(pair_index(a, b2));
                    }
                return SEL_MASK;
            };
            auto flush_pending = [&]() {
                if (!pending.empty()) group_run.clear();
                // ---- fused diagonals.  (1) Slices whose factor depends on the register bits only
                // (thread-only class with a one-entry table, no thread predicate) become ONE constant
                // diagonal op: per-unit constants in the payload, uniform operands, no loads.
                // (2) When several groups (different element sets) still wait, the thread-only slices
                // become ONE per-thread, per-unit factor array and the outer-only slices one
                // per-element, per-tile factor array: one dispatch, one multiply per element.
                auto thread_base = [&](int t) {
                    uint32_t base = 0;
                    for (size_t kb = 0; kb < tq.size(); kb++) if ((t >> kb) & 1) base |= 1u << (tq[kb] + VS);
                    return base;
                };
                if (fuse_min > 0) {
                    std::vector<PendingSlice> keep;
                    std::vector<const PendingSlice *> cm;
                    bool all_sign = true;
                    for (const PendingSlice &ps : pending) {
                        if (ps.cls == 0 && ps.tmask == 0 && ps.nf == 0) {
                            cm.push_back(&ps);
                            all_sign = all_sign && ps.sign;
                        } else {
                            keep.push_back(ps);
                        }
                    }
                    if (!cm.empty() && !all_sign) {          // (pure sign flips stay phase groups: no arithmetic)
                        std::vector<cd> fac(size_t(N), cd(1.0));
                        uint32_t emu = 0;
                        for (const PendingSlice *ps : cm) {
                            emu |= ps->emask;
                            for (int e = 0; e < N; e++) if ((ps->emask >> e) & 1u) fac[size_t(e)] *= ps->host[0];
                        }
                        uint32_t um = 0;
                        std::vector<double> sc;
                        // every selected element has one register bit in common: the mask-free form
                        // over the N / 2 elements with that bit set (constant 1 where nothing acts)
                        int common = -1;
                        if (!generic_only)
                            for (int a = 0; a < J && common < 0; a++) {
                                bool all = true;
                                for (int e = 0; e < N; e++) if (((emu >> e) & 1u) && !((e >> a) & 1)) all = false;
                                if (all) common = a;
                            }
                        uint32_t code_c = C_DIAGC;
                        if (common >= 0) {
                            code_c = uint32_t(C_DIAGCS + common);
                            for (int p2 = 0; p2 < N / 2; p2++) {
                                const int e = (((p2 >> common) << (common + 1)) | (p2 & ((1 << common) - 1))) | (1 << common);
                                sc.push_back(fac[size_t(e)].real()); sc.push_back(fac[size_t(e)].imag());
                            }
                        } else {
                            for (int u = 0; u < 16; u++) {
                                if (!((emu >> (u << VS)) & (VS ? 3u : 1u))) continue;
                                um |= 1u << u;
                                for (int j = 0; j < (1 << VS); j++) {
                                    const cd z = fac[size_t((u << VS) + j)];
                                    sc.push_back(z.real()); sc.push_back(z.imag());
                                }
                            }
                        }
                        std::vector<Unit> payload;
                        enc.push_scalars(payload, sc);
                        Unit h0, h1;
                        memset(&h1, 0, sizeof(h1));
                        h0.w[0] = code_c | (uint32_t(2 + payload.size()) << 16);
                        h0.w[1] = 0xffffu; h0.w[2] = 0; h0.w[3] = um;
                        r_ops.push_back(h0); r_ops.push_back(h1);
                        r_ops.insert(r_ops.end(), payload.begin(), payload.end());
                        r_nops++;
                        pending = keep;
                    }
                }
                {
                    std::vector<uint32_t> masks;
                    bool any_f = false;
                    for (const PendingSlice &ps : pending) {
                        if (ps.cls == 2 || ps.sign) continue;
                        if (std::find(masks.begin(), masks.end(), ps.emask) == masks.end()) masks.push_back(ps.emask);
                        any_f = any_f || ps.cls == 1;
                    }
                    if (fuse_min > 0 && int(masks.size()) >= fuse_min) {
                        std::vector<PendingSlice> keep;
                        std::vector<const PendingSlice *> gm, fm;
                        for (const PendingSlice &ps : pending) {
                            if (ps.cls == 0) gm.push_back(&ps);
                            else if (ps.cls == 1 && any_f) fm.push_back(&ps);
                            else keep.push_back(ps);
                        }
                        // every fused slice multiplies "the elements with register bit s set" for some
                        // slot s: the slot-factorised form (J factors per thread instead of 2^J)
                        auto slot_of = [&](uint32_t emask) -> int {
                            for (int a = 0; a < J; a++) {
                                uint32_t m = 0;
                                for (int e = 0; e < N; e++) if ((e >> a) & 1) m |= 1u << e;
                                if (m == emask) return a;
                            }
                            return -1;
                        };
                        bool by_slot = !generic_only;
                        for (const PendingSlice *ps : gm) by_slot = by_slot && slot_of(ps->emask) >= 0;
                        for (const PendingSlice *ps : fm) by_slot = by_slot && slot_of(ps->emask) >= 0;
                        if (by_slot) {
                            uint32_t smask = 0;
                            for (const PendingSlice *ps : gm) smask |= 1u << slot_of(ps->emask);
                            for (const PendingSlice *ps : fm) smask |= 1u << slot_of(ps->emask);
                            uint32_t g_off = 0xffffffffu;
                            if (!gm.empty()) {
                                const int per = VS ? 8 : 4;     // factors per thread (64 bytes)
                                std::vector<cd> G(size_t(nthr_round) * size_t(per), cd(1.0));   // [thread][slot]
                                for (const PendingSlice *ps : gm) {
                                    const int a = slot_of(ps->emask);
                                    for (int t = 0; t < nthr_round; t++) {
                                        const uint32_t base = thread_base(t);
                                        if ((base & ps->tmask) != ps->tmask) continue;
                                        uint32_t idx = 0;
                                        for (int f = 0; f < ps->nf; f++) idx |= host_field(base, ps->fields[f]);
                                        G[size_t(t) * size_t(per) + size_t(a)] *= ps->host[idx];
                                    }
                                }
                                if (VS && ((enc.tables.size() / enc.esz) & 1)) enc.push_table(std::vector<cd>(1, cd(1.0)));
                                g_off = enc.push_table(G);
                            }
                            uint32_t fidx = 0xffffu;
                            if (!fm.empty()) {
                                fidx = uint32_t(f_base) + uint32_t(r_Fdir.size());
                                Unit d;
                                memset(&d, 0, sizeof(d));
                                d.w[0] = uint32_t(r_Fsl.size());
                                d.w[1] = uint32_t(fm.size());
                                r_Fdir.push_back(d);
                                for (const PendingSlice *ps : fm) {
                                    Unit u;
                                    memset(&u, 0, sizeof(u));
                                    u.w[0] = ps->emask; u.w[1] = ps->table; u.w[2] = uint32_t(ps->oslot);
                                    r_Fsl.push_back(u);
                                }
                            }
                            Unit h0, h1;
                            memset(&h1, 0, sizeof(h1));
                            h0.w[0] = uint32_t(C_DIAGS) | (2u << 16);
                            h0.w[1] = fidx | ((gm.empty() ? 0u : 1u) << 16);
                            h0.w[2] = 0;
                            h0.w[3] = smask;
                            h1.w[0] = g_off;
                            r_ops.push_back(h0); r_ops.push_back(h1);
                            r_nops++;
                            pending = keep;
                            gm.clear(); fm.clear();
                        }
                        if (!gm.empty() || !fm.empty()) {
                        uint32_t emu = 0;
                        for (const PendingSlice *ps : gm) emu |= ps->emask;
                        for (const PendingSlice *ps : fm) emu |= ps->emask;
                        uint32_t um = 0;
                        for (int u = 0; u < 16; u++)
                            if ((emu >> (u << VS)) & (VS ? 3u : 1u)) um |= 1u << u;
                        if (__builtin_popcount(um) >= 12) um = 0xffffu;   // nearly full: the mask-free code path (factor 1 elsewhere)
                        uint32_t g_off = 0xffffffffu;
                        if (!gm.empty()) {
                            const int upe = 1 << VS;   // elements per 16-byte unit
                            std::vector<cd> G(size_t(16) * size_t(nthr_round) * size_t(upe), cd(1.0));   // [thread][unit][element]
                            for (const PendingSlice *ps : gm) {
                                for (int t = 0; t < nthr_round; t++) {
                                    const uint32_t base = thread_base(t);
                                    if ((base & ps->tmask) != ps->tmask) continue;
                                    uint32_t idx = 0;
                                    for (int f = 0; f < ps->nf; f++) idx |= host_field(base, ps->fields[f]);
                                    const cd z = ps->host[idx];
                                    for (int e = 0; e < N; e++)
                                        if ((ps->emask >> e) & 1u) G[size_t(t) * size_t(N) + size_t(e)] *= z;
                                }
                            }
                            if (VS && ((enc.tables.size() / enc.esz) & 1)) enc.push_table(std::vector<cd>(1, cd(1.0)));   // 16-byte alignment
                            g_off = enc.push_table(G);
                        }
                        uint32_t fidx = 0xffffu;
                        if (!fm.empty()) {
                            fidx = uint32_t(f_base) + uint32_t(r_Fdir.size());
                            Unit d;
                            memset(&d, 0, sizeof(d));
                            d.w[0] = uint32_t(r_Fsl.size());      // round-relative; rebased when the round is appended
                            d.w[1] = uint32_t(fm.size());
                            r_Fdir.push_back(d);
                            for (const PendingSlice *ps : fm) {
                                Unit u;
                                memset(&u, 0, sizeof(u));
                                u.w[0] = ps->emask; u.w[1] = ps->table; u.w[2] = uint32_t(ps->oslot);
                                r_Fsl.push_back(u);
                            }
                        }
                        Unit h0, h1;
                        memset(&h1, 0, sizeof(h1));
                        h0.w[0] = uint32_t(C_DIAGF) | (2u << 16);
                        h0.w[1] = fidx | ((gm.empty() ? 0u : 1u) << 16);
                        h0.w[2] = 0;
                        h0.w[3] = um;
                        h1.w[0] = g_off;
                        r_ops.push_back(h0); r_ops.push_back(h1);
                        r_nops++;
                        pending = keep;
                        }
                    }
                }
                std::vector<bool> done(pending.size(), false);
                for (size_t i = 0; i < pending.size(); i++) {
                    if (done[i]) continue;
                    std::vector<size_t> members;
                    bool allsign = true;
                    for (size_t j = i; j < pending.size(); j++) {
                        if (done[j] || pending[j].emask != pending[i].emask) continue;
                        done[j] = true;
                        members.push_back(j);
                        allsign = allsign && pending[j].sign;
                    }
                    // thread-only slices -> one factor per thread, computed here: a thread's tile
                    // position (hence its index into each of these tables) is the same in every tile
                    uint32_t g_off = 0xffffffffu;
                    bool parity = false;
                    uint32_t parity_mask = 0, parity_sign = 0;
                    {
                        std::vector<cd> G;
                        for (size_t j : members) {
                            const PendingSlice &ps = pending[j];
                            if (ps.cls != 0) continue;
                            if (G.empty()) G.assign(size_t(nthr_round), cd(1.0));
                            for (int t = 0; t < nthr_round; t++) {
                                uint32_t base = 0;
                                for (size_t kb = 0; kb < tq.size(); kb++) if ((t >> kb) & 1) base |= 1u << (tq[kb] + VS);
                                if ((base & ps.tmask) != ps.tmask) continue;
                                uint32_t idx = 0;
                                for (int f = 0; f < ps.nf; f++) idx |= host_field(base, ps.fields[f]);
                                G[size_t(t)] *= ps.host[idx];
                            }
                        }
                        // a +-1 factor that is the parity of some of the thread's position bits (any
                        // product of Z / CZ / CCZ-free sign gates: (-1)^(popcount(base & M)) up to a
                        // constant sign) needs no table: the kernel computes it from `base`
                        if (!G.empty() && allsign && !generic_only) {
                            uint32_t M = 0;
                            bool ok = true;
                            for (const cd &z : G) ok = ok && z.imag() == 0.0 && (z.real() == 1.0 || z.real() == -1.0);
                            for (size_t kb = 0; ok && kb < tq.size(); kb++)
                                if (G[size_t(1) << kb].real() != G[0].real()) M |= 1u << (tq[kb] + VS);
                            for (int t = 0; ok && t < nthr_round; t++) {
                                uint32_t base = 0;
                                for (size_t kb = 0; kb < tq.size(); kb++) if ((t >> kb) & 1) base |= 1u << (tq[kb] + VS);
                                const double want = G[0].real() * ((__builtin_popcount(base & M) & 1) ? -1.0 : 1.0);
                                ok = G[size_t(t)].real() == want;
                            }
                            if (ok) {
                                parity_mask = M;
                                parity_sign = G[0].real() < 0.0 ? 0x80000000u : 0u;
                                parity = true;
                                G.clear();
                            }
                        }
                        if (!G.empty()) g_off = enc.push_table(G);
                    }
                    // outer-only slices -> one factor per tile (up to six look-ups per entry)
                    uint32_t h_idx = 0xffffffffu;
                    std::vector<size_t> cross;
                    {
                        std::vector<size_t> outer;
                        for (size_t j : members) {
                            if (pending[j].cls == 1 && outer.size() < 6) outer.push_back(j);
                            else if (pending[j].cls != 0) cross.push_back(j);
                        }
                        if (!outer.empty()) {
                            Unit u[4];
                            memset(u, 0, sizeof(u));
                            u[0].w[0] = uint32_t(outer.size());
                            for (size_t t = 0; t < outer.size(); t++) {
                                u[1 + t / 2].w[(t & 1) * 2] = pending[outer[t]].table;
                                u[1 + t / 2].w[(t & 1) * 2 + 1] = uint32_t(pending[outer[t]].oslot);
                            }
                            h_idx = uint32_t(h_base) + uint32_t(r_H.size() / 4);
                            for (int q = 0; q < 4; q++) r_H.push_back(u[q]);
                        }
                    }
                    std::vector<Unit> payload;
                    for (size_t j : cross) {
                        const PendingSlice &ps = pending[j];
                        Unit d0, d1, d2;
                        memset(&d1, 0, sizeof(d1)); memset(&d2, 0, sizeof(d2));
                        d0.w[0] = ps.table;
                        d0.w[1] = uint32_t(ps.nf) | (uint32_t(ps.oslot) << 16);
                        d0.w[2] = ps.tmask;
                        d0.w[3] = ps.fields[0];
                        for (int f = 0; f < 4; f++) { d1.w[f] = ps.fields[1 + f]; d2.w[f] = ps.fields[5 + f]; }
                        payload.push_back(d0); payload.push_back(d1);
                        if (ps.nf > 5) payload.push_back(d2);
                    }
                    Unit h0, h1;
                    memset(&h1, 0, sizeof(h1));
                    h0.w[0] = uint32_t(C_PHASE) | (uint32_t(2 + payload.size()) << 16);
                    h0.w[1] = uint32_t(cross.size()) | (select_code(pending[i].emask) << 16);
                    h0.w[2] = (allsign ? 1u : 0u) | (parity ? 2u : 0u);
                    h0.w[3] = pending[i].emask;
                    h1.w[0] = g_off;
                    h1.w[1] = h_idx;
                    h1.w[2] = parity_mask;      // sign = parity_sign ^ parity(base & parity_mask)
                    h1.w[3] = parity_sign;
                    r_ops.push_back(h0); r_ops.push_back(h1);
                    r_ops.insert(r_ops.end(), payload.begin(), payload.end());
                    r_nops++;
                }
                pending.clear();
            };

            // Hadamard-like gates a * [[1, 1], [1, -1]] without controls: all but the last one of the
            // round run as add / subtract butterflies, the last one carries the product of the scales
            // (a global factor commutes with everything in between)
            auto hadamard_scale = [&](const qj_op_desc &od, double *a) -> bool {
                if (od.kind != QJ_OPK_DENSE1 || od.ncontrols != 0 || od.ntargets != 1) return false;
                if (od.data_offset < 0 || od.data_offset + 4 > ndata) return false;
                const cd m0 = enc.data_at(od.data_offset), m1 = enc.data_at(od.data_offset + 1),
                         m2 = enc.data_at(od.data_offset + 2), m3 = enc.data_at(od.data_offset + 3);
                if (m0.imag() != 0.0 || m1.imag() != 0.0 || m2.imag() != 0.0 || m3.imag() != 0.0) return false;
                if (m0.real() == 0.0 || m1.real() != m0.real() || m2.real() != m0.real() || m3.real() != -m0.real()) return false;
                *a = m0.real();
                return true;
            };
            if (hadamard_on && ri == pd.first_round) {      // counted over the whole pass: ONE gate carries the scale
                h_left = 0;
                round_scale = 1.0;
                for (int64_t rj = pd.first_round; rj < pd.first_round + pd.nrounds; rj++)
                    for (int64_t oi = rounds_in[rj].first_op; oi < rounds_in[rj].first_op + rounds_in[rj].nops; oi++) {
                        double a;
                        if (oi >= 0 && oi < nops && hadamard_scale(ops[oi], &a)) h_left++;
                    }
            }

            for (int64_t oi = rdesc.first_op; oi < rdesc.first_op + rdesc.nops; oi++) {
                const qj_op_desc &od = ops[oi];
                if (od.ncontrols < 0 || od.ncontrols > QJ_MAX_QUBITS) return bail("op: bad control count");
                uint64_t seen = 0;
                uint32_t lcmask = 0;          // local control positions
                HostOuter ho;
                for (int c = 0; c < od.ncontrols; c++) {
                    const int b = od.controls[c];
                    if (b < 0 || b >= nqubits) return bail("op: control bit out of range");
                    if ((seen >> b) & 1) return bail("op: duplicate qubit");
                    seen |= uint64_t(1) << b;
                    if (lpos[b] >= 0) lcmask |= 1u << lpos[b];
                    else ho.ocmask |= uint64_t(1) << b;
                }
                const uint32_t tmask = lcmask & ~regmask;
                const uint32_t rcmask = lcmask & regmask;
                uint32_t cmask_e = 0;         // elements that satisfy the register-slot controls
                for (int e = 0; e < N; e++) if ((elem_bits(e) & rcmask) == rcmask) cmask_e |= 1u << e;

                if (od.kind == QJ_OPK_DENSE1 || od.kind == QJ_OPK_DENSE2) {
                    flush_pending();
                    const int nt = (od.kind == QJ_OPK_DENSE1) ? 1 : 2;
                    if (od.ntargets != nt) return bail("op: dense target count mismatch");
                    const int64_t need = (nt == 1) ? 4 : 16;
                    if (od.data_offset < 0 || od.data_offset + need > ndata) return bail("op: matrix outside the data array");
                    int sl[2] = {0, 0};
                    for (int t = 0; t < nt; t++) {
                        const int b = od.targets[t];
                        if (b < 0 || b >= nqubits) return bail("op: target bit out of range");
                        if ((seen >> b) & 1) return bail("op: duplicate qubit");
                        seen |= uint64_t(1) << b;
                        if (lpos[b] < 0) return bail("op: dense target is not a local bit of its pass");
                        if (slot_of_pos[lpos[b]] < 0) return bail("op: dense target is not a register bit of its round");
                        sl[t] = slot_of_pos[lpos[b]];
                    }
                    const int oslot = add_outer(ho);
                    std::vector<cd> m(need);
                    for (int64_t i = 0; i < need; i++) m[i] = enc.data_at(od.data_offset + i);
                    std::vector<Unit> payload;
                    double h_scale = 0.0;
                    if (nt == 1 && h_left > 0 && hadamard_scale(od, &h_scale)) {
                        h_left--;
                        if (h_left > 0) {                 // not the last one: butterfly, scale deferred
                            round_scale *= h_scale;
                            const size_t at = r_ops.size();
                            push_op(C_GROUP1H, oslot, 0, tmask, cmask_e, 0, nullptr, payload);
                            r_ops[at].w[1] |= (1u << sl[0]) << 16;
                            group_run.clear();
                            continue;
                        }
                        for (cd &z : m) z *= round_scale;  // the last one: a real 2x2 with the accumulated scale
                        round_scale = 1.0;
                    }
                    if (nt == 1) {
                        bool real = true;
                        for (const cd &z : m) real = real && z.imag() == 0.0;
                        const bool is_x = real && m[0] == 0.0 && m[1] == 1.0 && m[2] == 1.0 && m[3] == 0.0;
                        if (is_x) {
                            push_op(C_PERM1 + sl[0], oslot, 0, tmask, cmask_e, 0, nullptr, payload);
                            group_run.clear();
                        } else {
                            uint32_t code;
                            if (real) {
                                code = C_GROUP1R;
                                enc.push_scalars(payload, {m[0].real(), m[1].real(), m[2].real(), m[3].real()});
                            } else if (m[0].imag() == 0.0 && m[3].imag() == 0.0 && m[1].real() == 0.0 && m[2].real() == 0.0) {
                                code = C_GROUP1X;
                                if (dtype == QJ_C128)
                                    enc.push_scalars(payload, {m[0].real(), m[1].imag(), m[2].imag(), m[3].real()});
                                else   // packed FP32x2: {a, d, -b, b, -c, c, 0, 0}
                                    enc.push_scalars(payload, {m[0].real(), m[3].real(), -m[1].imag(), m[1].imag(),
                                                               -m[2].imag(), m[2].imag(), 0.0, 0.0});
                            } else {
                                code = C_GROUP1C;
                                enc.push_complex(payload, m);
                            }
                            const uint32_t full = (N == 32) ? 0xffffffffu : 0xffffu;
                            const bool plain = cmask_e == full && tmask == 0 && oslot == 0xffff;
                            // plain gates on different slots commute: merge into the nearest earlier
                            // plain group of the same kind, unless a group in between holds this slot
                            bool merged = false;
                            if (plain) {
                                for (size_t gi = group_run.size(); gi-- > 0 && !merged;) {
                                    const size_t at = group_run[gi];
                                    const uint32_t gslots = r_ops[at].w[1] >> 16;
                                    if ((gslots >> sl[0]) & 1u) break;
                                    if ((r_ops[at].w[0] & 0xffffu) != code) continue;
                                    uint32_t below = 0;
                                    for (int q = 0; q < sl[0]; q++) below += (gslots >> q) & 1u;
                                    const size_t pos = at + 2 + size_t(below) * payload.size();
                                    r_ops.insert(r_ops.begin() + pos, payload.begin(), payload.end());
                                    r_ops[at].w[0] += uint32_t(payload.size()) << 16;
                                    r_ops[at].w[1] |= (1u << sl[0]) << 16;
                                    for (size_t gj = gi + 1; gj < group_run.size(); gj++) group_run[gj] += payload.size();
                                    merged = true;
                                }
                            } else {
                                group_run.clear();
                            }
                            if (!merged) {
                                const size_t at = r_ops.size();
                                push_op(code, oslot, 0, tmask, cmask_e, 0, nullptr, payload);
                                r_ops[at].w[1] |= (1u << sl[0]) << 16;
                                if (plain) group_run.push_back(at);
                            }
                        }
                    } else {
                        group_run.clear();
                        int a = sl[0], b = sl[1];
                        if (a > b) {  // canonical slot order: exchange the matrix-index bits
                            std::vector<cd> t(16);
                            auto sw = [](int i) { return ((i & 1) << 1) | (i >> 1); };
                            for (int i = 0; i < 4; i++)
                                for (int j = 0; j < 4; j++) t[sw(i) * 4 + sw(j)] = m[i * 4 + j];
                            m = t;
                            std::swap(a, b);
                        }
                        static const double swap_m[16] = {1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1};
                        bool is_swap = true;
                        for (int i = 0; i < 16; i++) is_swap = is_swap && m[i] == cd(swap_m[i]);
                        if (is_swap) {
                            push_op(C_PERM2 + pair_index(a, b), oslot, 0, tmask, cmask_e, 0, nullptr, payload);
                        } else {
                            bool real4 = true;
                            for (const cd &z : m) real4 = real4 && z.imag() == 0.0;
                            if (real4) {
                                std::vector<double> sc;
                                for (const cd &z : m) sc.push_back(z.real());
                                enc.push_scalars(payload, sc);
                                push_op(C_DENSE2R + pair_index(a, b), oslot, 0, tmask, cmask_e, 0, nullptr, payload);
                            } else {
                                enc.push_complex(payload, m);
                                push_op(C_DENSE2 + pair_index(a, b), oslot, 0, tmask, cmask_e, 0, nullptr, payload);
                            }
                        }
                    }
                } else if (od.kind == QJ_OPK_DIAG) {
                    const int nb = od.ntargets;
                    if (nb < 0 || nb > QJ_MAX_DIAG_BITS) return bail("op: diagonal table over too many bits");
                    if (od.data_offset < 0 || od.data_offset + (int64_t(1) << nb) > ndata) return bail("op: table outside the data array");
                    // classify the table bits
                    std::vector<int> regj, regslot;        // table bit -> register slot
                    std::vector<std::pair<int, int>> thr;  // (local position, table bit)
                    std::vector<std::pair<int, int>> outb; // (index bit, table bit)
                    for (int j = 0; j < nb; j++) {
                        const int b = od.targets[j];
                        if (b < 0 || b >= nqubits) return bail("op: table bit out of range");
                        if ((seen >> b) & 1) return bail("op: duplicate qubit");
                        seen |= uint64_t(1) << b;
                        if (lpos[b] < 0) outb.push_back({b, j});
                        else if (slot_of_pos[lpos[b]] >= 0) { regj.push_back(j); regslot.push_back(slot_of_pos[lpos[b]]); }
                        else thr.push_back({lpos[b], j});
                    }
                    std::sort(thr.begin(), thr.end());
                    std::sort(outb.begin(), outb.end());
                    const int nthr = int(thr.size()), nout = int(outb.size()), k = int(regj.size());
                    if (nout > 11) return bail("op: diagonal table over more than 11 bits outside the tile");
                    // thread fields: runs of consecutive local positions
                    uint32_t fields[10] = {0};
                    int nf = 0;
                    for (int i = 0; i < nthr; i++) {
                        if (nf > 0) {
                            const uint32_t fl = fields[nf - 1];
                            const int src = fl & 255, len = (fl >> 8) & 255;
                            if (src + len == thr[i].first) { fields[nf - 1] = fl + (1u << 8); continue; }
                        }
                        if (nf == 9) return bail("op: diagonal table needs too many bit fields");
                        fields[nf++] = uint32_t(thr[i].first) | (1u << 8) | (uint32_t(i) << 16);
                    }
                    for (int i = 0; i < nout; i++) {
                        ho.src[i] = uint8_t(outb[i].first);
                        ho.dst[i] = uint8_t(nthr + i);
                    }
                    ho.nbits = nout;
                    const int nsub = nthr + nout;
                    auto sub_index = [&](int s, int a) {   // sub-table index s, register assignment a -> host table index
                        int idx = 0;
                        for (int i = 0; i < nthr; i++) idx |= ((s >> i) & 1) << thr[i].second;
                        for (int i = 0; i < nout; i++) idx |= ((s >> (nthr + i)) & 1) << outb[i].second;
                        for (int i = 0; i < k; i++) idx |= ((a >> i) & 1) << regj[i];
                        return idx;
                    };
                    // slices along the register bits
                    struct Slice { int a; std::vector<cd> t; bool sign; };
                    std::vector<Slice> slices;
                    for (int a = 0; a < (1 << k); a++) {
                        Slice sl;
                        sl.a = a;
                        sl.t.resize(size_t(1) << nsub);
                        bool ident = true, sign = true;
                        for (int s = 0; s < (1 << nsub); s++) {
                            const cd z = enc.data_at(od.data_offset + sub_index(s, a));
                            sl.t[s] = z;
                            ident = ident && z == cd(1.0);
                            sign = sign && z.imag() == 0.0 && (z.real() == 1.0 || z.real() == -1.0);
                        }
                        sl.sign = sign;
                        if (!ident) slices.push_back(sl);
                    }
                    if (slices.empty()) continue;
                    const bool use_slices = k <= 2 || slices.size() <= 4;
                    const int oslot = add_outer(ho);
                    std::vector<Unit> payload;
                    if (use_slices) {
                        for (const Slice &sl : slices) {
                            uint32_t emask = 0;
                            for (int e = 0; e < N; e++) {
                                bool ok = (cmask_e >> e) & 1u;
                                for (int i = 0; i < k && ok; i++) ok = ((e >> regslot[i]) & 1) == ((sl.a >> i) & 1);
                                if (ok) emask |= 1u << e;
                            }
                            if (emask == 0) continue;
                            PendingSlice ps;
                            ps.emask = emask; ps.tmask = tmask; ps.oslot = oslot; ps.nf = nf;
                            for (int f = 0; f < 9; f++) ps.fields[f] = (f < nf) ? fields[f] : 0u;
                            ps.sign = sl.sign;
                            ps.table = 0;
                            if (oslot == 0xffff) {                 // no outer bits, no outer controls
                                ps.cls = 0;
                                ps.host = sl.t;
                            } else {
                                ps.cls = (nf == 0 && tmask == 0) ? 1 : 2;
                                ps.table = enc.push_table(sl.t);
                            }
                            pending.push_back(ps);
                        }
                    } else {
                        // general table: index = thread fields | outer bits | register bits
                        flush_pending();
                        group_run.clear();
                        if (nf > 7) return bail("op: diagonal table needs too many bit fields");
                        std::vector<cd> t(size_t(1) << nb);
                        for (int a = 0; a < (1 << k); a++)
                            for (int s = 0; s < (1 << nsub); s++)
                                t[size_t(a) << nsub | s] = enc.data_at(od.data_offset + sub_index(s, a));
                        Unit uf, uw;
                        memset(&uf, 0, sizeof(uf)); memset(&uw, 0, sizeof(uw));
                        for (int i = 0; i < 4; i++) uf.w[i] = fields[3 + i];
                        uint16_t w[8] = {0};
                        for (int i = 0; i < k; i++) w[regslot[i]] = uint16_t(1u << (nsub + i));
                        uw.w[0] = w[0] | (uint32_t(w[1]) << 16); uw.w[1] = w[2] | (uint32_t(w[3]) << 16); uw.w[2] = w[4];
                        payload.push_back(uf); payload.push_back(uw);
                        push_op(C_DIAGN, oslot, nf, tmask, cmask_e, enc.push_table(t), fields, payload);
                    }
                } else {
                    return bail("op: unknown kind");
                }
            }

            flush_pending();

            // append the round (closing the launch first when the image would overflow)
            const size_t f_units = f_dir.size() + f_slices.size() + r_Fdir.size() + r_Fsl.size();
            const size_t need_units = 2 + round_units.size() + 3 + outer_units.size() + r_outer.size() + h_units.size() +
                                      r_H.size() + op_units.size() + r_ops.size() + (h_units.size() + r_H.size()) / 4 +
                                      f_units + 16 * (f_dir.size() + r_Fdir.size());
            if (3 + r_outer.size() + r_ops.size() + r_H.size() + r_H.size() / 4 + 2 + 17 * r_Fdir.size() + r_Fsl.size() >
                    size_t(kMaxBlobUnits) ||
                r_outer.size() / 2 > size_t(kMaxOuter))
                return bail("round: too many ops for one round");
            if (need_units > size_t(kMaxBlobUnits) || (outer_units.size() + r_outer.size()) / 2 > size_t(kMaxOuter)) {
                // the outer slots / factor indices of this round were numbered relative to the open
                // launch: renumber
                close_launch();
                for (size_t i = 0; i < r_ops.size();) {
                    const uint32_t units = r_ops[i].w[0] >> 16;
                    if ((r_ops[i].w[0] & 0xffffu) == uint32_t(C_PHASE)) {  // slots live in the table descriptors
                        const int ntab = int(r_ops[i].w[1] & 0xffffu);
                        if (r_ops[i + 1].w[1] != 0xffffffffu) r_ops[i + 1].w[1] -= uint32_t(h_base);
                        size_t d = i + 2;
                        for (int t = 0; t < ntab; t++) {
                            const uint32_t nf = r_ops[d].w[1] & 0xffffu, oslot = r_ops[d].w[1] >> 16;
                            if (oslot != 0xffffu) r_ops[d].w[1] = nf | ((oslot - uint32_t(outer_base)) << 16);
                            d += (nf > 5) ? 3 : 2;
                        }
                    } else if ((r_ops[i].w[0] & 0xffffu) == uint32_t(C_DIAGF) ||
                               (r_ops[i].w[0] & 0xffffu) == uint32_t(C_DIAGS)) {   // index of its per-tile factors
                        const uint32_t fidx = r_ops[i].w[1] & 0xffffu;
                        if (fidx != 0xffffu) r_ops[i].w[1] = (r_ops[i].w[1] & 0xffff0000u) | (fidx - uint32_t(f_base));
                    } else {
                        const uint32_t oslot = r_ops[i].w[1] & 0xffffu;
                        if (oslot != 0xffffu) r_ops[i].w[1] = (r_ops[i].w[1] & 0xffff0000u) | (oslot - uint32_t(outer_base));
                    }
                    i += units;
                }
                for (Unit &u : r_Fsl) u.w[2] -= uint32_t(outer_base);
                for (size_t e = 0; e < r_H.size(); e += 4) {
                    const uint32_t ntab = r_H[e].w[0];
                    for (uint32_t t = 0; t < ntab; t++) r_H[e + 1 + t / 2].w[(t & 1) * 2 + 1] -= uint32_t(outer_base);
                }
            }
            if (r_nops > 0) {
                Unit u0, u1, u2;
                memset(&u0, 0, sizeof(u0)); memset(&u1, 0, sizeof(u1)); memset(&u2, 0, sizeof(u2));
                u0.w[0] = uint32_t(op_units.size());   // relative to the op stream; rebased in close_launch
                u0.w[1] = uint32_t(r_ops.size());   // units of the round's op stream
                u0.w[2] = vd[0] | (vd[1] << 16); u0.w[3] = vd[2] | (vd[3] << 16);
                for (size_t kbit = 0; kbit < tq.size() && kbit < 8; kbit++) {
                    const uint32_t td = swz_vec(1u << tq[kbit]) << 4;
                    u1.w[kbit >> 1] |= td << ((kbit & 1) * 16);
                    u2.w[kbit >> 2] |= uint32_t(tq[kbit] + VS) << ((kbit & 3) * 8);
                }
                round_units.push_back(u0); round_units.push_back(u1); round_units.push_back(u2);
                outer_units.insert(outer_units.end(), r_outer.begin(), r_outer.end());
                h_units.insert(h_units.end(), r_H.begin(), r_H.end());
                for (Unit &d : r_Fdir) d.w[0] += uint32_t(f_slices.size());
                f_dir.insert(f_dir.end(), r_Fdir.begin(), r_Fdir.end());
                f_slices.insert(f_slices.end(), r_Fsl.begin(), r_Fsl.end());
                op_units.insert(op_units.end(), r_ops.begin(), r_ops.end());
                launch_rounds++;
                launch_ops += r_nops;
                prog->total_rounds++;
                prog->total_mops += r_nops;
            }
        }
        close_launch();
    }

    if (image) {
        image->dtype = dtype;
        image->nqubits = nqubits;
        image->blob = blob_all;
        image->tables = enc.tables;
        image->launches = prog->launches;
        delete prog;
        return QJ_OK;
    }
    qj::DeviceGuard device_guard(h);
    auto upload = [&](void **dst, const void *src, size_t bytes) -> cudaError_t {
        if (bytes == 0) bytes = 16;   // never hand a null table pointer to the kernel
        cudaError_t e = cudaMalloc(dst, bytes);
        if (e != cudaSuccess || src == nullptr) return e;
        return cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, h->stream);
    };
    prog->images.resize(prog->launches.size());
    for (size_t li = 0; li < prog->launches.size(); li++) {
        const qj_program::Launch &L = prog->launches[li];
        memset(&prog->images[li], 0, sizeof(ProgParam));
        memcpy(prog->images[li].u, blob_all.data() + L.blob_off, size_t(L.geom.blob_units) * 16);
    }
    cudaError_t e = upload(&prog->d_tables, enc.tables.empty() ? nullptr : enc.tables.data(), enc.tables.size());
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);  // host vectors die with this frame
    if (e != cudaSuccess) {
        qj_program_destroy(h, prog);
        return fail(QJ_ERR_CUDA, std::string("program upload: ") + cudaGetErrorString(e));
    }
    *out = prog;
    return QJ_OK;
}

extern "C" int qj_program_create(qj_handle *h, int dtype, int nqubits, const qj_pass_desc *passes,
                                 int npasses, const qj_round_desc *rounds_in, int64_t nrounds_in,
                                 const qj_op_desc *ops, int64_t nops, const void *data, int64_t ndata,
                                 qj_program **out) {
    qj::DeviceGuard device_guard(h);
    QJ_REQUIRE(h && out, "null argument");
    return program_build(h, dtype, nqubits, passes, npasses, rounds_in, nrounds_in, ops, nops, data, ndata, out, nullptr);
}

extern "C" int qj_program_encode(int dtype, int nqubits, const qj_pass_desc *passes, int npasses,
                                 const qj_round_desc *rounds_in, int64_t nrounds_in, const qj_op_desc *ops,
                                 int64_t nops, const void *data, int64_t ndata, qj_program_image **out) {
    QJ_REQUIRE(out != nullptr, "null argument");
    auto *img = new qj_program_image();
    const int rc = program_build(nullptr, dtype, nqubits, passes, npasses, rounds_in, nrounds_in, ops, nops, data,
                                 ndata, nullptr, img);
    if (rc != QJ_OK) {
        delete img;
        return rc;
    }
    *out = img;
    return QJ_OK;
}

extern "C" int qj_program_image_sizes(const qj_program_image *img, int64_t *nlaunches, int64_t *blob_bytes,
                                      int64_t *table_bytes) {
    QJ_REQUIRE(img != nullptr, "null image");
    if (nlaunches) *nlaunches = int64_t(img->launches.size());
    if (blob_bytes) *blob_bytes = int64_t(img->blob.size()) * 16;
    if (table_bytes) *table_bytes = int64_t(img->tables.size());
    return QJ_OK;
}

extern "C" int qj_program_image_read(const qj_program_image *img, void *blob_out, void *tables_out,
                                     int64_t *launch_info) {
    QJ_REQUIRE(img != nullptr, "null image");
    if (blob_out && !img->blob.empty()) memcpy(blob_out, img->blob.data(), img->blob.size() * 16);
    if (tables_out && !img->tables.empty()) memcpy(tables_out, img->tables.data(), img->tables.size());
    if (launch_info) {
        for (size_t i = 0; i < img->launches.size(); i++) {
            const qj_program::Launch &L = img->launches[i];
            int64_t *o = launch_info + QJ_LAUNCH_INFO_FIELDS * i;
            o[0] = L.geom.T; o[1] = L.geom.r; o[2] = L.geom.nh; o[3] = L.geom.ntiles;
            o[4] = L.blob_off; o[5] = L.geom.blob_units; o[6] = L.geom.nH; o[7] = int64_t(L.smem);
            for (int b = 0; b < kMaxHiBits; b++) o[8 + b] = L.geom.hibit[b];
        }
    }
    return QJ_OK;
}

extern "C" int qj_program_image_destroy(qj_program_image *img) {
    delete img;
    return QJ_OK;
}

extern "C" int qj_program_destroy(qj_handle *h, qj_program *p) {
    qj::DeviceGuard device_guard(h);
    if (!p) return QJ_OK;
    if (h) cudaStreamSynchronize(h->stream);   // (without a handle cudaFree synchronises the device by itself)
    cudaFree(p->d_tables);
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
int launch_pass(qj_handle *h, const qj_program *p, void *state, int li, int zero_input, int64_t tile_begin = 0,
                int64_t tile_count = -1, void *out = nullptr, bool leave_a_slot = false) {
    const qj_program::Launch &L = p->launches[li];
    if (tile_count < 0) tile_count = L.geom.ntiles - tile_begin;
    QJ_REQUIRE(tile_begin >= 0 && tile_count >= 0 && tile_begin + tile_count <= L.geom.ntiles, "tile range out of bounds");
    if (tile_count == 0) return QJ_OK;
    // one thread per 16 vectors of the tile (at most 256); 128 registers per thread allow 512
    // resident threads per SM: two 64 KiB tiles or four 32 KiB tiles in different phases
    const int VS = (sizeof(T) == 8) ? 0 : 1;
    const int threads = std::max(1, std::min(kThreads, (1 << (L.geom.T - VS)) >> kVecRegBits));
    const int by_smem = (int)std::max<size_t>(1, (size_t(224) << 10) / (L.smem + 1024));
    const int per_sm = std::max(1, std::min(by_smem, 512 / threads));
    // The CTAs of a launch live until its last tile (grid-stride loop) and two or four of them hold
    // every register of an SM.  A pipelined sub-block launch therefore starts one CTA short of the
    // device's capacity: the free slot is where the high-priority handshake kernel of the exchange
    // runs while the pass is in flight (with a full grid it would wait for the launch to drain).
    const int64_t slots = int64_t(h->sm_count) * per_sm - (leave_a_slot ? 1 : 0);
    const unsigned grid = (unsigned)std::min<int64_t>(tile_count, std::max<int64_t>(1, slots));
    PassGeom geom = L.geom;
    geom.zero_input = zero_input;
    geom.tile_begin = tile_begin;
    geom.tile_end = tile_begin + tile_count;
    // (a byte difference of two 16-byte aligned device addresses; applied to the vector pointer in the kernel)
    geom.out_off = out ? (reinterpret_cast<intptr_t>(out) - reinterpret_cast<intptr_t>(state)) / 16 : 0;
    if constexpr (sizeof(T) == 4) {
        // (the complex64 instantiation lives in pass_kernels_f32.cu)
        const int rc = launch_k_pass_f32(h->device, grid, threads, L.smem, h->stream, state, &geom, p->d_tables,
                                         &p->images[li]);
        if (rc) return rc;
    } else {
        static bool configured[kMaxDevices] = {false};   // the attribute is per device, not per process
        if (!configured[h->device]) {
            QJ_CUDA_OK(cudaFuncSetAttribute(k_pass<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10));
            configured[h->device] = true;
        }
        k_pass<T><<<grid, threads, L.smem, h->stream>>>(
            reinterpret_cast<Cx<T> *>(state), geom, reinterpret_cast<const Cx<T> *>(p->d_tables), p->images[li]);
    }
    h->launches++;
    QJ_CUDA_OK(cudaGetLastError());
    return QJ_OK;
}

int run_launches(qj_handle *h, const qj_program *p, void *state, int first, int count, int flags = 0) {
    for (int li = first; li < first + count; li++) {
        const int zero = (li == first) ? (flags & 1) : 0;
        const int rc = (p->dtype == QJ_C128) ? launch_pass<double>(h, p, state, li, zero) : launch_pass<float>(h, p, state, li, zero);
        if (rc) return rc;
    }
    return QJ_OK;
}
}  // namespace

extern "C" int qj_program_run(qj_handle *h, const qj_program *p, void *state) {
    qj::DeviceGuard device_guard(h);
    QJ_REQUIRE(h && p && state, "null argument");
    QJ_REQUIRE((reinterpret_cast<uintptr_t>(state) & 15) == 0, "state must be 16-byte aligned");
    return run_launches(h, p, state, 0, (int)p->launches.size());
}

extern "C" int qj_program_run_ex(qj_handle *h, const qj_program *p, void *state, int first_launch, int nlaunches,
                                 int flags) {
    qj::DeviceGuard device_guard(h);
    QJ_REQUIRE(h && p && state, "null argument");
    QJ_REQUIRE((reinterpret_cast<uintptr_t>(state) & 15) == 0, "state must be 16-byte aligned");
    if (nlaunches < 0) nlaunches = (int)p->launches.size() - first_launch;
    QJ_REQUIRE(first_launch >= 0 && nlaunches >= 0 && first_launch + nlaunches <= (int)p->launches.size(),
               "launch range out of bounds");
    QJ_REQUIRE((flags & ~QJ_RUN_ZERO_INPUT) == 0, "unknown flag");
    if (nlaunches == 0) {
        // an empty program on |0...0> still has to prepare the state
        if (flags & QJ_RUN_ZERO_INPUT) return qj_initial_state(h, state, p->dtype, p->nqubits);
        return QJ_OK;
    }
    return run_launches(h, p, state, first_launch, nlaunches, flags);
}

extern "C" int qj_program_run_tiles(qj_handle *h, const qj_program *p, void *state, int launch, int64_t tile_begin,
                                    int64_t tile_count) {
    qj::DeviceGuard device_guard(h);
    QJ_REQUIRE(h && p && state, "null argument");
    QJ_REQUIRE(launch >= 0 && launch < (int)p->launches.size(), "launch index out of range");
    return (p->dtype == QJ_C128) ? launch_pass<double>(h, p, state, launch, 0, tile_begin, tile_count, nullptr, true)
                                 : launch_pass<float>(h, p, state, launch, 0, tile_begin, tile_count, nullptr, true);
}

extern "C" int qj_program_run_tiles_to(qj_handle *h, const qj_program *p, const void *state, void *out, int launch,
                                       int64_t tile_begin, int64_t tile_count) {
    qj::DeviceGuard device_guard(h);
    QJ_REQUIRE(h && p && state && out, "null argument");
    QJ_REQUIRE(((reinterpret_cast<uintptr_t>(state) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
               "state and out must be 16-byte aligned");
    QJ_REQUIRE(launch >= 0 && launch < (int)p->launches.size(), "launch index out of range");
    void *src = const_cast<void *>(state);
    return (p->dtype == QJ_C128) ? launch_pass<double>(h, p, src, launch, 0, tile_begin, tile_count, out, true)
                                 : launch_pass<float>(h, p, src, launch, 0, tile_begin, tile_count, out, true);
}

extern "C" int qj_program_launch_geometry(const qj_program *p, int launch, int64_t *out) {
    QJ_REQUIRE(p && out, "null argument");
    QJ_REQUIRE(launch >= 0 && launch < (int)p->launches.size(), "launch index out of range");
    const PassGeom &g = p->launches[launch].geom;
    out[0] = g.T; out[1] = g.r; out[2] = g.nh; out[3] = g.ntiles;
    for (int b = 0; b < kMaxHiBits; b++) out[4 + b] = (b < g.nh) ? g.hibit[b] : -1;
    return QJ_OK;
}

extern "C" int qj_program_run_launch(qj_handle *h, const qj_program *p, void *state, int launch) {
    qj::DeviceGuard device_guard(h);
    QJ_REQUIRE(h && p && state, "null argument");
    QJ_REQUIRE(launch >= 0 && launch < (int)p->launches.size(), "launch index out of range");
    return run_launches(h, p, state, launch, 1);
}
