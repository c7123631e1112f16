"""TEST INFRASTRUCTURE: numpy interpreter of the encoded pass-program image
(`qj_program_encode`, qibojit_b200/csrc/pass_kernels.cu).

It follows the kernel `k_pass` step by step -- tile geometry, the thread -> amplitude map of
every round, op headers, payload formats (including the packed FP32x2 operand pairs of complex64),
per-thread / per-tile phase factors, outer predicates -- so that the HOST side of the pass kernel
(planner + C++ encoder) can be checked on a machine without a GPU: a random circuit encoded and
interpreted here must give the state the CPU oracle gives.  The arithmetic is done in complex128
whatever the program dtype (the image's constants carry the program precision).
"""

import ctypes

import numpy as np

from qibojit_b200 import _capi, planner

C_GROUP1C, C_GROUP1R, C_GROUP1X, C_PERM1, C_DENSE2, C_PERM2, C_PHASE, C_DIAGN = 0, 1, 2, 3, 8, 18, 28, 29
C_DENSE2R, C_DIAGF, C_DIAGC, C_DIAGS, C_GROUP1H, C_DIAGCS = 30, 40, 41, 42, 43, 44
SEL_ALL, SEL_SLOT, SEL_PAIR, SEL_MASK = 0, 1, 6, 16
PAIRS = [(0, 1), (0, 2), (1, 2), (0, 3), (1, 3), (2, 3), (0, 4), (1, 4), (2, 4), (3, 4)]


def swz_vec(v):
    return v ^ (((v >> 3) ^ (v >> 6) ^ (v >> 9)) & 7)


def field_of(base, fl):
    return ((base >> (fl & 255)) & ((1 << ((fl >> 8) & 255)) - 1)) << (fl >> 16)


class _EncodeLib:
    """Stands in for the ctypes library inside planner.Program: `qj_program_create` becomes the
    device-free `qj_program_encode` and the image is kept instead of a device program."""

    def __init__(self):
        self.lib = _capi.load()
        self.images = []

    def qj_program_create(self, handle, dtype, nqubits, passes, npasses, rounds, nrounds, ops, nops, data, ndata, out):
        img = ctypes.c_void_p()
        rc = self.lib.qj_program_encode(dtype, nqubits, passes, npasses, rounds, nrounds, ops, nops, data, ndata,
                                        ctypes.byref(img))
        if rc == 0:
            self.images.append((img, dtype, nqubits))
            out._obj.value = len(self.images)        # a non-null fake handle
        return rc

    def qj_program_destroy(self, handle, prog):
        return 0

    def __getattr__(self, name):
        return getattr(self.lib, name)


class EncoderBackend:
    """Just enough of B200Backend for planner.Program to plan, serialise and encode."""

    def __init__(self, dtype):
        from qibojit_b200.matrices import CustomMatrices

        self.dtype = dtype
        self.custom_matrices = CustomMatrices(dtype)
        self._lib = _EncodeLib()

    def _handle(self):
        return None

    def _handle_or_none(self):
        return None


class Image:
    def __init__(self, lib, img, dtype):
        nl, bb, tb = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        _capi.check(lib.qj_program_image_sizes(img, ctypes.byref(nl), ctypes.byref(bb), ctypes.byref(tb)))
        self.units = np.zeros((bb.value // 16, 4), dtype=np.uint32)
        raw_tables = np.zeros(max(tb.value, 1), dtype=np.uint8)
        info = np.zeros((max(nl.value, 1), _capi.QJ_LAUNCH_INFO_FIELDS), dtype=np.int64)
        _capi.check(lib.qj_program_image_read(img, self.units.ctypes.data, raw_tables.ctypes.data, info.ctypes.data))
        self.c128 = dtype == _capi.QJ_C128
        self.tables = raw_tables[:tb.value].view(np.complex128 if self.c128 else np.complex64).astype(np.complex128)
        self.launches = [dict(T=int(r[0]), r=int(r[1]), nh=int(r[2]), ntiles=int(r[3]), blob_off=int(r[4]),
                              blob_units=int(r[5]), nH=int(r[6]), smem=int(r[7]), hibit=[int(x) for x in r[8:16]])
                         for r in info[:nl.value]]
        lib.qj_program_image_destroy(img)

    # payload readers ------------------------------------------------------------------------
    def scalars(self, units):
        """Units -> real scalars of the program precision."""
        raw = np.ascontiguousarray(units).view(np.uint8).reshape(-1)
        return raw.view(np.float64 if self.c128 else np.float32).astype(np.float64)

    def complex_elements(self, units, count):
        s = self.scalars(units)
        if self.c128:
            return s[0:2 * count:2] + 1j * s[1:2 * count:2]
        quad = s[:4 * count].reshape(count, 4)       # (gr, gr, -gi, gi)
        assert np.array_equal(quad[:, 0], quad[:, 1]) and np.array_equal(quad[:, 2], -quad[:, 3])
        return quad[:, 0] + 1j * quad[:, 3]


def run_image(image, state):
    """Apply every launch of `image` to `state` (complex128 numpy vector) in place."""
    VS = 0 if image.c128 else 1
    J = 4 if image.c128 else 5
    N = 1 << J
    for L in image.launches:
        U = image.units[L["blob_off"]:L["blob_off"] + L["blob_units"]]
        T, r, nh = L["T"], L["r"], L["nh"]
        assert T == r + nh
        Tv = T - VS
        nthr = max(1, (1 << Tv) >> 4)
        nrounds, nouter, off_rounds, off_outer = (int(x) for x in U[0])
        nH, off_H = int(U[1][0]), int(U[1][1])
        nF, off_F = int(U[1][2]), int(U[1][3])
        gbit = list(range(r)) + L["hibit"][:nh]            # local amplitude position -> index bit
        assert L["smem"] <= 200 << 10

        def outer_value(m, base_amp):
            o0, o1 = U[off_outer + 2 * m], U[off_outer + 2 * m + 1]
            ocmask = int(o0[0]) | (int(o0[1]) << 32)
            if (base_amp & ocmask) != ocmask:
                return -1
            srcw, dstw = [int(o0[3]), int(o1[0]), int(o1[1])], [int(o1[2]), int(o1[3])]
            v = 0
            for b in range(int(o0[2])):
                src = (srcw[b >> 2] >> ((b & 3) * 8)) & 255
                dst = (dstw[b >> 3] >> ((b & 7) * 4)) & 15
                v |= ((base_amp >> src) & 1) << dst
            return v

        free = [b for b in range(image_nqubits(state)) if b not in gbit]
        for tile_id in range(L["ntiles"]):
            base_amp = 0
            for i, b in enumerate(free):
                base_amp |= ((tile_id >> i) & 1) << b
            pos = np.arange(1 << T)
            gidx = np.full(1 << T, base_amp, dtype=np.int64)
            for p, b in enumerate(gbit):
                gidx |= ((pos >> p) & 1).astype(np.int64) << b
            tile = state[gidx].copy()
            s_outer = [outer_value(m, base_amp) for m in range(nouter)]
            s_H = []
            for m in range(nH):
                he = U[off_H + 4 * m:off_H + 4 * m + 4]
                acc = 1.0 + 0j
                for t in range(int(he[0][0])):
                    u = he[1 + (t >> 1)]
                    table, osl = (int(u[2]), int(u[3])) if t & 1 else (int(u[0]), int(u[1]))
                    v = outer_value(osl, base_amp)
                    if v >= 0:
                        acc *= image.tables[table + v]
                s_H.append(acc)
            s_F = np.ones((nF, N), dtype=np.complex128)     # per-tile, per-element factors of the fused diagonals
            for f in range(nF):
                first, nsl = int(U[off_F + f][0]), int(U[off_F + f][1])
                for sl in U[off_F + first:off_F + first + nsl]:
                    v = outer_value(int(sl[2]), base_amp)
                    if v < 0:
                        continue
                    for ee in range(N):
                        if (int(sl[0]) >> ee) & 1:
                            s_F[f, ee] *= image.tables[int(sl[1]) + v]

            tid = np.arange(nthr)
            for rd in range(nrounds):
                r0, r1, r2 = (U[off_rounds + 3 * rd + k] for k in range(3))
                vd = [int(r0[2]) & 0xffff, int(r0[2]) >> 16, int(r0[3]) & 0xffff, int(r0[3]) >> 16]
                # register slots -> local amplitude positions (the swizzle is an involution)
                rp = ([0] if VS else []) + [int(np.log2(swz_vec(v >> 4))) + VS for v in vd]
                nthread_bits = Tv - 4
                tdw = [int(x) for x in r1]
                tpos, td = [], []
                for k in range(nthread_bits):
                    td.append((tdw[k >> 1] >> ((k & 1) * 16)) & 0xffff)
                    tpos.append(((int(r2[0]) if k < 4 else int(r2[1])) >> ((k & 3) * 8)) & 255)
                    # the swizzled offset of a thread bit and the position used for predicates agree
                    assert swz_vec(td[k] >> 4) == 1 << (tpos[k] - VS)
                assert sorted(rp + tpos) == list(range(T)), "register + thread bits must cover the tile"
                base = np.zeros(nthr, dtype=np.int64)
                for k in range(nthread_bits):
                    base |= ((tid >> k) & 1) << tpos[k]
                e = np.arange(N)
                epos = np.zeros(N, dtype=np.int64)
                for j in range(J):
                    epos |= ((e >> j) & 1) << rp[j]
                where = base[:, None] | epos[None, :]            # (nthr, N) local positions
                assert len(np.unique(where)) == 1 << T
                x = tile[where]

                op = int(r0[0])
                op_end = op + int(r0[1])
                while op != op_end:
                    h0, h1 = U[op], U[op + 1]
                    pay = op + 2
                    units = int(h0[0]) >> 16
                    code = int(h0[0]) & 0xffff
                    op += units
                    emask = np.full(nthr, int(h0[3]), dtype=np.int64)
                    oi = 0
                    if (code not in (C_PHASE, C_DIAGF, C_DIAGC, C_DIAGS) and not C_DIAGCS <= code < C_DIAGCS + 5) and (int(h0[2]) != 0 or (int(h0[1]) & 0xffff) != 0xffff):
                        oslot, tmask = int(h0[1]) & 0xffff, int(h0[2])
                        ok = (base & tmask) == tmask
                        if oslot != 0xffff:
                            oi = s_outer[oslot]
                            ok &= oi >= 0
                            oi = max(oi, 0)
                        emask = np.where(ok, emask, 0)
                    sel_e = ((emask[:, None] >> e[None, :]) & 1).astype(bool)      # (nthr, N)

                    if code in (C_GROUP1C, C_GROUP1R, C_GROUP1X):
                        slots = int(h0[1]) >> 16
                        nu = {C_GROUP1C: 4, C_GROUP1R: 2 if image.c128 else 1, C_GROUP1X: 2}[code]
                        for A in range(J):
                            if not (slots >> A) & 1:
                                continue
                            u = U[pay:pay + nu]
                            pay += nu
                            if code == C_GROUP1C:
                                g = image.complex_elements(u, 4).reshape(2, 2)
                            elif code == C_GROUP1R:
                                g = image.scalars(u)[:4].reshape(2, 2).astype(np.complex128)
                            else:
                                s = image.scalars(u)
                                if image.c128:
                                    a, b, c, d = s[:4]
                                else:
                                    a, d, nb, b, nc, c = s[:6]
                                    assert nb == -b and nc == -c
                                g = np.array([[a, 1j * b], [1j * c, d]])
                            for e0 in range(N):
                                if (e0 >> A) & 1:
                                    continue
                                e1 = e0 | (1 << A)
                                m = sel_e[:, e0]
                                s0, s1 = x[m, e0].copy(), x[m, e1].copy()
                                x[m, e0] = g[0, 0] * s0 + g[0, 1] * s1
                                x[m, e1] = g[1, 0] * s0 + g[1, 1] * s1
                        assert pay == op
                    elif C_PERM1 <= code < C_PERM1 + 5:
                        A = code - C_PERM1
                        for e0 in range(N):
                            if (e0 >> A) & 1:
                                continue
                            e1 = e0 | (1 << A)
                            m = sel_e[:, e0]
                            x[m, e0], x[m, e1] = x[m, e1].copy(), x[m, e0].copy()
                    elif (C_DENSE2 <= code < C_DENSE2 + 10 or C_PERM2 <= code < C_PERM2 + 10
                          or C_DENSE2R <= code < C_DENSE2R + 10):
                        perm, real = C_PERM2 <= code < C_PERM2 + 10, code >= C_DENSE2R
                        a, b = PAIRS[code - (C_DENSE2R if real else C_PERM2 if perm else C_DENSE2)]
                        if perm:
                            g = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)
                        elif real:      # 16 real scalars, row-major
                            nu = 8 if image.c128 else 4
                            assert units == 2 + nu
                            g = image.scalars(U[pay:pay + nu])[:16].reshape(4, 4).astype(np.complex128)
                        else:
                            g = image.complex_elements(U[pay:pay + 16], 16).reshape(4, 4)
                        for e0 in range(N):
                            if (e0 >> a) & 1 or (e0 >> b) & 1:
                                continue
                            idx = [e0 | ((j & 1) << a) | ((j >> 1) << b) for j in range(4)]
                            m = sel_e[:, e0]
                            s = x[m][:, idx]
                            x[np.ix_(np.nonzero(m)[0], idx)] = s @ g.T
                    elif code == C_PHASE:
                        ntab, sel, allsign = int(h0[1]) & 0xffff, int(h0[1]) >> 16, int(h0[2]) & 1
                        parity = (int(h0[2]) >> 1) & 1
                        ph = np.ones(nthr, dtype=np.complex128)
                        if int(h1[0]) != 0xffffffff:
                            ph = ph * image.tables[int(h1[0]) + tid]
                        if int(h1[1]) != 0xffffffff:
                            ph = ph * s_H[int(h1[1])]
                        if parity:          # sign = constant sign x parity of some position bits of the thread
                            assert allsign and int(h1[0]) == 0xffffffff
                            par = np.array([bin(int(v) & int(h1[2])).count("1") & 1 for v in base])
                            ph = ph * np.where(par == 1, -1.0, 1.0) * (-1.0 if int(h1[3]) else 1.0)
                        d = pay
                        for _ in range(ntab):
                            d0, d1 = U[d], U[d + 1]
                            nf, osl = int(d0[1]) & 0xffff, int(d0[1]) >> 16
                            fields = [int(d0[3])] + [int(v) for v in d1]
                            if nf > 5:
                                fields += [int(v) for v in U[d + 2]]
                            d += 3 if nf > 5 else 2
                            ok = (base & int(d0[2])) == int(d0[2])
                            idx = np.zeros(nthr, dtype=np.int64)
                            if osl != 0xffff:
                                if s_outer[osl] < 0:
                                    ok &= False
                                else:
                                    idx |= s_outer[osl]
                            for f in fields[:nf]:
                                idx |= field_of(base, f)
                            z = image.tables[int(d0[0]) + np.where(ok, idx, 0)]
                            ph = ph * np.where(ok, z, 1.0)
                        assert d == op
                        if allsign:
                            assert np.all(ph.imag == 0) and np.all(np.abs(ph.real) == 1)
                        if sel == SEL_ALL:
                            chosen = np.ones(N, dtype=bool)
                        elif SEL_SLOT <= sel < SEL_SLOT + 5:
                            chosen = ((e >> (sel - SEL_SLOT)) & 1).astype(bool)
                        elif SEL_PAIR <= sel < SEL_PAIR + 10:
                            a, b = PAIRS[sel - SEL_PAIR]
                            chosen = (((e >> a) & 1) & ((e >> b) & 1)).astype(bool)
                        else:
                            chosen = ((int(h0[3]) >> e) & 1).astype(bool)
                        assert np.array_equal(chosen, ((int(h0[3]) >> e) & 1).astype(bool))
                        x[:, chosen] *= ph[:, None]
                    elif code == C_DIAGF:
                        fidx, has_g, um = int(h0[1]) & 0xffff, int(h0[1]) >> 16, int(h0[3])
                        upe = 1 << VS                               # elements per 16-byte unit
                        for u in range(16):
                            if not (um >> u) & 1:
                                continue
                            for k in range(upe):
                                ee = u * upe + k
                                z = np.ones(nthr, dtype=np.complex128)
                                if has_g:
                                    z = image.tables[int(h1[0]) + (tid * 16 + u) * upe + k]      # [thread][unit][element]
                                if fidx != 0xffff:
                                    z = z * s_F[fidx, ee]
                                x[:, ee] *= z
                        assert units == 2
                    elif code == C_GROUP1H:
                        slots = int(h0[1]) >> 16
                        assert units == 2 and np.all(sel_e)          # plain gates only
                        for A in range(J):
                            if not (slots >> A) & 1:
                                continue
                            for e0 in range(N):
                                if (e0 >> A) & 1:
                                    continue
                                e1 = e0 | (1 << A)
                                s0, s1 = x[:, e0].copy(), x[:, e1].copy()
                                x[:, e0] = s0 + s1                   # unnormalised: the scale rides on a later gate
                                x[:, e1] = s0 - s1
                    elif C_DIAGCS <= code < C_DIAGCS + 5:
                        A = code - C_DIAGCS
                        upe = 1 << VS
                        assert units == 2 + (N // 2) // upe
                        sc = image.scalars(U[pay:pay + (N // 2) // upe]).reshape(-1, 2)[:N // 2]
                        for p2 in range(N // 2):
                            ee = (((p2 >> A) << (A + 1)) | (p2 & ((1 << A) - 1))) | (1 << A)
                            x[:, ee] *= sc[p2, 0] + 1j * sc[p2, 1]
                    elif code == C_DIAGS:
                        fidx, has_g, smask = int(h0[1]) & 0xffff, int(h0[1]) >> 16, int(h0[3])
                        per = 8 if VS else 4                        # factors per thread in the table
                        for a in range(J):
                            if not (smask >> a) & 1:
                                continue
                            z = np.ones(nthr, dtype=np.complex128)
                            if has_g:
                                z = image.tables[int(h1[0]) + tid * per + a]
                            if fidx != 0xffff:
                                z = z * s_F[fidx, 1 << a]
                            x[:, ((e >> a) & 1).astype(bool)] *= z[:, None]
                        assert units == 2
                    elif code == C_DIAGC:
                        um = int(h0[3])
                        upe = 1 << VS
                        nun = bin(um).count("1")
                        assert units == 2 + nun
                        sc = image.scalars(U[pay:pay + nun]).reshape(nun, -1)     # c128: (re, im); c64: (re, im, re, im)
                        k = 0
                        for u in range(16):
                            if not (um >> u) & 1:
                                continue
                            for j in range(upe):
                                x[:, u * upe + j] *= sc[k, 2 * j] + 1j * sc[k, 2 * j + 1]
                            k += 1
                    elif code == C_DIAGN:
                        nf = int(h0[1]) >> 16
                        f4, w = U[pay], U[pay + 1]
                        fields = [int(h1[1]), int(h1[2]), int(h1[3])] + [int(v) for v in f4]
                        idxb = np.full(nthr, oi, dtype=np.int64)
                        for f in fields[:nf]:
                            idxb |= field_of(base, f)
                        wj = [int(w[0]) & 0xffff, int(w[0]) >> 16, int(w[1]) & 0xffff, int(w[1]) >> 16, int(w[2]) & 0xffff]
                        for ee in range(N):
                            idx = idxb.copy()
                            for j in range(J):
                                if (ee >> j) & 1:
                                    idx |= wj[j]
                            m = sel_e[:, ee]
                            x[m, ee] *= image.tables[int(h1[0]) + idx[m]]
                    else:
                        raise AssertionError(f"unknown op code {code}")
                tile[where] = x
            state[gidx] = tile
    return state


def image_nqubits(state):
    return int(state.size).bit_length() - 1


def encode(queue, nqubits, dtype, **kw):
    """Plan + serialise + encode `queue` with the real planner.Program code; returns
    [('image', Image) | ('raw', gate)] in execution order."""
    b = EncoderBackend(dtype)
    prog = planner.Program(b, queue, nqubits, dtype=dtype, **kw)
    out, k = [], 0
    for seg in prog.segments:
        if seg[0] == "program":
            img, tag, _ = b._lib.images[k]
            k += 1
            out.append(("image", Image(b._lib.lib, img, tag)))
        else:
            out.append(("raw", seg[1]))
    prog.segments = []
    return out
