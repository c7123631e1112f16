"""Pass planner: partitions a gate queue into multi-gate passes for ``qj_program_*``.

Upstream, ``Backend.execute_circuit`` (qibo) and ``MultiGpuOps.apply_gates``
(/root/reference/src/qibojit/backends/gpu.py:1467-1476) loop ``for gate in queue:
apply_gate(...)``: one pass over the state per gate.  The per-gate kernels here already run at
the HBM roofline, so the only way to make a circuit faster is fewer passes.

Three levels, all decided here on the host (pure Python, CPU-testable):

* a **pass** keeps a set of T index bits "local" (one tile of 2^T amplitudes resident in shared
  memory, always including the low contiguous run) and executes every gate whose non-diagonal
  targets are local; diagonal gates (Z, CZ, U1, CU1, RZ, ...) and controls need no locality;
* inside a pass the ops are grouped into **rounds**: each thread of the kernel gathers the
  2^J amplitudes spanned by J "register" bits of the tile and applies every op of the round to
  them, so dense targets of a round must be register bits of that round (gates that commute are
  hoisted into the open round);
* inside a round, diagonal gates are merged into phase **tables**, keyed by the set of register
  bits they touch, so that the kernel can slice a table along its register bits into per-thread
  constants (one table look-up per thread and slice).

The arithmetic of every gate is unchanged (same 2x2 / 4x4 complex mat-vec as
gates.py:16-38, 118-193; diagonal factors are multiplied together on the host in double
precision), only the order of memory traffic changes.
"""

import ctypes

import numpy as np

from . import _capi
from . import fusion

OP_DTYPE = np.dtype([
    ("kind", "<i4"), ("ntargets", "<i4"), ("ncontrols", "<i4"), ("reserved", "<i4"),
    ("data_offset", "<i8"),
    ("targets", "<i4", (_capi.QJ_MAX_DIAG_BITS,)),
    ("controls", "<i4", (_capi.QJ_MAX_QUBITS,)),
])
ROUND_DTYPE = np.dtype([
    ("nreg", "<i4"), ("reserved", "<i4"), ("first_op", "<i8"), ("nops", "<i8"),
    ("reg_bits", "<i4", (_capi.QJ_MAX_REG_BITS,)),
])
PASS_DTYPE = np.dtype([
    ("nlocal", "<i4"), ("reserved", "<i4"), ("first_round", "<i8"), ("nrounds", "<i8"),
    ("local_bits", "<i4", (_capi.QJ_MAX_LOCAL_BITS,)),
])
assert OP_DTYPE.itemsize == 264 and PASS_DTYPE.itemsize == 88 and ROUND_DTYPE.itemsize == 56

# diagonal in the computational basis: no locality needed
DIAGONAL_GATES = frozenset({
    "I", "Z", "S", "SDG", "T", "TDG", "RZ", "U1", "CZ", "CRZ", "CU1", "CCZ", "RZZ",
})

# complex128: 32 KiB tiles, four 128-thread CTAs per SM in different phases (measured 4-6 % faster
# than two 64 KiB tiles although a pass covers one qubit less); complex64: 64 KiB tiles
DEFAULT_TILE_BITS = {"complex128": 11, "complex64": 13}
# contiguous run of a tile in global memory: 256 bytes (whole sectors; measured as fast as 512-byte
# runs and it leaves one more tile bit for arbitrary high qubits: fewer passes)
DEFAULT_RUN_BITS = {"complex128": 3, "complex64": 5}
MAX_TILE_BITS = {"complex128": 12, "complex64": 13}   # 64 KiB tiles, two resident CTAs per SM
REG_BITS = {"complex128": 4, "complex64": 5}          # register bits per round (complex64: bit 0 + 4)
MAX_HI_BITS = 8
MIN_QUBITS = 6


class PlanOp:
    """One operation in index-bit space.  kind: 'dense' (1 or 2 targets), 'diag', 'raw'."""

    __slots__ = ("kind", "targets", "controls", "data", "gate", "bits", "tset", "dset")

    def __init__(self, kind, targets=(), controls=(), data=None, gate=None):
        self.kind = kind
        self.targets = tuple(targets)    # dense: matrix-index bit j <-> targets[j]; diag: table bit j
        self.controls = tuple(controls)
        self.data = data
        self.gate = gate
        self.bits = frozenset(self.targets) | frozenset(self.controls)
        # bits the op changes non-diagonally / bits it only reads (it commutes with any op that is
        # diagonal on them)
        self.tset = frozenset(self.targets) if kind == "dense" else frozenset()
        self.dset = self.bits - self.tset


def _is_diagonal(gate):
    name = gate.__class__.__name__
    if name in DIAGONAL_GATES:
        return True
    return bool(getattr(gate, "diagonal", False)) and name != "FusedGate"


def lower_gate(gate, nqubits, matrices):
    """qibo-style gate -> list of PlanOps (index bit m = nqubits - 1 - q, cpu.py:610)."""
    name = gate.__class__.__name__
    if name == "FanOut" or getattr(gate, "name", None) == "fanout":  # cpu.py:417-431
        c = nqubits - 1 - gate.control_qubits[0]
        x = np.array([[0, 1], [1, 0]], dtype=np.complex128)
        return [PlanOp("dense", (nqubits - 1 - t,), (c,), x) for t in gate.target_qubits]
    if name == "M" or hasattr(gate, "coefficients"):
        return [PlanOp("raw", gate=gate)]
    targets = [nqubits - 1 - q for q in gate.target_qubits]
    controls = [nqubits - 1 - q for q in gate.control_qubits]
    nt = len(targets)
    if nt > 2 and not _is_diagonal(gate):
        return [PlanOp("raw", gate=gate)]
    u = np.asarray(fusion.target_only_matrix(gate, matrices), dtype=np.complex128)
    if nt <= _capi.QJ_MAX_DIAG_BITS and (_is_diagonal(gate) or _matrix_is_diagonal(u)):
        # table bit j <-> targets reversed (first target = most significant matrix bit)
        return [PlanOp("diag", tuple(targets[::-1]), tuple(controls), np.diagonal(u).copy())]
    if nt > 2:
        return [PlanOp("raw", gate=gate)]
    return [PlanOp("dense", tuple(targets[::-1]), tuple(controls), u)]


def _matrix_is_diagonal(u):
    return u.ndim == 2 and not np.any(u - np.diag(np.diagonal(u)))


# ------------------------------------------------------------------------------- passes
def partition(ops, nqubits, tile_bits, run_bits):
    """Greedy partition of PlanOps into segments: ('pass', local_bits, [ops]) or ('raw', op)."""
    T = max(MIN_QUBITS, min(tile_bits, nqubits))
    r = max(1, min(run_bits, T))
    if T - r > MAX_HI_BITS:
        r = T - MAX_HI_BITS
    if T < nqubits:
        r = min(r, T - 2)    # a two-target gate on two high qubits needs two arbitrary tile bits
    segments = []
    remaining = list(ops)
    while remaining:
        if remaining[0].kind == "raw":
            segments.append(("raw", remaining.pop(0)))
            continue
        local = set(range(r))
        blocked_t, blocked_d = set(), set()
        taken, rest = [], []
        for i, op in enumerate(remaining):
            if len(blocked_t) == nqubits:
                rest.extend(remaining[i:])
                break
            if op.kind == "raw":
                blocked_t = set(range(nqubits))
                rest.append(op)
                continue
            if (op.tset & blocked_t) or (op.tset & blocked_d) or (op.dset & blocked_t):
                blocked_t |= op.tset
                blocked_d |= op.dset
                rest.append(op)
                continue
            if op.kind == "diag":
                taken.append(op)
                continue
            need = set(op.targets) - local
            if len(local) + len(need) <= T:
                local |= need
                taken.append(op)
            else:
                blocked_t |= op.tset
                blocked_d |= op.dset
                rest.append(op)
        if not taken:
            raise RuntimeError("pass planner made no progress (tile too small for the next gate)")
        b = 0
        while len(local) < T:  # pad with the lowest free bits: longer contiguous runs
            if b not in local:
                local.add(b)
            b += 1
        segments.append(("pass", sorted(local), taken))
        remaining = rest
    return segments


# ------------------------------------------------------------------------------- rounds
def schedule_rounds(ops, local_bits, nreg, fixed=()):
    """Ops of one pass -> [(reg_bits, [ops])]: every dense target of a round is one of its
    `nreg` register bits (`fixed` bits are register bits of every round).  An op is hoisted into
    the open round only over ops it commutes with, so non-commuting ops keep their order."""
    rounds = []
    remaining = list(ops)
    fixed = set(fixed)
    while remaining:
        regs = set(fixed)
        blocked_t, blocked_d = set(), set()
        taken, rest = [], []
        for op in remaining:
            free = not ((op.tset & blocked_t) or (op.tset & blocked_d) or (op.dset & blocked_t))
            if free and op.kind == "dense":
                need = op.tset - regs
                if len(regs) + len(need) <= nreg:
                    regs |= need
                else:
                    free = False
            if free:
                taken.append(op)
            else:
                blocked_t |= op.tset
                blocked_d |= op.dset
                rest.append(op)
        # pad with free local bits, preferably ones no diagonal op of the round touches (tables
        # without register bits cost one look-up per thread), highest first
        touched = set()
        for op in taken:
            touched |= op.dset
        for pool in ([b for b in reversed(local_bits) if b not in touched], list(reversed(local_bits))):
            for b in pool:
                if len(regs) >= nreg:
                    break
                regs.add(b)
        rounds.append((tuple(sorted(regs)), taken))
        remaining = rest
    return rounds


class _DiagAcc:
    def __init__(self):
        self.bits = []                      # table bit j <-> index bit bits[j]
        self.table = np.ones(1, dtype=np.complex128)

    def absorb(self, op):
        # controls of a diagonal op are table bits whose 0-branch is the identity: with the targets on
        # the low table bits the controlled entries are the top block
        obits = list(op.targets) + list(op.controls)
        nt, nc = len(op.targets), len(op.controls)
        data = np.asarray(op.data, dtype=np.complex128)
        if nc:
            otab = np.ones(1 << (nt + nc), dtype=np.complex128)
            otab[((1 << nc) - 1) << nt:] = data[:1 << nt]
        else:
            otab = data[:1 << nt]
        for b in obits:
            if b not in self.bits:
                self.bits.append(b)
                self.table = np.concatenate((self.table, self.table))
        # table[i] *= otab[the bits of i at the op's positions]: a broadcast product over the table
        # seen as a 2 x ... x 2 array (axis a <-> table bit n - 1 - a)
        n, m = len(self.bits), len(obits)
        axis_of = [n - 1 - self.bits.index(b) for b in obits]       # table axis of op bit k
        order = sorted(range(m), key=lambda k: axis_of[k])
        factor = otab.reshape((2,) * m).transpose([m - 1 - k for k in order])
        used = set(axis_of)
        factor = factor.reshape([2 if a in used else 1 for a in range(n)])
        self.table = (self.table.reshape((2,) * n) * factor).reshape(-1)

    def finish(self):
        """-> PlanOp with control-like bits split off (None when the product is the identity)."""
        bits, table = list(self.bits), self.table
        controls = []
        j = 0
        while j < len(bits):
            nd = table.reshape(1 << (len(bits) - 1 - j), 2, 1 << j)      # middle axis = table bit j
            if np.all(nd[:, 0, :] == 1.0):
                controls.append(bits.pop(j))
                table = np.ascontiguousarray(nd[:, 1, :]).reshape(-1)
            else:
                j += 1
        if np.all(table == 1.0):
            return None
        return PlanOp("diag", tuple(bits), tuple(controls), table)


def merge_round_diagonals(ops, reg_bits, max_diag_bits, local_bits=None):
    """Merge the (mutually commuting) diagonal ops of one round into phase tables, one family of
    tables per set of register bits touched; a dense op flushes the tables that touch its targets.

    Within a family the tables are kept apart by where their remaining bits live: all inside the
    tile ('T': the kernel folds these into one precomputed factor per thread), all outside ('O':
    one factor per tile) or mixed ('X': a per-thread look-up).  Ladders of two-qubit phases (QFT)
    fall entirely into T and O."""
    regs = frozenset(reg_bits)
    local = frozenset(local_bits) if local_bits is not None else None
    out, accs = [], []   # accs: [(signature, _DiagAcc)]

    def flush(entry):
        accs.remove(entry)
        op = entry[1].finish()
        if op is not None:
            out.append(op)

    def where(bits):
        rest = bits - regs
        if local is None or not rest:
            return "T"
        if rest <= local:
            return "T"
        if not (rest & local):
            return "O"
        return "X"

    for op in ops:
        if op.kind == "diag":
            sig = (op.bits & regs, where(op.bits))
            obits = set(op.bits)
            best, best_key = None, None
            for entry in accs:
                if entry[0] != sig:
                    continue
                abits = set(entry[1].bits)
                if len(obits | abits) > max_diag_bits:
                    continue
                key = (len(obits - abits), -len(obits & abits))
                if best is None or key < best_key:
                    best, best_key = entry, key
            if best is None:
                best = (sig, _DiagAcc())
                accs.append(best)
            best[1].absorb(op)
            continue
        for entry in [e for e in accs if e[0][0] & op.tset]:
            flush(entry)
        out.append(op)
    for entry in list(accs):
        flush(entry)
    return out


_SWAP_MATRIX = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)


def relabel_swaps_away(ops, nqubits):
    """For a circuit that starts from |0...0>: drop every uncontrolled SWAP and relabel the ops
    BEFORE it instead.  g_k ... g_1 with g_k = SWAP(a, b) equals (g'_{k-1} ... g'_1) SWAP(a, b)
    where g' is g with a and b exchanged; the SWAP then acts on |0...0>, which it leaves alone.
    (The distributed layer does the same with its qubit map.)  Returns `ops` unchanged when a raw
    gate (whose qubits cannot be relabelled here) is present."""
    if any(op.kind == "raw" for op in ops):
        return ops
    perm = list(range(nqubits))     # index bit -> relabelled index bit, for the ops still to come (going backwards)
    out = []
    for op in reversed(ops):
        if (op.kind == "dense" and len(op.targets) == 2 and not op.controls
                and np.array_equal(np.asarray(op.data).reshape(4, 4), _SWAP_MATRIX)):
            a, b = perm[op.targets[0]], perm[op.targets[1]]
            perm = [b if x == a else a if x == b else x for x in perm]
            continue
        out.append(PlanOp(op.kind, [perm[t] for t in op.targets], [perm[c] for c in op.controls], op.data, op.gate))
    out.reverse()
    return out


def fuse_one_qubit_runs(ops):
    """Multiply uncontrolled one-target gates that follow each other on the same bit (nothing in
    between touches that bit) into one 2x2: H then RX on a qubit costs one register update instead
    of two.  A one-bit uncontrolled phase next to such a gate is folded in as well."""
    out = []
    last = {}          # index bit -> position in `out` of the last op that touches it

    def as_matrix(op):
        if op.kind == "dense" and len(op.targets) == 1 and not op.controls:
            return np.asarray(op.data, dtype=np.complex128).reshape(2, 2)
        if op.kind == "diag" and len(op.targets) == 1 and not op.controls:
            return np.diag(np.asarray(op.data, dtype=np.complex128).reshape(2))
        return None

    for op in ops:
        m = as_matrix(op)
        if m is not None:
            t = op.targets[0]
            j = last.get(t)
            if j is not None:
                prev = as_matrix(out[j])
                if prev is not None and out[j].targets[0] == t and (op.kind == "dense" or out[j].kind == "dense"):
                    out[j] = PlanOp("dense", (t,), (), m @ prev)
                    continue
        if op.kind != "raw":
            for b in op.bits:
                last[b] = len(out)
        else:
            last = {b: len(out) for b in list(last)}   # a raw gate's qubits are not known here: it blocks every bit seen so far
            last[-1] = len(out)
        out.append(op)
    return out


# ------------------------------------------------------------------------------- pair blocks
# Cost model of the pass kernel in issue cycles per thread and op (both dtypes: a thread holds 16
# complex128 / 32 complex64 amplitudes and the FP64 / packed-FP32x2 pipes issue every other cycle):
# what `fuse_pair_blocks` weighs a 4x4 block against.
_COST_REAL1, _COST_CPLX1 = 128.0, 256.0      # one-target gate: real or axis-aligned / complex
_COST_REAL2, _COST_CPLX2 = 256.0, 512.0      # two-target gate: real / complex 4x4
_COST_DISPATCH = 60.0                        # header decode + branch of one op
_COST_GROUPED = 20.0                         # share of a dispatch for a one-target gate (grouped by up to J)
_COST_SIGN, _COST_PHASE = 80.0, 130.0        # +-1 diagonal / general phase op
_PAIR_MARGIN = 1.25                          # fuse only when the separate ops cost this much more


def _is_real(m):
    return not np.any(np.asarray(m).imag)


def _op_cost(op):
    d = np.asarray(op.data)
    if op.kind == "diag":
        return _COST_SIGN if np.all((d.imag == 0) & (np.abs(d.real) == 1)) else _COST_PHASE
    scale = 0.5 ** len(op.controls)
    if len(op.targets) == 1:
        m = d.reshape(2, 2)
        if np.array_equal(m, [[0, 1], [1, 0]]):
            return _COST_DISPATCH
        axis = not (m[0, 0].imag or m[1, 1].imag or m[0, 1].real or m[1, 0].real)
        body = _COST_REAL1 if (_is_real(m) or axis) else _COST_CPLX1
        return body * scale + (_COST_DISPATCH if op.controls else _COST_GROUPED)
    m = d.reshape(4, 4)
    if np.array_equal(m, _SWAP_MATRIX):
        return _COST_DISPATCH
    return (_COST_REAL2 if _is_real(m) else _COST_CPLX2) * scale + _COST_DISPATCH


def op_matrix(op, bits):
    """Matrix of a dense / diagonal PlanOp on the index bits `bits` (matrix-index bit j <-> bits[j];
    every bit of the op must be among them)."""
    k = len(bits)
    pos = {b: i for i, b in enumerate(bits)}
    dim = 1 << k
    out = np.zeros((dim, dim), dtype=np.complex128)
    data = np.asarray(op.data, dtype=np.complex128)
    nt = len(op.targets)
    tmask = 0
    for t in op.targets:
        tmask |= 1 << pos[t]
    for x in range(dim):
        if any(not (x >> pos[c]) & 1 for c in op.controls):
            out[x, x] = 1.0
            continue
        xin = 0
        for j, t in enumerate(op.targets):
            xin |= ((x >> pos[t]) & 1) << j
        if op.kind == "diag":
            out[x, x] = data.reshape(-1)[xin]
            continue
        m = data.reshape(1 << nt, 1 << nt)
        for yout in range(1 << nt):
            y = x & ~tmask
            for j, t in enumerate(op.targets):
                y |= ((yout >> j) & 1) << pos[t]
            out[y, x] = m[yout, xin]
    return out


def fuse_pair_blocks(ops):
    """Multiply runs of one- and two-bit ops on the same pair of index bits into ONE 4x4 dense op
    where the pass kernel runs that faster than the separate ops (a real 4x4 costs 8 multiply-adds
    per amplitude -- as much as two real rotations -- so RY RY CZ RY RY on a pair halves the
    arithmetic; qibo's ``Circuit.fuse(max_qubits=2)`` forms the same blocks unconditionally).

    A block grows from a two-bit op (the anchor) over the ops that are adjacent to it on both bits'
    timelines, so every member can be moved to the anchor's position without crossing an op it
    does not commute with.  Candidate blocks are taken best-gain first (lazy greedy: a candidate
    whose members were claimed meanwhile is re-evaluated) and only when the separate ops cost
    `_PAIR_MARGIN` times the block."""
    import heapq

    n = len(ops)
    if any(op.kind == "raw" for op in ops):
        # a raw gate's qubits are unknown here: fuse the stretches between raw gates
        out, seg = [], []
        for op in ops:
            if op.kind == "raw":
                out.extend(fuse_pair_blocks(seg))
                out.append(op)
                seg = []
            else:
                seg.append(op)
        out.extend(fuse_pair_blocks(seg))
        return out
    timeline = {}
    where = [dict() for _ in range(n)]     # op -> {bit: position in the bit's timeline}
    for i, op in enumerate(ops):
        for b in op.bits:
            tl = timeline.setdefault(b, [])
            where[i][b] = len(tl)
            tl.append(i)

    def single(op):
        return len(op.bits) == 1 and not op.controls

    free = [True] * n

    def grow(anchor):
        a, b = sorted(ops[anchor].bits)
        pair = frozenset((a, b))
        members = {anchor}
        first = {a: where[anchor][a], b: where[anchor][b]}
        last = dict(first)
        for step, edge in ((-1, first), (1, last)):
            moved = True
            while moved:
                moved = False
                cand = {}
                for bit in (a, b):
                    p = edge[bit] + step
                    if 0 <= p < len(timeline[bit]):
                        cand[bit] = timeline[bit][p]
                for bit, j in cand.items():
                    if not free[j] or j in members:
                        continue
                    if single(ops[j]):
                        members.add(j)
                        edge[bit] += step
                        moved = True
                    elif ops[j].bits == pair and cand.get(a) == j and cand.get(b) == j:
                        members.add(j)
                        edge[a] += step
                        edge[b] += step
                        moved = True
                        break
        return sorted(members), (a, b)

    def evaluate(anchor):
        members, (a, b) = grow(anchor)
        if len(members) < 2:
            return None
        m = np.eye(4, dtype=np.complex128)
        for j in members:
            m = op_matrix(ops[j], (a, b)) @ m
        m[np.abs(m) < 1e-300] = 0.0
        separate = sum(_op_cost(ops[j]) for j in members)
        if not np.any(m - np.diag(np.diagonal(m))):
            fused = PlanOp("diag", (a, b), (), np.diagonal(m).copy())
        else:
            fused = PlanOp("dense", (a, b), (), m)
        cost = _op_cost(fused)
        if separate < _PAIR_MARGIN * cost:
            return None
        return separate - cost, members, fused

    heap = []
    for i, op in enumerate(ops):
        if len(op.bits) == 2 and op.kind in ("dense", "diag"):
            ev = evaluate(i)
            if ev is not None:
                heapq.heappush(heap, (-ev[0], i, tuple(ev[1])))
    replaced = {}
    while heap:
        _, i, members = heapq.heappop(heap)
        if not free[i]:
            continue
        ev = evaluate(i)
        if ev is None:
            continue
        if tuple(ev[1]) != members:           # shrunk since it was queued: back into the queue
            heapq.heappush(heap, (-ev[0], i, tuple(ev[1])))
            continue
        for j in ev[1]:
            free[j] = False
        replaced[i] = ev[2]
    return [replaced.get(i, op) for i, op in enumerate(ops) if free[i] or i in replaced]


def plan_queue(queue, nqubits, matrices, tile_bits, run_bits, max_diag_bits=10, dtype="complex128",
               zero_state=False, pair_blocks=True):
    """Gate queue -> [('pass', local_bits, [(reg_bits, [PlanOp])]) | ('raw', gate)] (host logic).
    `zero_state`: the program will only ever run on |0...0> (SWAP gates become relabellings)."""
    ops = []
    for gate in queue:
        ops.extend(lower_gate(gate, nqubits, matrices))
    if zero_state:
        ops = relabel_swaps_away(ops, nqubits)
    ops = fuse_one_qubit_runs(ops)
    if pair_blocks:
        ops = fuse_pair_blocks(ops)
    nreg = REG_BITS[str(dtype)]
    fixed = (0,) if str(dtype) == "complex64" else ()
    mdb = min(max_diag_bits, _capi.QJ_MAX_DIAG_BITS)
    out = []
    for seg in partition(ops, nqubits, tile_bits, run_bits):
        if seg[0] == "raw":
            out.append(("raw", seg[1].gate))
            continue
        rounds = []
        for regs, rops in schedule_rounds(seg[2], seg[1], min(nreg, len(seg[1])), fixed):
            merged = merge_round_diagonals(rops, regs, mdb, seg[1])
            if merged:
                rounds.append((regs, merged))
        if rounds:
            out.append(("pass", seg[1], rounds))
    return out


# ------------------------------------------------------------------------------- programs
class Program:
    """A compiled gate queue: device-resident tile programs interleaved with raw gates that the
    per-gate kernels execute (dense gates on >= 3 targets, measurements)."""

    def __init__(self, backend, queue, nqubits, dtype=None, tile_bits=None, run_bits=None,
                 max_diag_bits=10, zero_state=False, pair_blocks=True):
        self.backend = backend
        self.pair_blocks = bool(pair_blocks)
        self.zero_state = bool(zero_state)   # valid on |0...0> only (SWAP gates relabelled away)
        self.nqubits = int(nqubits)
        self.dtype = str(dtype or backend.dtype)
        self.tile_bits = min(int(tile_bits or DEFAULT_TILE_BITS[self.dtype]), MAX_TILE_BITS[self.dtype])
        self.run_bits = int(run_bits or DEFAULT_RUN_BITS[self.dtype])
        self.max_diag_bits = min(int(max_diag_bits), _capi.QJ_MAX_DIAG_BITS)
        self.ngates = len(queue)
        self.closed = False
        self.segments = []       # ('program', handle, npasses) | ('raw', gate)
        self.passes = []         # [(local_bits, [(reg_bits, [PlanOp])])] for inspection
        self.upload_bytes = 0    # host bytes handed to qj_program_create (descriptors + matrices/tables)
        if self.nqubits < MIN_QUBITS:   # too small for a tile: gate by gate
            self.segments = [("raw", g) for g in queue]
            return
        pending = []
        for seg in plan_queue(queue, self.nqubits, backend.custom_matrices, self.tile_bits,
                              self.run_bits, self.max_diag_bits, self.dtype, self.zero_state,
                              self.pair_blocks):
            if seg[0] == "raw":
                self._flush(pending)
                pending = []
                self.segments.append(("raw", seg[1]))
            else:
                pending.append((seg[1], seg[2]))
        self._flush(pending)

    # -- serialisation
    def _flush(self, passes):
        if not passes:
            return
        np_dtype = np.dtype(self.dtype)
        nrounds = sum(len(p[1]) for p in passes)
        nops = sum(len(r[1]) for p in passes for r in p[1])
        op_arr = np.zeros(nops, dtype=OP_DTYPE)
        round_arr = np.zeros(nrounds, dtype=ROUND_DTYPE)
        pass_arr = np.zeros(len(passes), dtype=PASS_DTYPE)
        chunks, offset, k, ri = [], 0, 0, 0
        for pi, (local_bits, rounds) in enumerate(passes):
            pass_arr[pi]["nlocal"] = len(local_bits)
            pass_arr[pi]["first_round"] = ri
            pass_arr[pi]["nrounds"] = len(rounds)
            pass_arr[pi]["local_bits"][:len(local_bits)] = local_bits
            for regs, ops in rounds:
                round_arr[ri]["nreg"] = len(regs)
                round_arr[ri]["reg_bits"][:len(regs)] = regs
                round_arr[ri]["first_op"] = k
                round_arr[ri]["nops"] = len(ops)
                ri += 1
                for op in ops:
                    rec = op_arr[k]
                    data = np.ascontiguousarray(np.asarray(op.data, dtype=np_dtype).reshape(-1))
                    if op.kind == "dense":
                        rec["kind"] = _capi.QJ_OPK_DENSE1 if len(op.targets) == 1 else _capi.QJ_OPK_DENSE2
                    else:
                        rec["kind"] = _capi.QJ_OPK_DIAG
                    rec["ntargets"] = len(op.targets)
                    rec["targets"][:len(op.targets)] = op.targets
                    rec["ncontrols"] = len(op.controls)
                    rec["controls"][:len(op.controls)] = op.controls
                    rec["data_offset"] = offset
                    chunks.append(data)
                    offset += data.size
                    k += 1
        data = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np_dtype)
        b = self.backend
        handle = ctypes.c_void_p()
        _capi.check(b._lib.qj_program_create(
            b._handle(), _capi.QJ_C128 if self.dtype == "complex128" else _capi.QJ_C64, self.nqubits,
            pass_arr.ctypes.data, len(passes), round_arr.ctypes.data, nrounds, op_arr.ctypes.data, nops,
            data.ctypes.data, int(data.size), ctypes.byref(handle)))
        self.upload_bytes += pass_arr.nbytes + round_arr.nbytes + op_arr.nbytes + data.nbytes
        self.segments.append(("program", handle, len(passes)))
        self.passes.extend(passes)

    # -- execution
    def _check_open(self):
        if getattr(self, "closed", False):
            raise RuntimeError("this Program was closed (its device image is released): compile the circuit again")

    def _prepare_zero(self, state):
        """|0...0> by the initial-state kernel (when the first segment is not a pass program that
        can fuse the preparation in)."""
        b = self.backend
        _capi.check(b._lib.qj_initial_state(b._handle(), state.data_ptr(), b._tag(state), self.nqubits))

    def run(self, state, from_zero=False):
        """Apply the program to `state` in place.  `from_zero`: the input is |0...0> and `state` need
        not hold it (it may be uninitialised memory): the first pass writes every amplitude without
        reading any -- `initial_state_vector` (ops.py:14-18) fused into the first pass."""
        b = self.backend
        self._check_open()
        if state.numel() != (1 << self.nqubits) or str(state.dtype).replace("torch.", "") != self.dtype:
            raise ValueError("state does not match the program's qubit count / dtype")
        if from_zero and not (self.segments and self.segments[0][0] == "program"):
            self._prepare_zero(state)
            from_zero = False
        for seg in self.segments:
            if seg[0] == "program":
                flags = _capi.QJ_RUN_ZERO_INPUT if from_zero else 0
                from_zero = False
                _capi.check(b._lib.qj_program_run_ex(b._handle(), seg[1], state.data_ptr(), 0, -1, flags))
            else:
                state = seg[1].apply(b, state, self.nqubits)
        return state

    def run_timed(self, state, timer, from_zero=False):
        """Like run(), but every kernel launch goes through ``timer(kind, fraction, fn)`` where
        `fraction` is the launch's algorithmic traffic in units of one full read+write of the
        state -- bench.py brackets `fn` with CUDA events.  A pass that
        starts from |0...0> (`from_zero`) only writes: fraction 0.5."""
        from .backends.b200 import GATE_OPS

        b = self.backend
        self._check_open()
        if from_zero and not (self.segments and self.segments[0][0] == "program"):
            timer("init", 0.5, lambda: self._prepare_zero(state))
            from_zero = False
        for seg in self.segments:
            if seg[0] == "program":
                a = ctypes.c_int64()
                _capi.check(b._lib.qj_program_stats(seg[1], ctypes.byref(a), None, None))
                for i in range(a.value):
                    flags = _capi.QJ_RUN_ZERO_INPUT if (from_zero and i == 0) else 0
                    timer("pass0" if flags else "pass", 0.5 if flags else 1.0, lambda i=i, flags=flags: _capi.check(
                        b._lib.qj_program_run_ex(b._handle(), seg[1], state.data_ptr(), i, 1, flags)))
                from_zero = False
            else:
                g = seg[1]
                c = len(g.control_qubits)
                op = getattr(g, "op", None) or GATE_OPS.get(g.__class__.__name__)
                if op in ("apply_z", "apply_z_pow", "apply_swap"):
                    kind, frac = ("swap" if op == "apply_swap" else "diag"), 2.0 ** -(c + 1)
                elif op == "apply_fsim":
                    kind, frac = "fsim", 0.75 * 2.0 ** -c
                else:
                    kind, frac = f"dense{len(g.target_qubits)}" + (f"c{c}" if c else ""), 2.0 ** -c
                out = []
                timer(kind, frac, lambda: out.append(g.apply(b, state, self.nqubits)))
                state = out[0]
        return state

    def fma_per_amplitude(self):
        """Real multiply-adds per amplitude the pass kernel executes for this program (the
        arithmetic side of the roofline): 4 for a real or axis-aligned one-target gate, 8 for a
        complex one, 16 for a two-target gate (8 when its matrix is real), 4 per phase multiply;
        permutations and sign flips cost none; controls scale by 2^-c."""
        return float(sum(self.fma_per_pass()))

    def fma_per_pass(self):
        """The same count, one entry per pass."""
        return [self._fma_of(rounds) for _, rounds in self.passes]

    @staticmethod
    def _fma_of(rounds):
        total = 0.0
        if True:
            for _, ops in rounds:
                for op in ops:
                    frac = 2.0 ** -len(op.controls)
                    d = np.asarray(op.data)
                    if op.kind == "dense" and len(op.targets) == 1:
                        m = d.reshape(2, 2)
                        if np.array_equal(m, [[0, 1], [1, 0]]):
                            continue
                        real = not np.any(m.imag)
                        axis = not (m[0, 0].imag or m[1, 1].imag or m[0, 1].real or m[1, 0].real)
                        total += (4.0 if (real or axis) else 8.0) * frac
                    elif op.kind == "dense":
                        m = d.reshape(4, 4)
                        if np.array_equal(m, [[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]):
                            continue            # SWAP: register renaming
                        total += (8.0 if not np.any(m.imag) else 16.0) * frac
                    elif op.kind == "diag":
                        if np.all((d.imag == 0) & (np.abs(d.real) == 1)):
                            continue
                        total += 4.0 * frac * float(np.mean(d != 1.0))
        return total

    def stats(self):
        """{'launches', 'rounds', 'micro_ops', 'raw_gates'} summed over the segments."""
        out = dict(launches=0, rounds=0, micro_ops=0, raw_gates=0, passes=len(self.passes))
        for seg in self.segments:
            if seg[0] == "raw":
                out["raw_gates"] += 1
                continue
            a, r, m = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
            _capi.check(self.backend._lib.qj_program_stats(seg[1], ctypes.byref(a), ctypes.byref(r),
                                                          ctypes.byref(m)))
            out["launches"] += a.value
            out["rounds"] += r.value
            out["micro_ops"] += m.value
        return out

    def launches(self):
        """[(segment handle, launch index)] of every kernel launch, for per-pass timing."""
        out = []
        for seg in self.segments:
            if seg[0] != "program":
                continue
            a = ctypes.c_int64()
            _capi.check(self.backend._lib.qj_program_stats(seg[1], ctypes.byref(a), None, None))
            out.extend((seg[1], i) for i in range(a.value))
        return out

    def close(self):
        """Release the device images.  A closed program raises on run() instead of silently
        returning its input."""
        for seg in self.segments:
            if seg[0] == "program" and seg[1]:
                # the library synchronises the creating handle's stream when it is still alive;
                # a destroyed backend passes None (cudaFree synchronises by itself)
                self.backend._lib.qj_program_destroy(self.backend._handle_or_none(), seg[1])
        self.segments = []
        self.closed = True

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
