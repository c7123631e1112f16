"""Gate fusion: greedy block planner and fused-matrix builder.

Upstream this lives in qibo (``Circuit.fuse`` / ``FusedGate`` / ``Backend.matrix_fused``;
the reference only consumes the result, /root/reference/src/qibojit/backends/cpu.py:535-537).
qibo is not available offline, so the planner is restated here; it produces ``FusedGate``
blocks over at most ``max_qubits`` qubits whose dense matrix costs one pass over the state.
"""

import numpy as np

from . import gates as G


def target_only_matrix(gate, matrices):
    """2^t x 2^t matrix acting on the gate's targets (first target = most significant)."""
    name = gate.__class__.__name__
    if name == "FusedGate":
        return fused_matrix(gate, matrices)
    m = gate.target_matrix(matrices) if hasattr(gate, "target_matrix") else getattr(matrices, name)
    m = np.asarray(m)
    if name in ("U1", "CU1"):  # scalar phase (matrices.py:28-33)
        return np.array([[1, 0], [0, complex(m)]], dtype=np.complex128)
    if name in ("fSim", "GeneralizedfSim"):  # 5-vector (matrices.py:62-70)
        out = np.eye(4, dtype=np.complex128)
        out[1, 1], out[1, 2], out[2, 1], out[2, 2], out[3, 3] = m[0], m[1], m[2], m[3], m[4]
        return out
    t = len(gate.target_qubits)
    return m.astype(np.complex128).reshape(1 << t, 1 << t)


def full_matrix(gate, matrices):
    """Matrix on (controls..., targets...) with controls as the most significant bits."""
    u = target_only_matrix(gate, matrices)
    c = len(gate.control_qubits)
    if c == 0:
        return u
    dim = u.shape[0] << c
    out = np.eye(dim, dtype=np.complex128)
    out[dim - u.shape[0]:, dim - u.shape[0]:] = u
    return out


def embed_apply(block, u, positions, k):
    """Left-multiply the 2^k x 2^k `block` by `u` acting on block-qubit `positions`
    (position 0 = most significant qubit of the block)."""
    g = len(positions)
    cols = block.shape[1]
    t = block.reshape((2,) * k + (cols,))
    ut = np.asarray(u, dtype=np.complex128).reshape((2,) * (2 * g))
    t = np.tensordot(ut, t, axes=(list(range(g, 2 * g)), list(positions)))
    t = np.moveaxis(t, list(range(g)), list(positions))
    return t.reshape(1 << k, cols)


def _gate_key(gate):
    """Value summary of a gate (class, qubits, parameters) for cache validation."""
    def param(p):
        if hasattr(p, "tobytes"):
            a = np.ascontiguousarray(p)
            return (a.shape, str(a.dtype), a.tobytes())
        if isinstance(p, (list, tuple)):
            return tuple(param(x) for x in p)
        return repr(p)

    key = (gate.__class__.__name__, tuple(gate.target_qubits), tuple(gate.control_qubits),
           tuple(param(x) for x in getattr(gate, "parameters", ())))
    if hasattr(gate, "gates"):
        key += (tuple(_gate_key(g) for g in gate.gates),)
    return key


def fused_matrix(fgate, matrices):
    """Dense matrix of a FusedGate over its ``target_qubits`` (qibo ``matrix_fused``).

    The product is built and cached in complex128 and cast to the asking backend's dtype on return
    (a block first compiled under complex64 must not hand a float32-precision matrix to a later
    complex128 program); the cache is keyed on the inner gates' parameters, so re-parametrising an
    inner gate rebuilds it."""
    key = tuple(_gate_key(g) for g in fgate.gates)
    cached = getattr(fgate, "_matrix", None)
    if cached is not None and getattr(fgate, "_matrix_key", None) == key:
        return cached.astype(matrices.dtype)
    from .matrices import CustomMatrices

    wide = matrices if str(matrices.dtype) == "complex128" else CustomMatrices("complex128")
    bq = list(fgate.target_qubits)
    k = len(bq)
    block = np.eye(1 << k, dtype=np.complex128)
    for gate in fgate.gates:
        qs = list(gate.control_qubits) + list(gate.target_qubits)
        block = embed_apply(block, full_matrix(gate, wide), [bq.index(q) for q in qs], k)
    fgate._matrix = block
    fgate._matrix_key = key
    return block.astype(matrices.dtype)


def _fusable(gate):
    name = gate.__class__.__name__
    return name not in ("M", "FanOut") and not hasattr(gate, "coefficients")


def fuse(queue, max_qubits=2):
    """Greedy fusion of a gate list into FusedGate blocks of at most `max_qubits` qubits.

    Invariant: every open block is the last operation on each of its qubits, so open blocks
    commute with each other and may be merged; a gate that overlaps an open block either joins
    it or closes it, hence blocks sharing a qubit are always emitted in program order."""
    out = []
    open_on = {}  # qubit -> block (dict with 'qubits' list, 'gates' list)

    def close(block):
        if block.get("closed"):
            return
        block["closed"] = True
        for q in block["qubits"]:
            if open_on.get(q) is block:
                del open_on[q]
        out.append(block)

    for gate in queue:
        qs = list(gate.control_qubits) + list(gate.target_qubits)
        touching = []
        for q in qs:
            b = open_on.get(q)
            if b is not None and all(b is not t for t in touching):
                touching.append(b)
        if not _fusable(gate) or len(qs) > max_qubits:
            for b in touching:
                close(b)
            out.append({"qubits": qs, "gates": [gate], "closed": True, "raw": True})
            continue
        union = list(dict.fromkeys([q for b in touching for q in b["qubits"]] + qs))
        if len(union) <= max_qubits:
            if touching:
                block = touching[0]
                for other in touching[1:]:
                    block["gates"].extend(other["gates"])
                    other["closed"] = True
                    other["merged"] = True
                block["qubits"] = union
            else:
                block = {"qubits": union, "gates": []}
            block["gates"].append(gate)
            for q in union:
                open_on[q] = block
        else:
            for b in touching:
                close(b)
            block = {"qubits": list(qs), "gates": [gate]}
            for q in qs:
                open_on[q] = block
    for b in list(dict.fromkeys(id(b) for b in open_on.values())):
        pass
    seen = []
    for b in open_on.values():
        if all(b is not s for s in seen):
            seen.append(b)
    for b in seen:
        close(b)

    fused = []
    for b in out:
        if b.get("merged"):
            continue
        if b.get("raw") or len(b["gates"]) == 1:
            fused.append(b["gates"][0])
            continue
        fg = G.FusedGate(*sorted(b["qubits"]))
        for g in b["gates"]:
            fg.append(g)
        fused.append(fg)
    return fused
