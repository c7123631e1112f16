"""TEST INFRASTRUCTURE ONLY -- ctypes front end of the CPU oracle (oracle/qj_oracle.c).

Exposes the reference's *kernel-module* interface
(``/root/reference/src/qibojit/custom_operators/gates.py`` and ``ops.py``: same function
names, argument order and in-place/return-the-state behaviour) on numpy arrays, so the
parity tests call the oracle exactly the way the reference's backend calls numba.

Only ``tests/``, ``__graft_entry__.smoke()`` and bench.py's CPU-baseline / ``--impl
reference`` legs may import this module.  The product package never does.
"""

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libqj_oracle.so")

OP_GATE, OP_X, OP_Y, OP_Z, OP_ZPOW, OP_SWAP, OP_FSIM = range(7)


def build(force=False):
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    src = [os.path.join(_HERE, f) for f in ("qj_oracle.c", "qj_oracle_kernels.inc")]
    stale = not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libqj_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.qjo_max_threads.restype = ctypes.c_int
    return _lib


def _suffix(state):
    if state.dtype == np.complex128:
        return "c128"
    if state.dtype == np.complex64:
        return "c64"
    raise TypeError(f"oracle: unsupported state dtype {state.dtype}")


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _check(state):
    if not (isinstance(state, np.ndarray) and state.flags.c_contiguous):
        raise TypeError("oracle works in place on C-contiguous numpy arrays")


def _gate(gate, state):
    if gate is None:
        return None, None
    g = np.ascontiguousarray(np.asarray(gate, dtype=state.dtype).ravel())
    return g, _ptr(g)


def _qubits(qubits):
    if qubits is None:
        return None, None, 0
    q = np.ascontiguousarray(np.asarray(qubits, dtype=np.int32))
    return q, _ptr(q), int(q.size)


def max_threads():
    return int(lib().qjo_max_threads())


def set_threads(n):
    lib().qjo_set_threads(ctypes.c_int(int(n)))


# --------------------------------------------------------------------- gates.py
def _one(op, state, gate, qubits, nstates, m):
    _check(state)
    g, gp = _gate(gate, state)
    q, qp, nq = _qubits(qubits)
    fn = getattr(lib(), f"qjo_one_qubit_{_suffix(state)}")
    fn(_ptr(state), ctypes.c_int(op), gp, qp, ctypes.c_int(nq),
       ctypes.c_int64(int(nstates)), ctypes.c_int(int(m)))
    return state


def _two(op, state, gate, qubits, nstates, m1, m2, swap_targets):
    _check(state)
    g, gp = _gate(gate, state)
    q, qp, nq = _qubits(qubits)
    fn = getattr(lib(), f"qjo_two_qubit_{_suffix(state)}")
    fn(_ptr(state), ctypes.c_int(op), gp, qp, ctypes.c_int(nq), ctypes.c_int64(int(nstates)),
       ctypes.c_int(int(m1)), ctypes.c_int(int(m2)), ctypes.c_int(int(bool(swap_targets))))
    return state


def apply_gate_kernel(state, gate, nstates, m):
    return _one(OP_GATE, state, gate, None, nstates, m)


def multicontrol_apply_gate_kernel(state, gate, qubits, nstates, m):
    return _one(OP_GATE, state, gate, qubits, nstates, m)


def apply_x_kernel(state, gate, nstates, m):
    return _one(OP_X, state, None, None, nstates, m)


def multicontrol_apply_x_kernel(state, gate, qubits, nstates, m):
    return _one(OP_X, state, None, qubits, nstates, m)


def apply_y_kernel(state, gate, nstates, m):
    return _one(OP_Y, state, None, None, nstates, m)


def multicontrol_apply_y_kernel(state, gate, qubits, nstates, m):
    return _one(OP_Y, state, None, qubits, nstates, m)


def apply_z_kernel(state, gate, nstates, m):
    return _one(OP_Z, state, None, None, nstates, m)


def multicontrol_apply_z_kernel(state, gate, qubits, nstates, m):
    return _one(OP_Z, state, None, qubits, nstates, m)


def apply_z_pow_kernel(state, gate, nstates, m):
    return _one(OP_ZPOW, state, np.asarray(gate).reshape(1), None, nstates, m)


def multicontrol_apply_z_pow_kernel(state, gate, qubits, nstates, m):
    return _one(OP_ZPOW, state, np.asarray(gate).reshape(1), qubits, nstates, m)


def apply_two_qubit_gate_kernel(state, gate, nstates, m1, m2, swap_targets=False):
    return _two(OP_GATE, state, gate, None, nstates, m1, m2, swap_targets)


def multicontrol_apply_two_qubit_gate_kernel(state, gate, qubits, nstates, m1, m2,
                                             swap_targets=False):
    return _two(OP_GATE, state, gate, qubits, nstates, m1, m2, swap_targets)


def apply_swap_kernel(state, gate, nstates, m1, m2, swap_targets=False):
    return _two(OP_SWAP, state, None, None, nstates, m1, m2, swap_targets)


def multicontrol_apply_swap_kernel(state, gate, qubits, nstates, m1, m2, swap_targets=False):
    return _two(OP_SWAP, state, None, qubits, nstates, m1, m2, swap_targets)


def apply_fsim_kernel(state, gate, nstates, m1, m2, swap_targets=False):
    return _two(OP_FSIM, state, gate, None, nstates, m1, m2, swap_targets)


def multicontrol_apply_fsim_kernel(state, gate, qubits, nstates, m1, m2, swap_targets=False):
    return _two(OP_FSIM, state, gate, qubits, nstates, m1, m2, swap_targets)


def apply_multi_qubit_gate_kernel(state, gate, qubits, nstates, targets):
    _check(state)
    g, gp = _gate(gate, state)
    q, qp, nq = _qubits(qubits)
    t = np.ascontiguousarray(np.asarray(targets, dtype=np.int64))
    fn = getattr(lib(), f"qjo_multi_qubit_{_suffix(state)}")
    fn(_ptr(state), gp, qp, ctypes.c_int(nq), ctypes.c_int64(int(nstates)), _ptr(t),
       ctypes.c_int(int(t.size)))
    return state


apply_three_qubit_gate_kernel = apply_multi_qubit_gate_kernel
apply_four_qubit_gate_kernel = apply_multi_qubit_gate_kernel
apply_five_qubit_gate_kernel = apply_multi_qubit_gate_kernel


# ----------------------------------------------------------------------- ops.py
def initial_state_vector(state):
    _check(state)
    getattr(lib(), f"qjo_initial_state_{_suffix(state)}")(_ptr(state), ctypes.c_int64(state.size))
    return state


def _collapse(state, qubits, result, nqubits, normalize):
    _check(state)
    q, qp, nq = _qubits(qubits)
    fn = getattr(lib(), f"qjo_collapse_state_{_suffix(state)}")
    fn(_ptr(state), qp, ctypes.c_int(nq), ctypes.c_int64(int(result)), ctypes.c_int(int(nqubits)),
       ctypes.c_int(int(normalize)))
    return state


def collapse_state(state, qubits, result, nqubits):
    return _collapse(state, qubits, result, nqubits, False)


def collapse_state_normalized(state, qubits, result, nqubits):
    return _collapse(state, qubits, result, nqubits, True)


def measure_frequencies(frequencies, probs, nshots, nqubits, seed, nthreads):
    probs = np.ascontiguousarray(probs)
    if probs.dtype == np.float64:
        suf = "c128"
    elif probs.dtype == np.float32:
        suf = "c64"
    else:
        raise TypeError("probs must be float32 or float64")
    freq64 = np.ascontiguousarray(frequencies, dtype=np.int64)
    fn = getattr(lib(), f"qjo_measure_frequencies_{suf}")
    fn(_ptr(freq64), _ptr(probs), ctypes.c_int64(int(nshots)), ctypes.c_int(int(nqubits)),
       ctypes.c_int64(int(seed)), ctypes.c_int(int(nthreads)))
    if freq64 is not frequencies:
        frequencies[...] = freq64
    return frequencies


def transpose_state(pieces, state, nqubits, order):
    _check(state)
    pieces = [np.ascontiguousarray(p, dtype=state.dtype) for p in pieces]
    arr = (ctypes.c_void_p * len(pieces))(*[p.ctypes.data for p in pieces])
    o = np.ascontiguousarray(np.asarray(order, dtype=np.int64))
    fn = getattr(lib(), f"qjo_transpose_state_{_suffix(state)}")
    fn(arr, ctypes.c_int(len(pieces)), _ptr(state), ctypes.c_int(int(nqubits)), _ptr(o))
    return state


def swap_pieces(piece0, piece1, new_global, nlocal):
    _check(piece0)
    _check(piece1)
    fn = getattr(lib(), f"qjo_swap_pieces_{_suffix(piece0)}")
    fn(_ptr(piece0), _ptr(piece1), ctypes.c_int(int(new_global)), ctypes.c_int(int(nlocal)))


def calculate_probabilities(state, qubits, nqubits):
    """qibo ``Backend.calculate_probabilities`` semantics (parity unpinned, see qj_oracle.c)."""
    _check(state)
    bits = np.ascontiguousarray([nqubits - q - 1 for q in qubits], dtype=np.int32)
    rdtype = np.float64 if state.dtype == np.complex128 else np.float32
    probs = np.zeros(1 << len(qubits), dtype=rdtype)
    fn = getattr(lib(), f"qjo_probabilities_{_suffix(state)}")
    fn(_ptr(state), ctypes.c_int(int(nqubits)), _ptr(bits), ctypes.c_int(len(qubits)), _ptr(probs))
    return probs


def mt_doubles(seed, count):
    out = np.zeros(count, dtype=np.float64)
    lib().qjo_mt_doubles(ctypes.c_int64(int(seed)), ctypes.c_int(count), _ptr(out))
    return out


def mt_randints(seed, n, count):
    out = np.zeros(count, dtype=np.int64)
    lib().qjo_mt_randints(ctypes.c_int64(int(seed)), ctypes.c_int64(int(n)), ctypes.c_int(count),
                          _ptr(out))
    return out
