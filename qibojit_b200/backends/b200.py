"""``B200Backend`` -- the reference's backend operator surface on top of the C-ABI library.

Method names, argument meaning, in-place/return-the-state behaviour and error types mirror
``NumbaBackend`` (/root/reference/src/qibojit/backends/cpu.py) and ``CupyBackend``
(/root/reference/src/qibojit/backends/gpu.py) for the state-vector path only:

* dispatch: ``GATE_OPS`` cpu.py:23-36, ``apply_gate`` :369-377, ``_apply_gate`` :433-450,
  ``_as_custom_matrix`` :519-539, ``_create_qubits_tensor`` :565-569,
  ``_one/_two/_multi_qubit_base`` :606-635 / :581-604, density-matrix doubling :452-517
* state ops: ``zero_state`` :334-353, ``_collapse_statevector`` :541-563 /
  ``collapse_state`` gpu.py:616-644, ``sample_frequencies`` cpu.py:383-394,
  ``calculate_probabilities`` / ``sample_shots`` gpu.py:751-778

States are ``torch`` CUDA tensors (torch is the device-memory / stream plumbing); every
kernel is a call into ``libqibojit_b200.so``.  There is no CPU path: constructing the
backend without a CUDA device raises.
"""

from collections import Counter

import numpy as np

from .. import _capi
from ..matrices import CustomMatrices

try:  # qibo is an un-vendored dependency of the reference; absent in this image
    from qibo.backends import Backend as _QiboBackend  # type: ignore
    from qibo.config import SHOT_METROPOLIS_THRESHOLD  # type: ignore
except Exception:  # pragma: no cover - exercised whenever qibo is missing
    _QiboBackend = object
    SHOT_METROPOLIS_THRESHOLD = 100000  # qibo 0.3.4 config value

GATE_OPS = {
    "X": "apply_x",
    "CNOT": "apply_x",
    "TOFFOLI": "apply_x",
    "Y": "apply_y",
    "Z": "apply_z",
    "CY": "apply_y",
    "CZ": "apply_z",
    "U1": "apply_z_pow",
    "CU1": "apply_z_pow",
    "SWAP": "apply_swap",
    "fSim": "apply_fsim",
    "GeneralizedfSim": "apply_fsim",
}

_DTYPE_TAG = {"complex64": _capi.QJ_C64, "complex128": _capi.QJ_C128}


def _torch():
    import torch  # imported lazily: keeps `import qibojit_b200` cheap on CPU-only hosts

    return torch


class B200Backend(_QiboBackend):
    MAX_NUM_TARGETS = 10  # QJ_MAX_TARGETS (reference GPU backend: 7, gpu.py:48)

    def __init__(self, device=None):
        if _QiboBackend is not object:
            super().__init__()
        torch = _torch()
        self._lib = _capi.load()
        if not torch.cuda.is_available():
            raise RuntimeError(
                "B200Backend needs a CUDA device: qibojit_b200 has no CPU fallback "
                "(use the reference's numba backend on CPU-only hosts)."
            )
        import psutil

        self.name = "qibojit_b200"
        self.platform = "b200"
        self.engine = torch
        self.dtype = "complex128"
        self.ngpus = torch.cuda.device_count()
        self.supports_multigpu = True
        self.tensor_types = (torch.Tensor, np.ndarray)
        self.numeric_types = (int, float, complex, np.int32, np.int64, np.float32, np.float64,
                              np.complex64, np.complex128)
        self.custom_matrices = CustomMatrices(self.dtype)
        self.matrices = self.custom_matrices
        self.versions = {"qibojit_b200": self._lib.qj_version().decode(), "torch": torch.__version__}
        self.nthreads = len(psutil.Process().cpu_affinity())
        self._handles = {}
        self._device_index = torch.cuda.current_device() if device is None else None
        self.device = f"/GPU:{self._device_index}"
        if device is not None:
            self.set_device(device)

    # ------------------------------------------------------------------ plumbing
    def _handle(self):
        h = self._handles.get(self._device_index)
        if h is None:
            torch = _torch()
            import ctypes

            out = ctypes.c_void_p()
            stream = torch.cuda.current_stream(self._device_index).cuda_stream
            _capi.check(self._lib.qj_create(self._device_index, ctypes.c_void_p(stream),
                                            ctypes.byref(out)))
            h = out
            self._handles[self._device_index] = h
        return h

    def _handle_or_none(self):
        """The current device's handle if it exists (never creates one: used on tear-down paths)."""
        return getattr(self, "_handles", {}).get(getattr(self, "_device_index", None))

    def __del__(self):
        try:
            handles, self._handles = self._handles, {}     # nobody can pick up a destroyed handle
            for h in handles.values():
                self._lib.qj_destroy(h)
        except Exception:
            pass

    @property
    def torch_device(self):
        return _torch().device("cuda", self._device_index)

    def _tag(self, state):
        tag = _DTYPE_TAG.get(str(state.dtype).replace("torch.", ""))
        if tag is None:
            raise TypeError(f"state dtype must be complex64 or complex128, got {state.dtype}")
        return tag

    def _np_dtype(self, state):
        return np.complex128 if self._tag(state) == _capi.QJ_C128 else np.complex64

    def _host_gate(self, gate, state):
        """Row-major host buffer of the gate in the state dtype (gpu.py:936, 1065)."""
        if hasattr(gate, "detach"):
            gate = gate.detach().cpu().numpy()
        return np.ascontiguousarray(np.asarray(gate, dtype=self._np_dtype(state)).ravel())

    def synchronize(self):
        _capi.check(self._lib.qj_sync(self._handle()))

    def launch_count(self):
        return int(self._lib.qj_launch_count(self._handle()))

    def set_route(self, route):
        """0 = automatic kernel choice, 1 = register kernels, 2 = tile kernel (tests/profiling)."""
        _capi.check(self._lib.qj_set_route(self._handle(), int(route)))

    # ------------------------------------------------------------------ setters (cpu.py:121-191)
    def set_dtype(self, dtype):
        dtype = str(dtype)
        if dtype not in _DTYPE_TAG:
            raise ValueError(f"Unsupported dtype {dtype}: use complex64 or complex128.")
        if dtype != self.dtype:
            self.dtype = dtype
            self.custom_matrices = CustomMatrices(dtype)
            self.matrices = self.custom_matrices

    def set_precision(self, precision):  # older qibo spelling
        self.set_dtype({"single": "complex64", "double": "complex128"}[precision])

    def set_device(self, device):
        torch = _torch()
        try:
            kind, idx = str(device).strip("/").split(":")
            idx = int(idx)
        except ValueError:
            raise ValueError(f"Unknown device {device}.") from None
        if kind.upper() != "GPU" or not 0 <= idx < torch.cuda.device_count():
            raise ValueError(
                f"Device {device} is not available for {self.name} ({self.platform}) backend."
            )
        self._device_index = idx
        self.device = f"/GPU:{idx}"

    def set_seed(self, seed):
        np.random.seed(seed)

    def set_threads(self, nthreads):
        """Number of Metropolis chains of the high-shot sampler (cpu.py:182-191, 391-393):
        pass the numba backend's thread count to reproduce its samples bit for bit."""
        if not isinstance(nthreads, int) or nthreads < 1:
            raise ValueError("nthreads must be a positive integer")
        self.nthreads = nthreads

    # ------------------------------------------------------------------ arrays
    def cast(self, array, dtype=None, copy=False):
        torch = _torch()
        if dtype is None:
            dtype = self.dtype
        tdtype = getattr(torch, str(dtype).replace("torch.", "")) if not isinstance(dtype, torch.dtype) else dtype
        if isinstance(array, torch.Tensor):
            out = array.to(device=self.torch_device, dtype=tdtype)
            return out.clone() if copy and out is array else out
        arr = np.asarray(array)
        return torch.as_tensor(arr, device=self.torch_device).to(tdtype)

    def to_numpy(self, array):
        if hasattr(array, "detach"):
            return array.detach().cpu().numpy()
        return np.asarray(array)

    def assert_allclose(self, value, target, rtol=1e-7, atol=0.0):
        np.testing.assert_allclose(self.to_numpy(value), self.to_numpy(target), rtol=rtol, atol=atol)

    # ------------------------------------------------------------------ state preparation
    def zero_state(self, nqubits, density_matrix=False, dtype=None):
        torch = _torch()
        if dtype is None:
            dtype = self.dtype
        n = 1 << nqubits
        total_qubits = 2 * nqubits if density_matrix else nqubits
        state = torch.empty(1 << total_qubits, dtype=getattr(torch, str(dtype)), device=self.torch_device)
        _capi.check(self._lib.qj_initial_state(self._handle(), state.data_ptr(), self._tag(state),
                                               total_qubits))
        return state.reshape(n, n) if density_matrix else state

    def plus_state(self, nqubits, density_matrix=False, dtype=None):
        torch = _torch()
        if dtype is None:
            dtype = self.dtype
        n = 1 << nqubits
        tdtype = getattr(torch, str(dtype))
        if density_matrix:
            return torch.full((n, n), 1.0 / n, dtype=tdtype, device=self.torch_device)
        return torch.full((n,), 1.0 / np.sqrt(n), dtype=tdtype, device=self.torch_device)

    # ------------------------------------------------------------------ gate application
    def apply_gate(self, gate, state, nqubits, inverse=False):
        if len(state.shape) == 2:
            return self._apply_gate_density_matrix(gate, state, nqubits, inverse)
        return self._apply_gate(gate, state, nqubits)

    def apply_gate_half_density_matrix(self, gate, state, nqubits):
        """First half of the doubling trick only: U acting on the row index of rho
        (qibo ``Backend.apply_gate_half_density_matrix``; pinned by the reference at
        tests/test_gates.py:414-422)."""
        matrix = self._as_custom_matrix(gate)
        qubits = self._create_qubits_tensor(gate, nqubits) + nqubits
        targets = gate.target_qubits
        shape = state.shape
        flat = state.reshape(-1)
        name = gate.__class__.__name__
        if len(targets) == 1:
            op = GATE_OPS.get(name, "apply_gate")
            flat = self._one_qubit_base(flat, 2 * nqubits, *targets, op, matrix, qubits)
        elif len(targets) == 2:
            op = GATE_OPS.get(name, "apply_two_qubit_gate")
            flat = self._two_qubit_base(flat, 2 * nqubits, *targets, op, matrix, qubits)
        else:
            flat = self._multi_qubit_base(flat, 2 * nqubits, targets, matrix, qubits)
        return flat.reshape(shape)

    def apply_channel(self, channel, state, nqubits):
        if len(state.shape) != 2:
            raise NotImplementedError("channels on state vectors are sampled by qibo, not the backend")
        # cpu.py:400-415
        all_unitary = getattr(channel, "_all_unitary_operators", False)
        state_copy = None if all_unitary else state.clone()
        new_state = (1 - channel.coefficient_sum) * state
        for coeff, gate in zip(channel.coefficients, channel.gates):
            state = self.apply_gate(gate, state, nqubits)
            new_state = new_state + coeff * state
            if all_unitary:
                state = self.apply_gate(gate, state, nqubits, inverse=True)
            else:
                state = state_copy.clone()
        return new_state

    def _apply_gate(self, gate, state, nqubits):
        if getattr(gate, "name", None) == "fanout":
            return self._apply_fanout_gate(gate, state, nqubits)
        matrix = self._as_custom_matrix(gate)
        qubits = self._create_qubits_tensor(gate, nqubits)
        targets = gate.target_qubits
        name = gate.__class__.__name__
        if len(targets) == 1:
            op = GATE_OPS.get(name, "apply_gate")
            return self._one_qubit_base(state, nqubits, *targets, op, matrix, qubits)
        if len(targets) == 2:
            op = GATE_OPS.get(name, "apply_two_qubit_gate")
            return self._two_qubit_base(state, nqubits, *targets, op, matrix, qubits)
        return self._multi_qubit_base(state, nqubits, targets, matrix, qubits)

    def _apply_fanout_gate(self, gate, state, nqubits):
        # cpu.py:417-431: a loop of CNOTs
        control = gate.control_qubits[0]
        for target in gate.target_qubits:
            qubits = np.array(sorted([nqubits - control - 1, nqubits - target - 1]), dtype=np.int32)
            state = self._one_qubit_base(state, nqubits, target, "apply_x", None, qubits)
        return state

    def _apply_gate_density_matrix(self, gate, state, nqubits, inverse=False):
        # cpu.py:452-498: rho flattened to a 2n-qubit vector, U on the row index then conj(U)
        # on the column index.
        name = gate.__class__.__name__
        if name in ("Y", "CY"):
            return self._apply_ygate_density_matrix(gate, state, nqubits)
        if name == "FanOut" or getattr(gate, "name", None) == "fanout":
            # a loop of CNOTs (cpu.py:417-431), each its own inverse and real: U rho U^dagger in two halves
            from .. import gates as G

            for t in gate.target_qubits:
                state = self._apply_gate_density_matrix(G.CNOT(gate.control_qubits[0], t), state, nqubits)
            return state
        if inverse:
            # cpu.py:464-468 inverts the gate's MATRIX (not the kernel-format buffer of U1 / fSim,
            # which is a scalar / 5-vector) and applies it with the general kernels
            from ..fusion import target_only_matrix

            matrix = np.linalg.inv(np.asarray(target_only_matrix(gate, self.custom_matrices), dtype=np.complex128))
            matrix = matrix.astype(self.dtype)
        else:
            matrix = self._as_custom_matrix(gate)
        qubits = self._create_qubits_tensor(gate, nqubits)
        qubits_dm = qubits + nqubits
        targets = gate.target_qubits
        targets_dm = tuple(q + nqubits for q in targets)
        shape = state.shape
        flat = state.reshape(-1)
        conj = None if matrix is None else np.conj(np.asarray(matrix))
        if len(targets) == 1:
            op = GATE_OPS.get(name, "apply_gate") if not inverse else "apply_gate"
            flat = self._one_qubit_base(flat, 2 * nqubits, *targets, op, matrix, qubits_dm)
            flat = self._one_qubit_base(flat, 2 * nqubits, *targets_dm, op, conj, qubits)
        elif len(targets) == 2:
            op = GATE_OPS.get(name, "apply_two_qubit_gate") if not inverse else "apply_two_qubit_gate"
            flat = self._two_qubit_base(flat, 2 * nqubits, *targets, op, matrix, qubits_dm)
            flat = self._two_qubit_base(flat, 2 * nqubits, *targets_dm, op, conj, qubits)
        else:
            flat = self._multi_qubit_base(flat, 2 * nqubits, targets, matrix, qubits_dm)
            flat = self._multi_qubit_base(flat, 2 * nqubits, targets_dm, conj, qubits)
        return flat.reshape(shape)

    def _apply_ygate_density_matrix(self, gate, state, nqubits):
        # cpu.py:500-517: the second half must use the general kernel so conj(Y) is applied
        matrix = self._as_custom_matrix(gate)
        qubits = self._create_qubits_tensor(gate, nqubits)
        qubits_dm = qubits + nqubits
        targets = gate.target_qubits
        targets_dm = tuple(q + nqubits for q in targets)
        shape = state.shape
        flat = state.reshape(-1)
        flat = self._one_qubit_base(flat, 2 * nqubits, *targets, "apply_y", matrix, qubits_dm)
        flat = self._one_qubit_base(flat, 2 * nqubits, *targets_dm, "apply_gate", np.conj(matrix), qubits)
        return flat.reshape(shape)

    def matrix(self, gate):
        """2^t x 2^t matrix of the gate's target part (what the channel inverse inverts)."""
        from ..fusion import target_only_matrix

        return np.asarray(target_only_matrix(gate, self.custom_matrices))

    def matrix_fused(self, fgate):
        """Dense matrix of a ``FusedGate`` block (qibo ``Backend.matrix_fused``, called at
        cpu.py:535-537)."""
        from ..fusion import fused_matrix

        return fused_matrix(fgate, self.custom_matrices)

    def _as_custom_matrix(self, gate):
        # cpu.py:519-539
        name = gate.__class__.__name__
        if name == "FusedGate":
            return self.matrix_fused(gate)
        if name == "FanOut":
            return None
        if hasattr(gate, "target_matrix"):
            return gate.target_matrix(self.custom_matrices)
        _matrix = getattr(self.custom_matrices, name)
        if getattr(gate, "parameters", ()):  # qibo ParametrizedGate
            return _matrix(*gate.parameters)
        return _matrix(2 ** len(gate.target_qubits)) if callable(_matrix) else _matrix

    def _create_qubits_tensor(self, gate, nqubits):
        # cpu.py:565-569
        qubits = [nqubits - q - 1 for q in gate.control_qubits]
        qubits.extend(nqubits - q - 1 for q in gate.target_qubits)
        return np.array(sorted(qubits), dtype=np.int32)

    @staticmethod
    def _qubits_arg(qubits):
        if qubits is None:
            return None, 0
        q = np.ascontiguousarray(np.asarray(qubits, dtype=np.int32))
        return q, int(q.size)

    def _one_qubit_base(self, state, nqubits, target, kernel, gate, qubits):
        # cpu.py:606-616 / gpu.py:1009-1036
        m = nqubits - target - 1
        q, nq = self._qubits_arg(qubits)
        qp = q.ctypes.data if q is not None else None
        h, tag, ptr = self._handle(), self._tag(state), state.data_ptr()
        lib = self._lib
        if kernel == "apply_gate":
            g = self._host_gate(gate, state)
            rc = lib.qj_apply_gate(h, ptr, tag, nqubits, m, g.ctypes.data, qp, nq)
        elif kernel == "apply_x":
            rc = lib.qj_apply_x(h, ptr, tag, nqubits, m, qp, nq)
        elif kernel == "apply_y":
            rc = lib.qj_apply_y(h, ptr, tag, nqubits, m, qp, nq)
        elif kernel == "apply_z":
            rc = lib.qj_apply_z(h, ptr, tag, nqubits, m, qp, nq)
        elif kernel == "apply_z_pow":
            g = self._host_gate(gate, state)
            rc = lib.qj_apply_z_pow(h, ptr, tag, nqubits, m, g.ctypes.data, qp, nq)
        else:
            raise ValueError(f"unknown one-qubit kernel {kernel}")
        _capi.check(rc)
        return state

    def _two_qubit_base(self, state, nqubits, target1, target2, kernel, gate, qubits):
        # cpu.py:618-635 / gpu.py:1038-1076
        if target1 > target2:
            swap_targets = 1
            m1, m2 = nqubits - target1 - 1, nqubits - target2 - 1
        else:
            swap_targets = 0
            m1, m2 = nqubits - target2 - 1, nqubits - target1 - 1
        q, nq = self._qubits_arg(qubits)
        qp = q.ctypes.data if q is not None else None
        h, tag, ptr = self._handle(), self._tag(state), state.data_ptr()
        lib = self._lib
        if kernel == "apply_two_qubit_gate":
            g = self._host_gate(gate, state)
            rc = lib.qj_apply_two_qubit_gate(h, ptr, tag, nqubits, m1, m2, swap_targets,
                                             g.ctypes.data, qp, nq)
        elif kernel == "apply_swap":
            rc = lib.qj_apply_swap(h, ptr, tag, nqubits, m1, m2, qp, nq)
        elif kernel == "apply_fsim":
            g = self._host_gate(gate, state)
            rc = lib.qj_apply_fsim(h, ptr, tag, nqubits, m1, m2, swap_targets, g.ctypes.data, qp, nq)
        else:
            raise ValueError(f"unknown two-qubit kernel {kernel}")
        _capi.check(rc)
        return state

    def _multi_qubit_base(self, state, nqubits, targets, gate, qubits):
        # cpu.py:581-604 / gpu.py:975-1007
        assert gate is not None
        if qubits is None:
            qubits = np.array(sorted(nqubits - q - 1 for q in targets), dtype=np.int32)
        ntargets = len(targets)
        if ntargets > self.MAX_NUM_TARGETS:
            raise ValueError(
                f"Number of target qubits must be <= {self.MAX_NUM_TARGETS} but is {ntargets}."
            )
        q, nq = self._qubits_arg(qubits)
        tmasks = np.array([1 << (nqubits - t - 1) for t in tuple(targets)[::-1]], dtype=np.int64)
        g = self._host_gate(gate, state)
        _capi.check(self._lib.qj_apply_multi_qubit_gate(
            self._handle(), state.data_ptr(), self._tag(state), nqubits, g.ctypes.data,
            q.ctypes.data, nq, tmasks.ctypes.data, ntargets))
        return state

    # ------------------------------------------------------------------ measurement
    def collapse_state(self, state, qubits, shot, nqubits, normalize=True, density_matrix=False):
        if density_matrix:
            return self._collapse_density_matrix(state, qubits, shot, nqubits, normalize)
        return self._collapse_statevector(state, qubits, shot, nqubits, normalize)

    def _collapse_density_matrix(self, state, qubits, shot, nqubits, normalize=True):
        """P rho P / tr(P rho) with P the projector on outcome `shot` of `qubits` (qibo
        ``collapse_density_matrix``): rho flattened to a 2n-qubit vector is collapsed on the row
        AND the column copy of the measured qubits by the same zeroing kernel as a state vector
        (ops.py:47-56 on 2n index bits); the normalisation is the trace, not a 2-norm."""
        qubits = sorted(int(q) for q in qubits)
        if hasattr(shot, "detach"):
            shot = shot.detach().cpu().numpy()
        shot = int(np.asarray(shot).flat[0]) if hasattr(shot, "shape") or hasattr(shot, "__len__") else int(shot)
        shape = state.shape
        flat = state.reshape(-1)
        both = qubits + [q + nqubits for q in qubits]
        flat = self._collapse_statevector(flat, both, (shot << len(qubits)) | shot, 2 * nqubits, normalize=False)
        state = flat.reshape(shape)
        if normalize:
            state /= state.diagonal().sum()
        return state

    def _collapse_statevector(self, state, qubits, shot, nqubits, normalize=True):
        # cpu.py:541-563
        bits = np.array([nqubits - q - 1 for q in reversed(list(qubits))], dtype=np.int32)
        if hasattr(shot, "detach"):
            shot = shot.detach().cpu().numpy()
        shot = (int(np.asarray(shot).flat[0]) if hasattr(shot, "shape") or hasattr(shot, "__len__")
                else int(shot))
        _capi.check(self._lib.qj_collapse_state(
            self._handle(), state.data_ptr(), self._tag(state), nqubits,
            bits.ctypes.data if bits.size else None, int(bits.size), shot, int(bool(normalize))))
        return state

    def calculate_norm(self, state, order=2):
        import ctypes

        if order != 2:
            raise NotImplementedError("only the 2-norm is provided by the kernel library")
        out = ctypes.c_double()
        flat = state.reshape(-1)
        nq = int(flat.numel()).bit_length() - 1
        _capi.check(self._lib.qj_norm2(self._handle(), flat.data_ptr(), self._tag(flat), nq,
                                       ctypes.byref(out)))
        return float(np.sqrt(out.value))

    def max_deviation(self, state, reference):
        """max_i |state[i] - reference| for a scalar `reference`, on the device without a
        state-sized temporary: the analytic check of QFT|0...0> = 2^(-n/2) at benchmark sizes."""
        import ctypes

        out = ctypes.c_double()
        flat = state.reshape(-1)
        nq = int(flat.numel()).bit_length() - 1
        ref = complex(reference)
        _capi.check(self._lib.qj_max_deviation(self._handle(), flat.data_ptr(), self._tag(flat), nq,
                                               ref.real, ref.imag, ctypes.byref(out)))
        return float(out.value)

    def calculate_probabilities(self, state, qubits, nqubits, density_matrix=False):
        torch = _torch()
        if density_matrix:
            # qibo `calculate_probabilities_density_matrix`: |marginal of the diagonal| in `qubits` order
            qubits = [int(q) for q in qubits]
            diag = state.diagonal().reshape((2,) * nqubits)
            rest = [q for q in range(nqubits) if q not in qubits]
            if rest:
                diag = diag.sum(dim=rest)
            kept = sorted(qubits)
            diag = diag.permute([kept.index(q) for q in qubits]) if len(qubits) > 1 else diag
            return diag.abs().reshape(-1)
        qubits = list(qubits)
        bits = np.array([nqubits - q - 1 for q in qubits], dtype=np.int32)
        rdtype = torch.float64 if self._tag(state) == _capi.QJ_C128 else torch.float32
        probs = torch.empty(1 << len(qubits), dtype=rdtype, device=state.device)
        _capi.check(self._lib.qj_calculate_probabilities(
            self._handle(), state.data_ptr(), self._tag(state), nqubits,
            bits.ctypes.data if bits.size else None, int(bits.size), probs.data_ptr()))
        return probs

    def _real_tag(self, probs):
        torch = _torch()
        if probs.dtype == torch.float64:
            return _capi.QJ_C128
        if probs.dtype == torch.float32:
            return _capi.QJ_C64
        raise TypeError("probabilities must be float32 or float64")

    def sample_shots(self, probabilities, nshots):
        """Inverse-CDF sampling on the device with uniforms from the host generator that
        ``set_seed`` seeds (numpy legacy ``RandomState``, as ``np.random.choice(p=...)`` uses)."""
        torch = _torch()
        probs = self.cast(probabilities, dtype=probabilities.dtype if hasattr(probabilities, "dtype") and str(probabilities.dtype).replace("torch.", "") in ("float32", "float64") else "float64")
        nq = int(probs.numel()).bit_length() - 1
        uniforms = np.ascontiguousarray(np.random.random_sample(int(nshots)))
        shots = torch.empty(int(nshots), dtype=torch.int64, device=probs.device)
        cdf = torch.empty(probs.numel(), dtype=torch.float64, device=probs.device)
        _capi.check(self._lib.qj_sample_shots(
            self._handle(), probs.data_ptr(), self._real_tag(probs), nq, uniforms.ctypes.data,
            int(nshots), shots.data_ptr(), cdf.data_ptr()))
        return shots

    def calculate_frequencies(self, samples):
        torch = _torch()
        samples = samples if hasattr(samples, "detach") else torch.as_tensor(np.asarray(samples))
        res, counts = torch.unique(samples, return_counts=True)
        return Counter(dict(zip(res.cpu().tolist(), counts.cpu().tolist())))

    def sample_frequencies(self, probabilities, nshots):
        # cpu.py:383-394
        torch = _torch()
        if nshots < SHOT_METROPOLIS_THRESHOLD:
            return self.calculate_frequencies(self.sample_shots(probabilities, nshots))
        seed = int(np.random.randint(0, int(1e8), size=1, dtype=np.int64)[0])
        probs = probabilities if hasattr(probabilities, "detach") else self.cast(probabilities, dtype=str(np.asarray(probabilities).dtype))
        nq = int(probs.numel()).bit_length() - 1
        freqs = torch.zeros(probs.numel(), dtype=torch.int64, device=probs.device)
        self.measure_frequencies_op(freqs, probs, nshots, nq, seed, self.nthreads)
        nz = torch.nonzero(freqs).reshape(-1)
        return Counter(dict(zip(nz.cpu().tolist(), freqs[nz].cpu().tolist())))

    def measure_frequencies_op(self, frequencies, probs, nshots, nqubits, seed, nthreads):
        """ops.measure_frequencies (ops.py:86-108) on device tensors."""
        _capi.check(self._lib.qj_measure_frequencies(
            self._handle(), frequencies.data_ptr(), probs.data_ptr(), self._real_tag(probs),
            int(nshots), int(nqubits), int(seed), int(nthreads)))
        return frequencies

    # ------------------------------------------------------------------ shard primitives
    # (used by qibojit_b200.distributed; the reference keeps pieces in host RAM and swaps them
    #  with ops.swap_pieces on the CPU, gpu.py:1497-1507)
    def shard_zeros(self, nlocal, dtype, one_at_zero=False):
        torch = _torch()
        if one_at_zero:
            return self.zero_state(nlocal, dtype=dtype)
        return torch.zeros(1 << nlocal, dtype=getattr(torch, str(dtype)), device=self.torch_device)

    def shard_spare(self, shard):
        """Another uninitialised buffer like `shard`, or None when the device has no room for it."""
        torch = _torch()
        try:
            return torch.empty_like(shard)
        except torch.OutOfMemoryError:
            return None

    def shard_from(self, piece, dtype):
        """This rank's piece of a caller-supplied state (host array or tensor on any device) as a
        device shard of its own (the caller keeps its state, cpu.py:96-119 semantics of `cast`)."""
        return self.cast(piece, dtype=str(dtype), copy=True).reshape(-1)

    def shard_reset(self, shard, nlocal, one_at_zero=False):
        if one_at_zero:
            _capi.check(self._lib.qj_initial_state(self._handle(), shard.data_ptr(), self._tag(shard), nlocal))
        else:
            shard.zero_()
        return shard

    def run_local_segment(self, shard, nlocal, segment):
        """Run the local gates between two exchanges: compiled once into multi-gate passes
        (``planner.Program``, cached on the segment), or gate by gate when programs are off."""
        if not getattr(self, "use_programs", True):
            for gate in segment.gates:
                shard = gate.apply(self, shard, nlocal)
            return shard
        if segment.compiled is None:
            from ..planner import Program

            segment.compiled = Program(self, segment.gates, nlocal,
                                       dtype=str(shard.dtype).replace("torch.", ""))
        return segment.compiled.run(shard)

    def _launch_geometry(self, handle, launch):
        import ctypes

        geom = (ctypes.c_int64 * 12)()
        _capi.check(self._lib.qj_program_launch_geometry(handle, launch, geom))
        return {"T": int(geom[0]), "r": int(geom[1]), "ntiles": int(geom[3]),
                "hibits": [int(v) for v in geom[4:12] if v >= 0]}

    @staticmethod
    def _tile_split(geom, nlocal, ntop):
        """How a launch's tiles split over the top `ntop` index bits: (free, tiles per block) where
        `free` are the positions (0 = lowest of the ntop bits) of the top bits OUTSIDE the tile -- the
        tiles whose number has the value v in its top len(free) bits are exactly the amplitudes whose
        free bits spell v, the tile range [v * per, (v + 1) * per) -- or None if the contiguous part of
        the tile reaches into the top bits."""
        if geom["r"] > nlocal - ntop:
            return None
        free = [i for i in range(ntop) if (nlocal - ntop + i) not in geom["hibits"]]
        if geom["ntiles"] % (1 << len(free)):
            return None
        return free, geom["ntiles"] >> len(free)

    def shard_scale(self, shard, nlocal, phase):
        ph = np.asarray(phase, dtype=self._np_dtype(shard)).reshape(1)
        _capi.check(self._lib.qj_apply_phase(self._handle(), shard.data_ptr(), self._tag(shard),
                                             nlocal, ph.ctypes.data))
        return shard

    def _staging(self, nelem, dtype, which):
        torch = _torch()
        key = (which, dtype)
        buf = getattr(self, "_stage_bufs", {}).get(key)
        if buf is None or buf.numel() < nelem:
            if not hasattr(self, "_stage_bufs"):
                self._stage_bufs = {}
            buf = torch.empty(nelem, dtype=dtype, device=self.torch_device)
            self._stage_bufs[key] = buf
        return buf[:nelem]

    # -- peer-memory transport (NVLink): every rank maps its peers' shards through CUDA IPC and a
    #    swap kernel trades sub-blocks in place; the NCCL transport below stays as the fallback
    #    (QJ_PEER_EXCHANGE=0, or IPC mapping not possible)
    def _peer_pointers(self, shard, comm):
        """{rank: device pointer to that rank's shard as seen from this process}, cached per shard."""
        import ctypes

        torch = _torch()
        # every rank's current shard address: a cached mapping is valid only while NONE of them moved
        # (states are re-allocated between circuit executions)
        mine = torch.tensor([int(shard.data_ptr())], dtype=torch.int64, device=self.torch_device)
        everyone = torch.empty(comm.world, dtype=torch.int64, device=self.torch_device)
        comm.dist.all_gather_into_tensor(everyone, mine, group=comm.group)
        key = tuple(everyone.cpu().tolist())
        cache = self.__dict__.setdefault("_peer_cache", {})
        if key in cache:
            return cache[key]
        handle = (ctypes.c_char * 64)()
        off = ctypes.c_int64()
        _capi.check(self._lib.qj_ipc_export(ctypes.c_void_p(shard.data_ptr()), handle, ctypes.byref(off)))
        gathered = [None] * comm.world
        comm.dist.all_gather_object(gathered, (bytes(handle.raw), int(off.value)), group=comm.group)
        opened = self.__dict__.setdefault("_ipc_opened", {})
        ptrs = {}
        for r, (hb, o) in enumerate(gathered):
            if r == comm.rank:
                ptrs[r] = int(shard.data_ptr())
                continue
            if hb not in opened:
                base = ctypes.c_void_p()
                buf = ctypes.create_string_buffer(hb, 64)
                _capi.check(self._lib.qj_ipc_open(buf, ctypes.byref(base)))
                opened[hb] = int(base.value)
            ptrs[r] = opened[hb] + o
        cache[key] = ptrs
        return ptrs

    def release_peer_mappings(self):
        """Unmap every peer shard (call on all ranks before the shards are freed: an allocation must
        not be released while another process still maps it)."""
        import ctypes

        for base in self.__dict__.pop("_ipc_opened", {}).values():
            self._lib.qj_ipc_close(ctypes.c_void_p(base))
        self.__dict__.pop("_peer_cache", None)
        self.__dict__.pop("_handshake_ptrs", None)

    def _peer_enabled(self, shard, comm):
        import os

        if os.environ.get("QJ_PEER_EXCHANGE", "1") == "0" or comm.world == 1:
            return False
        if getattr(self, "_peer_broken", False):
            return False
        try:
            self._peer_pointers(shard, comm)
            return True
        except Exception as exc:          # IPC export / mapping not possible on this system
            import sys

            self._peer_broken = True
            sys.stderr.write(f"qibojit_b200: peer-memory exchange unavailable ({exc}); using NCCL send/recv\n")
            return False

    def _stream_barrier(self, comm):
        """Cross-rank barrier in stream order: nothing enqueued after it on any rank starts before
        everything enqueued before it on every rank is done (no host synchronisation)."""
        torch = _torch()
        flag = self.__dict__.get("_barrier_flag")
        if flag is None:
            flag = self._barrier_flag = torch.zeros(1, dtype=torch.int32, device=self.torch_device)
        comm.dist.all_reduce(flag, group=comm.group)

    def _exchange_peer(self, shard, nlocal, lbits, pairs, comm):
        """pairs: [(peer rank, my sub-block value, its sub-block value)] -- one swap kernel per peer,
        this rank moves the half of the pair's amplitudes its rank order assigns to it."""
        ptrs = self._peer_pointers(shard, comm)
        bits = np.ascontiguousarray(np.asarray(lbits, dtype=np.int32))
        tag, h = self._tag(shard), self._handle()
        self._stream_barrier(comm)
        for peer, mine_val, peer_val in pairs:
            _capi.check(self._lib.qj_swap_bits_peer(h, shard.data_ptr(), ptrs[peer], tag, nlocal, bits.ctypes.data,
                                                    len(lbits), mine_val, peer_val, 0 if comm.rank < peer else 1, 2))
        self._stream_barrier(comm)

    def _peer_flags(self, comm):
        """(this rank's flag words, {rank: that rank's flag words as mapped here}) for
        `qj_peer_handshake`; slot r of every array is written by rank r only."""
        torch = _torch()
        flags = self.__dict__.get("_handshake_flags")
        if flags is None:
            flags = self._handshake_flags = torch.zeros(64, dtype=torch.int32, device=self.torch_device)
            self._handshake_epoch = {}
            torch.cuda.synchronize(self._device_index)
        ptrs = self.__dict__.get("_handshake_ptrs")
        if ptrs is None:
            ptrs = self._handshake_ptrs = self._peer_pointers(flags, comm)
        return flags, ptrs

    def _peer_handshake(self, comm, peers, timeout=30.0):
        """Stream-ordered rendezvous with `peers` (ranks) on the handle's current stream: what each side
        enqueued before it is complete and visible to the other before anything enqueued after it starts."""
        import ctypes

        flags, ptrs = self._peer_flags(comm)
        n = len(peers)
        slots = (ctypes.c_void_p * n)(*[ptrs[p] + 4 * comm.rank for p in peers])
        src = (ctypes.c_int32 * n)(*peers)
        epoch = self._handshake_epoch
        for p in peers:
            epoch[p] = (epoch.get(p, 0) + 1) & 0xFFFFFFFF
        eps = (ctypes.c_uint32 * n)(*[epoch[p] for p in peers])
        _capi.check(self._lib.qj_peer_handshake(self._handle(), ctypes.c_void_p(flags.data_ptr()), slots, src, eps,
                                                n, ctypes.c_double(timeout)))

    def run_segment_then_exchange(self, shard, nlocal, segment, lbits, rank_bits, rank, comm, chunk_bytes=1 << 29,
                                  spare=None):
        """A local segment followed by the exchange of `lbits` <-> `rank_bits`, with the LAST pass of
        the segment pipelined against the exchange.  With a `spare` buffer of the shard's size the
        exchange is out of place and uses no SM: the last pass runs block by block over the exchanged
        bits its tiles do not contain (XOR order: at step d this rank finishes the block that holds
        the sub-blocks of the ranks at distance d, which finish the block holding this rank's
        sub-block at the same step); as soon as a block is complete on both sides (a stream-ordered
        handshake over mapped flag words) the copy engines pull this rank's sub-blocks out of the
        peers' shards over NVLink into their final place in `spare`, while the pass works on the next
        block.  The sub-block that stays is written by the pass straight into `spare` (or copied, when
        its block also holds peers' sub-blocks).  Returns (spare, bytes sent): the caller keeps
        `shard` as its next spare.
        Falls back to "segment, then in-place exchange" (returning `shard`) without a spare or
        whenever the geometry does not allow it (exchanged bits not the shard's top bits, NCCL
        transport, gate-by-gate execution)."""
        import ctypes
        import os

        torch = _torch()
        k = len(lbits)
        esize = shard.element_size()
        moved = ((1 << k) - 1) * (1 << (nlocal - k)) * esize

        def plain():
            out = self.run_local_segment(shard, nlocal, segment)
            if k == 1:
                peer = rank ^ (1 << rank_bits[0])
                self.shard_exchange(out, nlocal, lbits[0], peer, (rank >> rank_bits[0]) & 1, comm, chunk_bytes)
            else:
                self.shard_exchange_multi(out, nlocal, lbits, rank_bits, rank, comm, chunk_bytes)
            return out, moved

        if (spare is None or os.environ.get("QJ_OVERLAP_EXCHANGE", "1") == "0"
                or not getattr(self, "overlap_exchange", True)
                or not getattr(self, "use_programs", True) or not self._peer_enabled(shard, comm)):
            return plain()
        if spare.dtype != shard.dtype or spare.numel() != shard.numel() or spare.data_ptr() == shard.data_ptr():
            raise ValueError("spare must be another buffer of the shard's size and dtype")
        if segment.compiled is None:
            from ..planner import Program

            segment.compiled = Program(self, segment.gates, nlocal, dtype=str(shard.dtype).replace("torch.", ""))
        prog = segment.compiled
        # what THIS rank's program allows: (eligible, exchanged bits outside the last launch's tile,
        # pieces per sub-block).  Ranks compile their own programs (gates controlled on global qubits
        # differ), so the protocol is fixed by agreement: identical bit sets everywhere, the coarsest
        # slicing, or the plain order on every rank.
        handle, last, geom, free, slices = None, -1, None, [], 1
        eligible = bool(prog.segments) and prog.segments[-1][0] == "program" and \
            list(lbits) == list(range(nlocal - k, nlocal))
        if eligible:
            handle = prog.segments[-1][1]
            nl = ctypes.c_int64()
            _capi.check(self._lib.qj_program_stats(handle, ctypes.byref(nl), None, None))
            last = int(nl.value) - 1
            eligible = last >= 0
        if eligible:
            geom = self._launch_geometry(handle, last)
            split = self._tile_split(geom, nlocal, k)
            eligible = split is not None
        if eligible:
            free = split[0]                     # exchanged bits outside the tile (positions in lbits)
            # a block goes piece by piece when it is a single sub-block: a piece is a tile range AND
            # must be one contiguous byte range (what the copy engine moves), i.e. the index bits right
            # below the exchanged ones must lie outside the launch's tile as well
            want = max(1, int(os.environ.get("QJ_OVERLAP_SLICES", "8")))
            while len(free) == k and slices * 2 <= want:
                finer = self._tile_split(geom, nlocal, k + slices.bit_length())
                if finer is None or len(finer[0]) != k + slices.bit_length():
                    break
                slices *= 2
        agreed = getattr(segment, "_overlap_agreed", None)
        agreed = agreed[1] if agreed is not None and agreed[0] is prog else None     # (per compiled program)
        if agreed is None:
            mask = sum(1 << i for i in free)
            mine_desc = [int(eligible), -int(eligible), mask, -mask, slices]
            t = torch.tensor(mine_desc, dtype=torch.int64, device=self.torch_device)
            comm.dist.all_reduce(t, op=comm.dist.ReduceOp.MIN, group=comm.group)
            lo_e, hi_e, lo_m, hi_m, min_slices = [int(v) for v in t.cpu().tolist()]
            agreed = (lo_e == 1 and -hi_e == 1 and lo_m == -hi_m, int(min_slices))
            segment._overlap_agreed = (prog, agreed)
        if not agreed[0]:
            return plain()
        slices = agreed[1]
        nfree = len(free)
        per_block = geom["ntiles"] >> nfree
        inside = [i for i in range(k) if i not in free]

        self.overlapped_exchanges = getattr(self, "overlapped_exchanges", 0) + 1
        self.overlapped_inside_tile = getattr(self, "overlapped_inside_tile", 0) + (1 if inside else 0)
        h = self._handle()
        ptr, out = shard.data_ptr(), spare.data_ptr()
        src_ptrs = self._peer_pointers(shard, comm)       # the peers' shards (the sources of the pulls)
        self._peer_flags(comm)
        # everything but the last launch
        for seg in prog.segments[:-1]:
            if seg[0] == "program":
                _capi.check(self._lib.qj_program_run(h, seg[1], ptr))
            else:
                seg[1].apply(self, shard, nlocal)
        if last > 0:
            _capi.check(self._lib.qj_program_run_ex(h, handle, ptr, 0, last, 0))
        main = torch.cuda.current_stream(self._device_index)
        side = self.__dict__.get("_side_stream")
        if side is None:
            side = self._side_stream = torch.cuda.Stream(device=self._device_index, priority=-1)
        side_ptr, main_ptr = ctypes.c_void_p(side.cuda_stream), ctypes.c_void_p(main.cuda_stream)
        mine = sum(((rank >> j) & 1) << i for i, j in enumerate(rank_bits))
        sub_bytes = (1 << (nlocal - k)) * esize

        def rank_of(a):
            r = rank
            for i, j in enumerate(rank_bits):
                r = (r & ~(1 << j)) | (((a >> i) & 1) << j)
            return r

        def block_of(a):                                   # a sub-block's block: its free bits, packed
            return sum(((a >> i) & 1) << j for j, i in enumerate(free))

        def on_side(fn):
            _capi.check(self._lib.qj_set_stream(h, side_ptr))
            try:
                fn()
            finally:
                _capi.check(self._lib.qj_set_stream(h, main_ptr))

        def after_main():
            done = torch.cuda.Event()
            done.record(main)
            side.wait_event(done)

        def pull(a, off, nbytes):
            _capi.check(self._lib.qj_copy_async(h, ctypes.c_void_p(out + a * sub_bytes + off),
                                                ctypes.c_void_p(src_ptrs[rank_of(a)] + mine * sub_bytes + off), nbytes))

        trace = os.environ.get("QJ_OVERLAP_TRACE") == "1"
        if trace:
            t_begin, t_passes, t_end = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            t_begin.record(main)
        my_block = block_of(mine)
        everyone = []
        for d in list(range(1, 1 << nfree)) + [0]:
            block = my_block ^ d
            # (same XOR offsets on every rank, in the same order: each pull of a step is a perfect
            # matching over the ranks -- nobody's links serve two readers while a sibling's idle)
            partners = [mine ^ x for x in range(1, 1 << k) if block_of(x) == d]
            everyone += partners
            if not partners:
                # the block is this rank's own sub-block: from the shard straight into the new buffer
                _capi.check(self._lib.qj_program_run_tiles_to(h, handle, ptr, ctypes.c_void_p(out), last,
                                                              block * per_block, per_block))
                continue
            piece = per_block // slices
            for s in range(slices):
                _capi.check(self._lib.qj_program_run_tiles(h, handle, ptr, last, block * per_block + s * piece, piece))
                after_main()

                def step(s=s):
                    # both sides finished this piece of what they trade
                    self._peer_handshake(comm, [rank_of(a) for a in partners])
                    for a in partners:
                        pull(a, s * (sub_bytes // slices), sub_bytes // slices)
                    if d == 0 and s == slices - 1:
                        # the block also held the sub-block that stays
                        _capi.check(self._lib.qj_copy_async(h, ctypes.c_void_p(out + mine * sub_bytes),
                                                            ctypes.c_void_p(ptr + mine * sub_bytes), sub_bytes))
                on_side(step)
        if trace:
            t_passes.record(main)
        on_side(lambda: self._peer_handshake(comm, [rank_of(a) for a in everyone]))   # every pull out of this shard is done
        fin = torch.cuda.Event()
        fin.record(side)
        main.wait_event(fin)
        if trace:
            t_end.record(main)
            t_end.synchronize()
            self.overlap_trace = {"last_pass_ms": t_begin.elapsed_time(t_passes), "total_ms": t_begin.elapsed_time(t_end),
                                  "blocks": 1 << nfree, "slices": slices, "inside_bits": len(inside)}
        return spare, moved

    def shard_exchange(self, shard, nlocal, lbit, peer, is_upper, comm, chunk_bytes=1 << 29):
        """Global<->local qubit swap with rank `peer` (ops.swap_pieces semantics): the amplitudes
        of this shard whose local bit `lbit` equals (1 - is_upper) are exchanged with the peer's
        complementary half, chunk by chunk through a small staging buffer (nothing state-sized is
        allocated).  When `lbit` is the top local bit both halves are contiguous and travel
        straight from / into the shard; otherwise they are packed / unpacked by the library's
        strided-copy kernels.  Returns the bytes sent."""
        dist = comm.dist
        half = 1 << (nlocal - 1)
        esize = shard.element_size()
        if self._peer_enabled(shard, comm):
            self._exchange_peer(shard, nlocal, [lbit], [(peer, 1 - int(is_upper), int(is_upper))], comm)
            return half * esize
        chunk = max(2, min(half, chunk_bytes // esize))
        contiguous = lbit == nlocal - 1
        tag = self._tag(shard)
        h = self._handle()
        region = None
        if contiguous:
            region = shard[:half] if is_upper else shard[half:]
        for c0 in range(0, half, chunk):
            n = min(chunk, half - c0)
            recv = self._staging(n, shard.dtype, "recv")
            if contiguous:
                send = region[c0:c0 + n]
            else:
                send = self._staging(n, shard.dtype, "send")
                _capi.check(self._lib.qj_swap_pack(h, shard.data_ptr(), send.data_ptr(), tag, nlocal,
                                                   lbit, int(is_upper), c0, n))
            # NCCL moves real pairs; the complex views share storage
            torch = _torch()
            ops = [dist.P2POp(dist.isend, torch.view_as_real(send), peer, group=comm.group),
                   dist.P2POp(dist.irecv, torch.view_as_real(recv), peer, group=comm.group)]
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            if contiguous:
                region[c0:c0 + n].copy_(recv)
            else:
                _capi.check(self._lib.qj_swap_unpack(h, shard.data_ptr(), recv.data_ptr(), tag, nlocal,
                                                     lbit, int(is_upper), c0, n))
        return half * esize

    def shard_exchange_multi(self, shard, nlocal, lbits, rank_bits, rank, comm, chunk_bytes=1 << 29):
        """Swap k global qubits (rank bits `rank_bits`) with k local ones (index bits `lbits`,
        ascending, paired in order) in ONE all-to-all among the 2^k ranks that differ in those
        rank bits: this rank sends the sub-block whose local bits spell `a` to the rank whose
        exchanged rank bits spell `a`, and stores what that rank sends in the same slots.
        (2^k - 1) / 2^k of the shard crosses the links instead of k / 2.  Chunked through a
        bounded staging buffer.  Returns the bytes sent."""
        dist = comm.dist
        torch = _torch()
        k = len(lbits)
        mine = sum(((rank >> j) & 1) << i for i, j in enumerate(rank_bits))
        others = [a for a in range(1 << k) if a != mine]
        peers = {}
        for a in others:
            r = rank
            for i, j in enumerate(rank_bits):
                r = (r & ~(1 << j)) | (((a >> i) & 1) << j)
            peers[a] = r
        sub = 1 << (nlocal - k)
        esize = shard.element_size()
        if self._peer_enabled(shard, comm):
            # pairs meet in XOR-distance order: at step d this rank and its partner swap with each other
            order = [mine ^ d for d in range(1, 1 << k)]
            self._exchange_peer(shard, nlocal, list(lbits), [(peers[a], a, mine) for a in order], comm)
            return len(others) * sub * esize
        # two chunk slots: chunk c + 1 is packed (and its transfers queued) while chunk c is on
        # the links, chunk c is unpacked while chunk c + 1 travels
        chunk = max(2, min(sub, (chunk_bytes // esize // len(others) // 2) & ~1023 or 2))   # whole 16-byte vectors
        tag = self._tag(shard)
        h = self._handle()
        bits = np.ascontiguousarray(np.asarray(lbits, dtype=np.int32))

        def launch(c0, slot):
            n = min(chunk, sub - c0)
            ops, bufs = [], []
            for a in others:
                send = self._staging(n, shard.dtype, ("send", a, slot))
                recv = self._staging(n, shard.dtype, ("recv", a, slot))
                _capi.check(self._lib.qj_swap_pack_bits(h, shard.data_ptr(), send.data_ptr(), tag, nlocal,
                                                        bits.ctypes.data, k, a, c0, n))
                ops.append(dist.P2POp(dist.isend, torch.view_as_real(send), peers[a], group=comm.group))
                ops.append(dist.P2POp(dist.irecv, torch.view_as_real(recv), peers[a], group=comm.group))
                bufs.append((a, recv))
            return c0, n, dist.batch_isend_irecv(ops), bufs

        def finish(job):
            c0, n, reqs, bufs = job
            for req in reqs:
                req.wait()
            for a, recv in bufs:
                _capi.check(self._lib.qj_swap_unpack_bits(h, shard.data_ptr(), recv.data_ptr(), tag, nlocal,
                                                          bits.ctypes.data, k, a, c0, n))

        pending = None
        for i, c0 in enumerate(range(0, sub, chunk)):
            job = launch(c0, i & 1)
            if pending is not None:
                finish(pending)
            pending = job
        if pending is not None:
            finish(pending)
        return len(others) * sub * esize

    # ------------------------------------------------------------------ circuits
    def execute_circuit(self, circuit, initial_state=None, nshots=None):
        """Stand-in for qibo's ``Backend.execute_circuit``: zero state (or a
        cast copy of `initial_state`), then the circuit's queue.  The queue is compiled once per
        (circuit, dtype) into multi-gate passes (``planner.Program``) so the whole circuit costs a
        handful of passes over the state instead of one per gate; ``self.use_programs = False``
        restores the gate-by-gate loop of the reference."""
        nqubits = circuit.nqubits
        programs = getattr(self, "use_programs", True) and nqubits >= 4 and len(circuit.queue) > 1
        if initial_state is None and programs:
            # from |0...0> the SWAP gates of the circuit are relabellings (planner.relabel_swaps_away)
            # and the state preparation is fused into the first pass: the state starts as
            # uninitialised memory that the first pass only writes
            torch = _torch()
            state = torch.empty(1 << nqubits, dtype=getattr(torch, str(self.dtype)), device=self.torch_device)
            return self.compile_circuit(circuit, zero_state=True).run(state, from_zero=True)
        if initial_state is None:
            state = self.zero_state(nqubits)
        else:
            state = self.cast(initial_state, copy=True)
        if programs:
            return self.compile_circuit(circuit, zero_state=False).run(state)
        for gate in circuit.queue:
            state = gate.apply(self, state, nqubits)
        return state

    @staticmethod
    def circuit_fingerprint(queue):
        """Comparable summary of a gate queue -- classes, qubits AND parameter values -- so that a
        cached program is not reused after `circuit.set_parameters(...)` (the matrices are baked
        into the program image).  The summary is the tuple itself, compared by value: Python's
        hash() is no identity (hash(-1.0) == hash(-2.0)), so nothing here goes through it; arrays
        enter as a blake2b digest of their bytes."""
        import hashlib

        def param(p):
            if hasattr(p, "tobytes"):
                a = np.ascontiguousarray(p)
                return (a.shape, str(a.dtype), hashlib.blake2b(a.tobytes(), digest_size=16).digest())
            if isinstance(p, (list, tuple)):
                return tuple(param(x) for x in p)
            if isinstance(p, (bool, int, float, complex, str, bytes, type(None))):
                return (type(p).__name__, p)
            if isinstance(p, np.generic):
                return (p.dtype.str, p.tobytes())
            return repr(p)

        out = []
        for g in queue:
            item = (g.__class__.__name__, tuple(g.target_qubits), tuple(g.control_qubits),
                    tuple(param(x) for x in getattr(g, "parameters", ())))
            if hasattr(g, "gates"):               # FusedGate
                item += (B200Backend.circuit_fingerprint(g.gates),)
            out.append(item)
        return tuple(out)

    def compile_circuit(self, circuit, **options):
        """Compile (and cache on the circuit object) the multi-gate pass program of `circuit`."""
        from ..planner import Program

        key = (self.dtype, self._device_index, tuple(sorted(options.items())))
        fingerprint = (len(circuit.queue), self.circuit_fingerprint(circuit.queue))
        cache = circuit.__dict__.setdefault("_qj_programs", {})
        entry = cache.get(key)
        if entry is not None and entry[0] != fingerprint:
            entry[1].close()          # re-parametrised circuit: one program per key, no pile-up
            entry = None
        if entry is None:
            entry = (fingerprint, Program(self, circuit.queue, circuit.nqubits, dtype=self.dtype, **options))
            cache[key] = entry
        return entry[1]

    def execute_distributed_circuit(self, circuit, initial_state=None, nshots=None):
        from ..distributed import execute_distributed_circuit

        return execute_distributed_circuit(self, circuit, initial_state, nshots)
