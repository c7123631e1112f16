"""ctypes binding of include/qibojit_b200.h.

The library is the product: if it is missing, cannot be built, or no CUDA device is
present, calls raise -- there is no CPU fallback anywhere in this package.
"""

import ctypes
import os

from . import build as _build

QJ_C64, QJ_C128 = 0, 1
QJ_ERR_INVALID, QJ_ERR_CUDA, QJ_ERR_NODEVICE, QJ_ERR_UNSUPPORTED = -1, -2, -3, -4

_c = ctypes
_P = _c.c_void_p
_I = _c.c_int
_L = _c.c_int64

# name -> (restype, argtypes); mirrors include/qibojit_b200.h one to one
SIGNATURES = {
    "qj_create": (_I, [_I, _P, _c.POINTER(_P)]),
    "qj_destroy": (_I, [_P]),
    "qj_set_stream": (_I, [_P, _P]),
    "qj_sync": (_I, [_P]),
    "qj_last_error": (_c.c_char_p, []),
    "qj_version": (_c.c_char_p, []),
    "qj_launch_count": (_L, [_P]),
    "qj_set_route": (_I, [_P, _I]),
    "qj_initial_state": (_I, [_P, _P, _I, _I]),
    "qj_apply_gate": (_I, [_P, _P, _I, _I, _I, _P, _P, _I]),
    "qj_apply_x": (_I, [_P, _P, _I, _I, _I, _P, _I]),
    "qj_apply_y": (_I, [_P, _P, _I, _I, _I, _P, _I]),
    "qj_apply_z": (_I, [_P, _P, _I, _I, _I, _P, _I]),
    "qj_apply_z_pow": (_I, [_P, _P, _I, _I, _I, _P, _P, _I]),
    "qj_apply_phase": (_I, [_P, _P, _I, _I, _P]),
    "qj_apply_two_qubit_gate": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _I]),
    "qj_apply_swap": (_I, [_P, _P, _I, _I, _I, _I, _P, _I]),
    "qj_apply_fsim": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _I]),
    "qj_apply_multi_qubit_gate": (_I, [_P, _P, _I, _I, _P, _P, _I, _P, _I]),
    "qj_collapse_state": (_I, [_P, _P, _I, _I, _P, _I, _L, _I]),
    "qj_norm2": (_I, [_P, _P, _I, _I, _c.POINTER(_c.c_double)]),
    "qj_max_deviation": (_I, [_P, _P, _I, _I, _c.c_double, _c.c_double, _c.POINTER(_c.c_double)]),
    "qj_calculate_probabilities": (_I, [_P, _P, _I, _I, _P, _I, _P]),
    "qj_measure_frequencies": (_I, [_P, _P, _P, _I, _L, _I, _L, _I]),
    "qj_sample_shots": (_I, [_P, _P, _I, _I, _P, _L, _P, _P]),
    "qj_swap_pieces_peer": (_I, [_P, _P, _P, _I, _I, _I, _I]),
    "qj_swap_bits_peer": (_I, [_P, _P, _P, _I, _I, _P, _I, _I, _I, _I, _I]),
    "qj_ipc_export": (_I, [_P, _P, _c.POINTER(_L)]),
    "qj_ipc_open": (_I, [_P, _c.POINTER(_P)]),
    "qj_ipc_close": (_I, [_P]),
    "qj_peer_handshake": (_I, [_P, _P, _P, _P, _P, _I, _c.c_double]),
    "qj_copy_async": (_I, [_P, _P, _P, _L]),
    "qj_swap_pack": (_I, [_P, _P, _P, _I, _I, _I, _I, _L, _L]),
    "qj_swap_unpack": (_I, [_P, _P, _P, _I, _I, _I, _I, _L, _L]),
    "qj_swap_pack_bits": (_I, [_P, _P, _P, _I, _I, _P, _I, _I, _L, _L]),
    "qj_swap_unpack_bits": (_I, [_P, _P, _P, _I, _I, _P, _I, _I, _L, _L]),
    "qj_program_create": (_I, [_P, _I, _I, _P, _I, _P, _L, _P, _L, _P, _L, _c.POINTER(_P)]),
    "qj_program_run": (_I, [_P, _P, _P]),
    "qj_program_run_launch": (_I, [_P, _P, _P, _I]),
    "qj_program_run_ex": (_I, [_P, _P, _P, _I, _I, _I]),
    "qj_program_run_tiles": (_I, [_P, _P, _P, _I, _L, _L]),
    "qj_program_run_tiles_to": (_I, [_P, _P, _P, _P, _I, _L, _L]),
    "qj_program_launch_geometry": (_I, [_P, _I, _P]),
    "qj_program_stats": (_I, [_P, _c.POINTER(_L), _c.POINTER(_L), _c.POINTER(_L)]),
    "qj_program_destroy": (_I, [_P, _P]),
    "qj_program_encode": (_I, [_I, _I, _P, _I, _P, _L, _P, _L, _P, _L, _c.POINTER(_P)]),
    "qj_program_image_sizes": (_I, [_P, _c.POINTER(_L), _c.POINTER(_L), _c.POINTER(_L)]),
    "qj_program_image_read": (_I, [_P, _P, _P, _P]),
    "qj_program_image_destroy": (_I, [_P]),
}

QJ_OPK_DENSE1, QJ_OPK_DENSE2, QJ_OPK_DIAG = 1, 2, 3
QJ_RUN_ZERO_INPUT = 1
QJ_MAX_DIAG_BITS, QJ_MAX_LOCAL_BITS, QJ_MAX_QUBITS, QJ_MAX_REG_BITS = 12, 16, 48, 8
QJ_LAUNCH_INFO_FIELDS = 16

_lib = None


class QjError(RuntimeError):
    pass


def library_path():
    return _build.LIB_PATH


def load(build_if_missing=True):
    """dlopen the C-ABI library (building it in-tree first when nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path) or (build_if_missing and _build.is_stale()):
        if not build_if_missing:
            raise ImportError(f"{path} is missing: run `python -m qibojit_b200.build`")
        try:
            _build.build()
        except Exception as exc:  # stale-but-present library is still usable on a GPU box
            if not os.path.exists(path):
                raise ImportError(f"cannot build {path}: {exc}") from exc
    lib = _c.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library diverge
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc == 0:
        return
    msg = load().qj_last_error().decode()
    if rc == QJ_ERR_INVALID:
        raise ValueError(msg)
    if rc == QJ_ERR_NODEVICE:
        raise RuntimeError(msg)
    if rc == QJ_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise QjError(msg)
