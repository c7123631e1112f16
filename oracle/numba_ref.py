"""TEST INFRASTRUCTURE ONLY -- the reference's own numba kernels as a CPU arm.

Loads ``qibojit.custom_operators.gates`` / ``ops`` (the unmodified reference modules placed under
``oracle/_ref`` by ``oracle/make_ref.py``) through a package stub, exactly as
``tests/golden/make_golden.py`` does from /root/reference, and sets the numba thread count the way
the reference backend does (``len(psutil.Process().cpu_affinity())``, backends/cpu.py:86-89,188-191).
Only bench.py's ``cpu_baseline`` / ``--impl reference`` legs and tests/ may import this module.
"""

import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref", "qibojit")

_modules = None


def host_threads():
    """The thread count the reference backend picks (cpu.py:86-89)."""
    try:
        import psutil

        return len(psutil.Process().cpu_affinity())
    except Exception:
        return os.cpu_count() or 1


def load(nthreads=None):
    """-> (gates module, ops module, threads in use); raises ImportError when oracle/_ref is absent."""
    global _modules
    if _modules is None:
        from . import make_ref

        if not make_ref.available() and not make_ref.build():
            raise ImportError("oracle/_ref is missing (run `python oracle/make_ref.py` where /root/reference exists)")
        n = int(nthreads or host_threads())
        # torchrun exports OMP_NUM_THREADS=1; the reference sizes its pool from the CPU affinity
        os.environ["NUMBA_NUM_THREADS"] = str(n)
        os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join("/tmp", f"numba_cache_{os.getuid()}"))
        if "qibojit" not in sys.modules:
            pkg = types.ModuleType("qibojit")
            pkg.__path__ = [REF_ROOT]
            sys.modules["qibojit"] = pkg
        import numba
        import qibojit.custom_operators.gates as G
        import qibojit.custom_operators.ops as O

        numba.set_num_threads(min(n, numba.config.NUMBA_NUM_THREADS))
        _modules = (G, O, numba.get_num_threads())
    return _modules
