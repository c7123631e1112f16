"""qibojit_b200 -- B200-native (sm_100a) drop-in for qibojit's state-vector gate path.

``qibo.set_backend("qibojit_b200", platform="b200")`` resolves through ``MetaBackend``
exactly like the reference package does (/root/reference/src/qibojit/__init__.py:1-5).
"""

__version__ = "0.1.0"

from qibojit_b200.backends import MetaBackend  # noqa: E402,F401
