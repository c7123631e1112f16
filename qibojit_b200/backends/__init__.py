"""Plugin entry point (mirrors /root/reference/src/qibojit/backends/__init__.py:8-46)."""

PLATFORMS = ("b200",)


class MetaBackend:
    """Meta-backend class which takes care of loading the qibojit_b200 backend."""

    @staticmethod
    def load(platform: str = None):
        """Load the backend for ``platform`` (default and only platform: ``"b200"``)."""
        if platform is None:
            platform = "b200"
        if platform == "b200":
            from qibojit_b200.backends.b200 import B200Backend

            return B200Backend()
        raise ValueError(
            f"Unsupported platform {platform}, please use one of the following: {PLATFORMS}."
        )

    def list_available(self) -> dict:
        """Which platforms can be constructed here (needs the library and a CUDA device)."""
        available = {}
        for platform in PLATFORMS:
            try:
                MetaBackend.load(platform=platform)
                available[platform] = True
            except Exception:  # ImportError / RuntimeError when no device or no library
                available[platform] = False
        return available
