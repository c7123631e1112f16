"""In-tree nvcc build of the C-ABI library (sm_100a only)."""

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libqibojit_b200.so")
SOURCES = ["pass_kernels.cu", "pass_kernels_f32.cu", "capi.cu", "gate_kernels.cu", "tile_kernels.cu", "ops_kernels.cu"]
HEADERS = ["common.cuh", "pass_device.cuh", os.path.join("..", "..", "include", "qibojit_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-diag-suppress", "128",
]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libqibojit_b200.so")
    return nvcc


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > built for d in deps if os.path.exists(d))


OBJ_CACHE = os.path.join(LIB_DIR, "objcache")


def _object_key(src, extra):
    """Content hash of everything one translation unit's object depends on: its source, the
    headers, the flags and the compiler."""
    import hashlib

    h = hashlib.sha256()
    ver = subprocess.run([_nvcc(), "--version"], capture_output=True, text=True).stdout
    h.update(ver.encode())
    h.update(" ".join(NVCC_FLAGS + extra).encode())
    for name in [src] + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read() + b"\0")
    return h.hexdigest()[:20]


def build(force=False, verbose=False, use_cache=True):
    """Compile qibojit_b200/csrc/*.cu into qibojit_b200/lib/libqibojit_b200.so: one nvcc per
    translation unit, all at once, then one link.

    nvcc's optimiser needs about 8 minutes for the complex64 instantiation of the pass kernel
    (pass_kernels_f32.cu; superlinear in its 32 unrolled register elements -- the complex128 one takes
    18 s, and faster-compiling spellings of the kernel measured 8 % slower at run time), so objects are
    kept in lib/objcache/ under a content hash of their inputs (source, headers, flags, nvcc version)
    and a unit whose inputs are byte-identical is not compiled again.  `force` re-links from current
    objects even if the library looks fresh; `use_cache=False` compiles every unit from scratch."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(OBJ_CACHE, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []
    procs, objs = [], []
    for src in SOURCES:
        stem = src.replace(".cu", "")
        obj = os.path.join(OBJ_CACHE, f"{stem}.{_object_key(src, extra)}.o")
        objs.append(obj)
        if use_cache and not verbose and os.path.exists(obj):
            continue
        tmp = obj + ".tmp"
        cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", tmp]
        procs.append((src, obj, tmp, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    logs, failed = [], []
    for src, obj, tmp, proc in procs:
        out = proc.communicate()[0]
        logs.append(out)
        if proc.returncode != 0:
            failed.append(f"{src}:\n{out}")
            continue
        os.replace(tmp, obj)
        stem = src.replace(".cu", "")
        for old in os.listdir(OBJ_CACHE):          # one object per unit
            if old.startswith(stem + ".") and old.endswith(".o") and os.path.join(OBJ_CACHE, old) != obj:
                os.remove(os.path.join(OBJ_CACHE, old))
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(failed))
    link = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    if verbose:
        print("\n".join(logs))
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, use_cache="--no-cache" not in sys.argv))
