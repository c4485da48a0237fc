"""In-tree build of libecdna_b200.so (nvcc, sm_100a only) and of the C++ host CLI."""
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "libecdna_b200.so")
CLI_PATH = os.path.join(PKG_DIR, "host", "ecdna")
CSRC = os.path.join(PKG_DIR, "csrc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-cudart", "static", "-diag-suppress", "128",
              # every fused multiply-add in this library is written explicitly (the CPU oracle is compiled
              # with -ffp-contract=off and must see the same roundings)
              "-fmad=false"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libecdna_b200.so cannot be built (there is no CPU fallback)")


def _digest(sources, extra=""):
    import hashlib
    h = hashlib.sha256(extra.encode())
    for s in sorted(sources):
        h.update(os.path.basename(s).encode())
        with open(s, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale(target, sources, extra=""):
    """Stale = built from different source CONTENT (mtimes do not survive copying the tree)."""
    stamp = target + ".srchash"
    if not os.path.exists(target) or not os.path.exists(stamp):
        return True
    return open(stamp).read().strip() != _digest(sources, extra)


def _stamp(target, sources, extra=""):
    with open(target + ".srchash", "w") as f:
        f.write(_digest(sources, extra))


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libecdna_b200.so and host/*.cpp into the `ecdna` CLI."""
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(ROOT, "include", "ecdna_b200.h")]
    flags = " ".join(NVCC_FLAGS)
    if force or _stale(LIB_PATH, srcs, flags):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
            "-o", LIB_PATH, os.path.join(CSRC, "capi.cu")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        _stamp(LIB_PATH, srcs, flags)
        if verbose:
            print(r.stderr)
    host_dir = os.path.join(PKG_DIR, "host")
    host_srcs = [os.path.join(host_dir, f) for f in sorted(os.listdir(host_dir)) if f.endswith((".cpp", ".h"))]
    hdr = [os.path.join(ROOT, "include", "ecdna_b200.h")]
    if host_srcs and (force or _stale(CLI_PATH, host_srcs + hdr)):
        cmd = ["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-o", CLI_PATH] + [
            s for s in host_srcs if s.endswith(".cpp")] + ["-L", PKG_DIR, "-lecdna_b200", "-Wl,-rpath,$ORIGIN/..",
                                                            "-ldl", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ (host CLI) failed:\n" + r.stdout + r.stderr)
        _stamp(CLI_PATH, host_srcs + hdr)
    return LIB_PATH
