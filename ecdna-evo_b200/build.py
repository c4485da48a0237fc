"""In-tree build of libecdna_b200.so (nvcc, sm_100a only) and of the C++ host CLI."""
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "libecdna_b200.so")
CLI_PATH = os.path.join(PKG_DIR, "host", "ecdna")
CSRC = os.path.join(PKG_DIR, "csrc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-diag-suppress", "128",
              # every fused multiply-add in this library is written explicitly (the CPU oracle is compiled
              # with -ffp-contract=off and must see the same roundings)
              "-fmad=false"]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static"]
OBJ_DIR = os.path.join(PKG_DIR, "_obj")


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


def build_debug(defines=("-DECDNA_DEBUG_BOUNDS",), suffix="dbg"):
    """libecdna_b200_dbg.so: the same sources with -DECDNA_DEBUG_BOUNDS (device-side asserts on every window
    address, record index and output index).  Select it with ECDNA_B200_LIB=<path> (scripts/sanitize_cases.py).
    Other `defines` / `suffix` build an experimental variant next to the product library the same way."""
    out = os.path.join(PKG_DIR, f"libecdna_b200_{suffix}.so")
    units = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ_DIR, exist_ok=True)

    def one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + f".{suffix}.o")
        r = subprocess.run([_nvcc()] + NVCC_FLAGS + list(defines) + ["-c", "-o", obj, src], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(one, units))
    r = subprocess.run([_nvcc()] + LINK_FLAGS + ["-o", out] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc (link) failed:\n" + r.stdout + r.stderr)
    return out


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libecdna_b200.so and host/*.cpp into the `ecdna` CLI."""
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(ROOT, "include", "ecdna_b200.h")]
    flags = " ".join(NVCC_FLAGS + LINK_FLAGS)
    if force or _stale(LIB_PATH, srcs, flags):
        # one object per translation unit (the kernel is instantiated per tile width in its own .cu), in parallel
        from concurrent.futures import ThreadPoolExecutor
        os.makedirs(OBJ_DIR, exist_ok=True)
        units = [f for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
        headers = [x for x in srcs if not x.endswith(".cu")]

        def compile_unit(u):
            src, obj = os.path.join(CSRC, u), os.path.join(OBJ_DIR, u[:-3] + ".o")
            if not force and not _stale(obj, [src] + headers, flags):
                return u, 0, ""
            cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode == 0:
                _stamp(obj, [src] + headers, flags)
            return u, r.returncode, r.stdout + r.stderr

        with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 1)) as ex:
            results = list(ex.map(compile_unit, units))
        bad = [(u, out) for u, rc, out in results if rc != 0]
        if bad:
            raise RuntimeError("nvcc failed:\n" + "\n".join(f"== {u}\n{out}" for u, out in bad))
        if verbose:
            for u, _, out in results:
                print("==", u)
                print(out)
        cmd = [_nvcc()] + LINK_FLAGS + ["-o", LIB_PATH] + [os.path.join(OBJ_DIR, u[:-3] + ".o") for u in units]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc (link) failed:\n" + r.stdout + r.stderr)
        _stamp(LIB_PATH, srcs, flags)
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
