"""
In-tree build of libh264scroll_b200.so: nvcc (sm_100a only) for csrc/*.cu, gcc for host/*.c.
The .so stays inside this directory so that it travels with the gpurun snapshot.
"""
import os
import subprocess
import shutil

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libh264scroll_b200.so")
OBJ = os.path.join(PKG, "build")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# experiment knobs (e.g. H264SCROLL_NVCC_EXTRA="-DCOMPOSE_MPT=4 -DCOMPOSE_MIN_CTAS=4"); empty for the shipped build
NVCC_FLAGS += os.environ.get("H264SCROLL_NVCC_EXTRA", "").split()
CFLAGS = ["-std=c99", "-O2", "-g", "-Wall", "-Wextra", "-fPIC"]


def _sources():
    cu = sorted(f for f in os.listdir(os.path.join(PKG, "csrc")) if f.endswith(".cu"))
    c = sorted(f for f in os.listdir(os.path.join(PKG, "host")) if f.endswith(".c"))
    return cu, c


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


STAMP = os.path.join(OBJ, "flags.stamp")


def _flags_changed():
    """The objects were built with other flags (an experiment variant built through H264SCROLL_NVCC_EXTRA): everything is
    stale, whatever the mtimes say - tests and the bench must never run a variant by accident."""
    want = " ".join(NVCC_FLAGS + CFLAGS)
    try:
        return open(STAMP).read() != want
    except OSError:
        return os.path.isdir(OBJ) and any(f.endswith(".o") for f in os.listdir(OBJ))


def build(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cu, c = _sources()
    hdrs = [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    hdrs += [os.path.join(PKG, "csrc", f) for f in os.listdir(os.path.join(PKG, "csrc")) if f.endswith((".cuh", ".h"))]
    hdrs += [os.path.join(PKG, "host", f) for f in os.listdir(os.path.join(PKG, "host")) if f.endswith(".h")]
    os.makedirs(OBJ, exist_ok=True)
    force = force or _flags_changed()
    inc = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(PKG, "csrc"), "-I" + os.path.join(PKG, "host")]
    objs, log = [], []
    for f in cu:
        src, o = os.path.join(PKG, "csrc", f), os.path.join(OBJ, f + ".o")
        if force or _stale(o, [src] + hdrs):
            r = subprocess.run([nvcc] + NVCC_FLAGS + inc + ["-c", src, "-o", o], capture_output=True, text=True)
            log.append(r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s" % (f, r.stderr))
        objs.append(o)
    for f in c:
        src, o = os.path.join(PKG, "host", f), os.path.join(OBJ, f + ".o")
        if force or _stale(o, [src] + hdrs):
            r = subprocess.run(["gcc"] + CFLAGS + inc + ["-c", src, "-o", o], capture_output=True, text=True)
            log.append(r.stderr)
            if r.returncode != 0:
                raise RuntimeError("gcc failed for %s:\n%s" % (f, r.stderr))
        objs.append(o)
    open(STAMP, "w").write(" ".join(NVCC_FLAGS + CFLAGS))
    if force or _stale(LIB, objs):
        r = subprocess.run([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                 "-cudart", "static"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr)
    if verbose:
        print("".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
