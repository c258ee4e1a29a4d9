"""Build eig_b200/libeigb200.so with nvcc for sm_100a (cross-compiles without a GPU)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libeigb200.so")
SRCS = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))
DEPS = SRCS + glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + [os.path.join(HERE, "..", "include", "eigb200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=default", "--expt-relaxed-constexpr"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in DEPS)


def _compile(verbose):
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for s in SRCS:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        cmd = [NVCC] + FLAGS + os.environ.get("EB_NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    ok = True
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        ok &= p.returncode == 0
    if not ok:
        raise RuntimeError("nvcc failed")
    # link to a temporary name, then rename: a process that dlopens the library never sees a half-written file
    tmp = LIB + ".tmp.%d" % os.getpid()
    subprocess.run([NVCC, "-shared", "-o", tmp] + objs + ["-lcudart"], check=True)
    os.replace(tmp, LIB)


def build(force=False, verbose=False):
    """Build under an exclusive file lock: the ranks of a torchrun launch may all find the library stale at the same time (a source
    edited after the last build); one of them builds, the others wait and then find it fresh."""
    if not force and not needs_build():
        return LIB
    import fcntl
    with open(os.path.join(HERE, ".build.lock"), "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        try:
            if force or needs_build():
                _compile(verbose)
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
