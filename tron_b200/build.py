"""Build libtron_b200.so and the `tron` CLI, in tree, for sm_100a.

    python -m tron_b200.build [--force]

Outputs: tron_b200/lib/libtron_b200.so, tron_b200/bin/tron (git-ignored; they
travel to the GPU box with the gpurun snapshot).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
BINDIR = os.path.join(HERE, "bin")
LIB = os.path.join(LIBDIR, "libtron_b200.so")
EXE = os.path.join(BINDIR, "tron")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "--use_fast_math", "-lineinfo", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fvisibility=default", "-w"] + ARCH + os.environ.get("TRON_NVCC_EXTRA", "").split()
CU_SOURCES = ["plan.cu", "grid.cu", "grid_tile.cu", "grid_scatter.cu", "grid_wide.cu", "degrid.cu", "degrid_wide.cu", "fft.cu", "combine.cu", "cgnr.cu", "legacy.cu", "comm.cu"]
HOST_SOURCES = ["ra.c", "float16.cpp"]


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


HASHFILE = LIB + ".srchash"


def source_hash():
    """sha256 over the sources the library is built from (and the build flags)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in sorted(sources()):
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _recorded_hash():
    try:
        return open(HASHFILE).read().strip()
    except OSError:
        return None


def sources():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    out += [os.path.join(HERE, "..", "include", f) for f in ("tron.h", "ra.h", "float16.h")]
    return out


def build(force=False, verbose=False, check_stale=False):
    """Build the library and the CLI.  By default an existing library is used as is (tests, bench.py
    and every rank of a torchrun launch call this; the snapshot on the GPU box has fresh mtimes, so a
    staleness check there would rebuild -- concurrently on all ranks).  `force` / `check_stale`
    (the `python -m tron_b200.build` entry) rebuild when sources are newer."""
    if os.path.exists(LIB) and os.path.exists(EXE) and not force and not check_stale:
        # An existing library is used as is -- unless it was built from OTHER sources than the ones in the tree
        # (edited csrc/*.cu and a test run on the previous binary would report green for code that never ran).
        # The check is a content hash written next to the library, not mtimes: the snapshot on the GPU box has
        # arbitrary mtimes and every rank of a torchrun launch comes through here.  A rebuild is serialised by a
        # file lock.
        if shutil.which("nvcc") is None or _recorded_hash() == source_hash():
            return LIB
        import fcntl
        with open(os.path.join(LIBDIR, ".build.lock"), "w") as lock:
            fcntl.flock(lock, fcntl.LOCK_EX)
            if _recorded_hash() != source_hash():
                return build(force=True, verbose=verbose)
        return LIB
    if shutil.which("nvcc") is None:
        if os.path.exists(LIB):
            return LIB
        raise RuntimeError("nvcc not found and no prebuilt libtron_b200.so")
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(BINDIR, exist_ok=True)
    deps = sources()
    if not force and not _stale(LIB, deps) and not _stale(EXE, deps):
        return LIB
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for f in CU_SOURCES:
        o = os.path.join(objdir, f + ".o")
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, f), "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    o = os.path.join(objdir, "ra.o")
    _run(["gcc", "-O2", "-fPIC", "-std=gnu99", "-c", os.path.join(CSRC, "ra.c"), "-o", o])
    objs.append(o)
    o = os.path.join(objdir, "float16.o")
    _run(["g++", "-O2", "-fPIC", "-std=c++14", "-c", os.path.join(CSRC, "float16.cpp"), "-o", o])
    objs.append(o)
    log = []
    for cmd, p in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), out))
    _run(["nvcc", "-shared"] + ARCH + objs + ["-ldl", "-o", LIB])
    _run(["nvcc"] + NVCC_FLAGS + [os.path.join(CSRC, "tron_main.cu"), "-o", EXE,
                                  "-L" + LIBDIR, "-ltron_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../lib"])
    shutil.rmtree(objdir, ignore_errors=True)
    with open(HASHFILE, "w") as fh:
        fh.write(source_hash() + "\n")
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, check_stale=True))
