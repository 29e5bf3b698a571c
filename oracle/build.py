"""Build recipes for the test oracle.  TEST INFRASTRUCTURE ONLY.

* ``build_oracle()``  -> oracle/liboracle.so       (gcc, the C restatement)
* ``build_ref()``     -> oracle/_ref/{tron_ref, libtronref.so, libtronref_mc64.so}
  compiled from the reference sources WHERE THEY LIE under /root/reference/src
  (never copied), with the reference Makefile's flags (src/Makefile:3-4) plus an
  explicit sm_100 target and minus ``-ccbin gcc-6`` (absent here).  Only run in
  the build container; the GPU box uses the prebuilt files.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
REF_OUT = os.path.join(HERE, "_ref")
ORACLE_SO = os.path.join(HERE, "liboracle.so")

# reference Makefile CFLAGS (src/Makefile:3) + target arch
REF_CFLAGS = ["-O3", "-Wno-deprecated-gpu-targets", "--use_fast_math", "-D_FORCE_INLINES",
              "-DCUDA_HOST_MALLOC", "-gencode", "arch=compute_100,code=sm_100", "-w"]
REF_LFLAGS = ["-lcufft", "-lm", "-lcublas"]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def _run(cmd, cwd=None):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build_oracle(force=False):
    srcs = [os.path.join(HERE, "tron_oracle.c"), os.path.join(HERE, "tron_oracle.h")]
    if not force and os.path.exists(ORACLE_SO) and (_newer(ORACLE_SO, srcs) or shutil.which("gcc") is None):
        return ORACLE_SO
    # -march=native: rebuilt on whichever host runs it (a copied tree has fresh mtimes, so the GPU box
    # recompiles for its own CPU); written to a temporary name and renamed so a reader never sees half a file
    tmp = ORACLE_SO + ".%d.tmp" % os.getpid()
    _run(["gcc", "-O3", "-march=native", "-fopenmp", "-fno-fast-math", "-ffp-contract=off",
          "-std=gnu99", "-shared", "-fPIC", "-o", tmp, srcs[0], "-lm"])
    os.replace(tmp, ORACLE_SO)
    return ORACLE_SO


def ref_available():
    return os.path.isfile(os.path.join(REF_SRC, "tron.cu"))


def build_ref(force=False):
    """Compile the unmodified reference.  Returns the output dir or None."""
    if not ref_available() or shutil.which("nvcc") is None:
        return REF_OUT if os.path.isfile(os.path.join(REF_OUT, "libtronref.so")) else None
    os.makedirs(REF_OUT, exist_ok=True)
    harness = os.path.join(HERE, "ref_harness.cu")
    ref_files = [os.path.join(REF_SRC, f) for f in
                 ("tron.cu", "tron.h", "ra.cu", "ra.h", "float16.cu", "float16.h", "float2math.h")]
    tmp = os.path.join(REF_OUT, "obj")
    os.makedirs(tmp, exist_ok=True)

    # 1. the stock CLI binary: tron.cu + ra.cu + float16.cu, as src/Makefile:12-16
    exe = os.path.join(REF_OUT, "tron_ref")
    if force or not _newer(exe, ref_files):
        objs = []
        for f in ("tron", "ra", "float16"):
            o = os.path.join(tmp, f + ".o")
            _run(["nvcc"] + REF_CFLAGS + ["-dc", os.path.join(REF_SRC, f + ".cu"), "-o", o])
            objs.append(o)
        _run(["nvcc", "-gencode", "arch=compute_100,code=sm_100"] + objs + REF_LFLAGS + ["-o", exe])

    # 2. harness libraries (stock MAXCHAN, and the widened parity-only variant)
    for name, extra in (("libtronref.so", []), ("libtronref_mc64.so", ["-DTRONREF_MAXCHAN=64"])):
        so = os.path.join(REF_OUT, name)
        if not force and _newer(so, ref_files + [harness]):
            continue
        tag = name.replace(".so", "")
        objs = []
        for f, src in (("harness", harness), ("ra", os.path.join(REF_SRC, "ra.cu")),
                       ("float16", os.path.join(REF_SRC, "float16.cu"))):
            o = os.path.join(tmp, "%s_%s.o" % (tag, f))
            _run(["nvcc"] + REF_CFLAGS + extra + ["-I", REF_SRC, "-Xcompiler", "-fPIC", "-dc", src, "-o", o])
            objs.append(o)
        _run(["nvcc", "-shared", "-gencode", "arch=compute_100,code=sm_100"] + objs + REF_LFLAGS + ["-o", so])
    shutil.rmtree(tmp, ignore_errors=True)
    return REF_OUT


if __name__ == "__main__":
    print(build_oracle(force=True))
    print(build_ref(force=True))
