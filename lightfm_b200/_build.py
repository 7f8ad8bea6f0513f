"""In-tree build of libfm_cuda.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m lightfm_b200._build [--force] [--verbose]

Objects go to lightfm_b200/csrc/build/, the library to lightfm_b200/csrc/libfm_cuda.so
(git-ignored, but shipped to the GPU box with the working tree).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libfm_cuda.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"]

# (source, extra flags).  Replay / predict are the bit-parity paths: no FMA contraction.
SOURCES = [
    ("lfm_replay.cu", ["--fmad=false"]),
    ("lfm_predict.cu", ["--fmad=false"]),
    ("lfm_hogwild.cu", []),
    ("lfm_host.cu", []),
]
HEADERS = ["lfm_common.cuh", "lfm_hogwild_fast.cuh", "lfm_replay_fast.cuh", "lfm_replay_dataflow.cuh", os.path.join("..", "..", "include", "lfm_cuda.h")]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libfm_cuda.so")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = _nvcc()
    bdir = os.path.join(CSRC, "build")
    os.makedirs(bdir, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs, jobs = [], []
    for src, extra in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(bdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for src, proc in jobs:  # the translation units compile side by side
        out, _ = proc.communicate()
        if verbose or proc.returncode != 0:
            sys.stderr.write(out)
        if proc.returncode != 0:
            failed.append(src)
    if failed:
        raise RuntimeError("nvcc failed on %s" % ", ".join(failed))
    if force or _stale(OUT, objs):
        tmp = OUT + ".tmp.%d" % os.getpid()   # link aside, then rename: the library is never half-written
        cmd = [nvcc] + ARCH + ["-shared", "-o", tmp] + objs
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("link failed")
        os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
