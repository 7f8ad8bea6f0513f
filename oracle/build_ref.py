"""Build the UNMODIFIED reference (lyst/lightfm) hot path into ``oracle/_ref``.

TEST INFRASTRUCTURE ONLY.  Nothing under ``lightfm_b200/`` may import this.

The reference's native path is one Cython template
(``/root/reference/lightfm/_lightfm_fast.pyx.template``).  Its shipped,
pre-generated C (Cython 0.29) does not compile against CPython 3.12, so the
recipe re-runs the reference's own ``setup.py cythonize`` step on a scratch
COPY of the tree (``/root/reference`` is read-only) and compiles the result
with gcc.  No reference source is modified except one build directive in the
scratch copy's setup.py (``legacy_implicit_noexcept`` -- Cython 3 otherwise
rejects the qsort/bsearch callback types; it does not change arithmetic).

Outputs (all git-ignored, all travel to the GPU box with the snapshot):

  oracle/_ref/fast/lightfm/     shipped flags  (-ffast-math -march=native -fopenmp)
  oracle/_ref/strict/lightfm/   LIGHTFM_NO_CFLAGS=1 (IEEE, no FMA contraction, portable)
  oracle/_ref/csrc/             the generated C, so ``rebuild_native()`` can
                                recompile with -march=native for the *GPU box's*
                                host CPU when /root/reference is absent.

Usage:  python oracle/build_ref.py            (needs /root/reference)
        python oracle/build_ref.py --native   (box-side recompile from csrc/)
"""
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REFERENCE = os.environ.get("LFM_REFERENCE_DIR", "/root/reference")
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"

# the reference's offline test modules (the other three need the network), run UNMODIFIED
# against the drop-in package by tests/test_gpu_reference_suite.py
REF_TESTS = ["__init__.py", "test_api.py", "test_evaluation.py", "test_fast_functions.py", "test_data.py"]

PY_FILES = ["__init__.py", "_lightfm_fast.py", "lightfm.py", "evaluation.py",
            "cross_validation.py", "data.py", "version.py"]


def _run(cmd, cwd, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(cmd, cwd=cwd, env=e, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-4000:])
        raise RuntimeError("command failed: %s" % " ".join(cmd))
    return r.stdout


def _ext_suffix():
    return sysconfig.get_config_var("EXT_SUFFIX")


def _compile(csrc, dst_pkg, flags, name="_lightfm_fast_openmp"):
    """gcc one generated C file into a CPython extension inside dst_pkg."""
    inc = sysconfig.get_paths()["include"]
    out = os.path.join(dst_pkg, name + _ext_suffix())
    cmd = [GCC, "-shared", "-fPIC", "-O3", "-fopenmp", "-fwrapv",
           "-fno-strict-aliasing", "-DNDEBUG", "-I", inc] + flags + \
          [csrc, "-o", out, "-lm"]
    _run(cmd, cwd=dst_pkg)
    return out


def _install_py(src_pkg, dst_pkg):
    os.makedirs(dst_pkg, exist_ok=True)
    for f in PY_FILES:
        shutil.copy(os.path.join(src_pkg, f), os.path.join(dst_pkg, f))
    # datasets/ needs `requests` + network; the package __init__ does not import it.


def build_from_reference():
    if not os.path.isdir(REFERENCE):
        raise RuntimeError("%s not present; use --native on the GPU box" % REFERENCE)
    tmp = tempfile.mkdtemp(prefix="lfm_ref_build_")
    tree = os.path.join(tmp, "ref")
    shutil.copytree(REFERENCE, tree, ignore=shutil.ignore_patterns(".git", "doc", "examples"))
    os.chmod(tree, 0o755)
    for root, dirs, files in os.walk(tree):
        for d in dirs:
            os.chmod(os.path.join(root, d), 0o755)
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
    sp = os.path.join(tree, "setup.py")
    s = open(sp).read()
    needle = "compiler_directives={'language_level' : \"3\"}"
    assert needle in s, "reference setup.py changed; update build_ref.py"
    s = s.replace(needle, "compiler_directives={'language_level' : \"3\", "
                          "'legacy_implicit_noexcept': True}")
    open(sp, "w").write(s)
    env = {"CC": GCC, "LDSHARED": GCC + " -shared"}
    _run([sys.executable, "setup.py", "cythonize"], cwd=tree, env=env)

    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(os.path.join(OUT, "csrc"))
    gen_c = os.path.join(tree, "lightfm", "_lightfm_fast_openmp.c")
    shutil.copy(gen_c, os.path.join(OUT, "csrc", "_lightfm_fast_openmp.c"))
    for f in PY_FILES:
        shutil.copy(os.path.join(tree, "lightfm", f), os.path.join(OUT, "csrc", f))

    for variant, flags in (("fast", ["-ffast-math", "-march=native"]),
                           ("strict", [])):
        pkg = os.path.join(OUT, variant, "lightfm")
        _install_py(os.path.join(tree, "lightfm"), pkg)
        _compile(os.path.join(OUT, "csrc", "_lightfm_fast_openmp.c"), pkg, flags)
    shutil.rmtree(tmp, ignore_errors=True)
    install_tests()
    print("reference built into", OUT)


def install_tests():
    """Copy the reference's offline test modules into oracle/_ref/tests (git-ignored output)."""
    src = os.path.join(REFERENCE, "tests")
    if not os.path.isdir(src):
        raise RuntimeError("%s not present" % src)
    dst = os.path.join(OUT, "tests")
    os.makedirs(dst, exist_ok=True)
    for f in REF_TESTS:
        shutil.copy(os.path.join(src, f), os.path.join(dst, f))
        os.chmod(os.path.join(dst, f), 0o644)
    return dst


def rebuild_native():
    """Recompile the generated C with the shipped flags for THIS host's CPU.

    Used by ``bench.py --impl reference`` on the GPU box so that
    ``-march=native`` matches the box's host cores (what a user gets from
    ``pip install lightfm`` there).  Returns the variant directory to import
    from, falling back to the prebuilt portable 'strict' build.
    """
    csrc = os.path.join(OUT, "csrc", "_lightfm_fast_openmp.c")
    if not os.path.exists(csrc):
        raise RuntimeError("oracle/_ref/csrc missing: run build_ref.py where /root/reference exists")
    pkg = os.path.join(OUT, "native", "lightfm")
    try:
        _install_py(os.path.join(OUT, "csrc"), pkg)
        _compile(csrc, pkg, ["-ffast-math", "-march=native"])
        return os.path.join(OUT, "native")
    except Exception as exc:  # pragma: no cover - depends on box toolchain
        sys.stderr.write("native rebuild failed (%s); using prebuilt strict variant\n" % exc)
        return os.path.join(OUT, "strict")


if __name__ == "__main__":
    if "--native" in sys.argv:
        print(rebuild_native())
    elif "--tests" in sys.argv:
        print(install_tests())
    else:
        build_from_reference()
