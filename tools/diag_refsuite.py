"""GPU diagnostic: run the reference's test_api.py over the shim in several ways (order dependence?)."""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
tmp = tempfile.mkdtemp(prefix="refsuite_")
pkg = os.path.join(tmp, "lightfm")
os.makedirs(pkg)
for f in ("__init__.py", "lightfm.py", "evaluation.py", "cross_validation.py", "data.py", "version.py"):
    shutil.copy(os.path.join(REF, "csrc", f), os.path.join(pkg, f))
open(os.path.join(pkg, "_lightfm_fast.py"), "w").write(
    "from lightfm_b200._lightfm_fast import *  # noqa\n"
    "from lightfm_b200 import _lightfm_fast as _m\n"
    "globals()['__test_in_positives'] = getattr(_m, '__test_in_positives')\n")
tdir = os.path.join(tmp, "reftests")
os.makedirs(tdir)
shutil.copy(os.path.join(REF, "tests", "test_api.py"), tdir)
# a probe that repeats the failing scenario inside the same kind of process, with diagnostics
open(os.path.join(tdir, "test_probe.py"), "w").write('''
import numpy as np, scipy.sparse as sp
from lightfm.lightfm import LightFM
def test_probe():
    import sys, os
    sys.path.insert(0, os.path.join(os.environ["LFM_ROOT"], "tests"))
    import helpers as H
    orc = H.oracle_native()
    for rep in range(300):
        train = sp.rand(10, 100, format="csr", random_state=42)
        model = LightFM()
        model.fit_partial(train)
        rank_input = sp.csr_matrix(np.ones((10, 100)))
        ranks = model.predict_rank(rank_input, num_threads=2).todense()
        bad = [r for r in range(10) if not np.all(np.sort(ranks[r]) == np.arange(100))]
        arr = {k: getattr(model, k) for k in H.MODEL_ARRAYS}
        want = np.zeros(1000, np.float32)
        t32 = sp.csr_matrix(rank_input, dtype=np.float32)
        orc.predict_ranks(orc.CSRMatrix(sp.identity(100, dtype=np.float32, format="csr")),
                          orc.CSRMatrix(sp.identity(10, dtype=np.float32, format="csr")), orc.CSRMatrix(t32),
                          orc.CSRMatrix(sp.csr_matrix((10, 100), dtype=np.float32)), want,
                          H.holder(orc, arr, H.Hyper(d=10)), 1)
        mism = int((np.asarray(ranks).ravel() != want).sum())
        if bad or mism or rep % 50 == 0:
            print("rep", rep, "bad rows", bad, "mismatches vs oracle", mism)
        for r in bad[:2]:
            s = model.predict(r, np.arange(100, dtype=np.int32))
            row = np.asarray(ranks[r]).ravel()
            order = np.argsort(-s, kind="stable")
            want = np.empty(100); want[order] = np.arange(100)
            diff = np.flatnonzero(row != want)
            print("   row", r, "n_diff", len(diff), [(int(i), float(row[i]), float(want[i]), float(s[i])) for i in diff[:6]],
                  "dtype", model.item_embeddings.dtype, model.user_embeddings.flags.c_contiguous,
                  "finite", np.isfinite(s).all(), "unique scores", len(np.unique(s)))
        assert not bad
''')
env = dict(os.environ)
env["PYTHONPATH"] = os.pathsep.join([tmp, ROOT, env.get("PYTHONPATH", "")])
env["LFM_ROOT"] = ROOT
for label, args in (("probe alone (300 random models, ranks vs oracle)", ["reftests/test_probe.py", "-s"]),
                    ("whole test_api.py then probe", ["reftests/test_api.py", "reftests/test_probe.py", "-s"])):
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider"] + args, cwd=tmp, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    print("=====", label, "rc", r.returncode)
    print("\n".join(ln for ln in r.stdout.splitlines() if "rep " in ln or "   row" in ln or "FAILED" in ln or "passed" in ln or "failed" in ln or "Error" in ln))
