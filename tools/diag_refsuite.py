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
    for rep in range(5):
        train = sp.rand(10, 100, format="csr", random_state=42)
        model = LightFM()
        model.fit_partial(train)
        rank_input = sp.csr_matrix(np.ones((10, 100)))
        ranks = model.predict_rank(rank_input, num_threads=2).todense()
        bad = [r for r in range(10) if not np.all(np.sort(ranks[r]) == np.arange(100))]
        print("rep", rep, "bad rows", bad)
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
for label, args in (("probe alone", ["reftests/test_probe.py", "-s"]),
                    ("test_predict_ranks alone", ["reftests/test_api.py", "-k", "test_predict_ranks"]),
                    ("whole test_api.py", ["reftests/test_api.py"]),
                    ("whole test_api.py then probe", ["reftests/test_api.py", "reftests/test_probe.py", "-s"])):
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider"] + args, cwd=tmp, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    print("=====", label, "rc", r.returncode)
    print("\n".join(ln for ln in r.stdout.splitlines() if ln.startswith(("rep", "   row", "FAILED")) or "passed" in ln or "failed" in ln))
