"""Generate tests/golden/tierb_bands.json: the reference's own run-to-run band of held-out
precision@10 / AUC on the tier-B problems (SURVEY 8(c) tier B), from

  * the oracle restatement at num_threads=1 (bit-equal to the real reference, deterministic), and
  * the REAL reference (oracle/_ref/fast: shipped flags, OpenMP) at 8 threads (Hogwild, racy),

five seeds each.  The GPU tests (tests/test_gpu_tierb.py) train the same problems on the B200 and
require the GPU mean to sit inside this band (widened by a stated tolerance).  Run where
/root/reference was available to build oracle/_ref:

    python oracle/build_ref.py && python tests/golden/make_tierb.py [case ...]
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers as H  # noqa: E402

OUT = os.path.join(HERE, "tierb_bands.json")


def main():
    cases = sys.argv[1:] or list(H.TIERB)
    bands = json.load(open(OUT)) if os.path.exists(OUT) else {}
    orc = H.oracle_native()
    ref = H.reference_native("fast")
    for name in cases:
        fit, train, test, users = H.tierb_problem(name)
        rec = {"config": {k: v for k, v in H.TIERB[name].items()}, "digest": H.data_digest(fit),
               "eval_users": int(len(users)), "seeds": list(H.TIERB_SEEDS), "runs": {}}
        for side, api, nt in (("oracle_1thread", orc, 1), ("reference_8threads", ref, 8)):
            out = []
            for seed in H.TIERB_SEEDS:
                t0 = time.time()
                arr = H.tierb_fit(api, name, seed, nt)
                p, a = H.eval_subset(arr, train, test, users)
                out.append({"seed": seed, "p_at_10": p, "auc": a, "fit_s": round(time.time() - t0, 2)})
                print(name, side, out[-1], flush=True)
            rec["runs"][side] = out
        allp = [r["p_at_10"] for v in rec["runs"].values() for r in v]
        alla = [r["auc"] for v in rec["runs"].values() for r in v]
        rec["band"] = {"p_at_10": [min(allp), max(allp)], "auc": [min(alla), max(alla)]}
        bands[name] = rec
        json.dump(bands, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
