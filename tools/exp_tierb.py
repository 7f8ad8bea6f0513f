"""GPU experiment: what moves held-out precision@10 of the hogwild WARP kernel at C2 shape
(tier-B problem c2_warp / c2_kos): atomic-return accumulators on/off x interactions in flight.
Usage: python tools/exp_tierb.py [case ...]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402

cu = H.cuda_native()
bands = json.load(open(os.path.join(H.GOLDEN_DIR, "tierb_bands.json")))
for name in (sys.argv[1:] or ["c2_warp", "c2_kos"]):
    fit, train, test, users = H.tierb_problem(name)
    print(json.dumps({"case": name, "reference_band": bands[name]["band"]}), flush=True)
    for atomg in (1, 0):
        for divisor in (128, 1024, 8192):
            cu.module.set_atomic_accumulators(atomg)
            cu.module.set_inflight_divisor(divisor)
            runs = [H.eval_subset(H.tierb_fit(cu, name, seed, 8), train, test, users) for seed in (0, 1, 2)]
            print(json.dumps({"case": name, "atomic_accumulators": atomg, "inflight_cap": max(64, fit.nnz // divisor),
                              "p_at_10": round(float(np.mean([r[0] for r in runs])), 4),
                              "auc": round(float(np.mean([r[1] for r in runs])), 4)}), flush=True)
cu.module.set_atomic_accumulators(1)
cu.module.set_inflight_divisor(128)
