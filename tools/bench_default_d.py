"""GPU: WARP at the class default no_components=10 (generic scalar-lane kernel) on the C2 shape."""
import json, os, sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
from lightfm_b200 import _lightfm_fast as fast
fast.set_mode("hogwild")
for d in (10, 48):
    prob = B.Problem(B.N_USERS, B.N_ITEMS, B.NNZ, d, seed=2, device="cuda", pin=False)
    plan = fast.ResidentPlan("warp", fast.CSRMatrix(prob.itf), fast.CSRMatrix(prob.usf), fast.CSRMatrix(prob.pos),
                             prob.row, prob.col, prob.data, prob.data, prob.holder(fast), 0.0, 0.0)
    for w in range(2):
        plan.epoch(seed=w, num_threads=8)
    cs = [plan.epoch(seed=10 + s, num_threads=8) for s in range(3)]
    plan.close()
    ms = sum(c["train_kernel_ms"] for c in cs) / 3
    print(json.dumps({"d": d, "train_ms": round(ms, 2), "M_inter_per_s": round(cs[0]["positives"] / ms / 1e3, 1)}), flush=True)
