"""GPU experiment: time the WARP fast-path kernel variants on the C2 workload (resident plan)."""
import os
import sys
import json

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from lightfm_b200 import _lightfm_fast as fast  # noqa: E402

fast.set_mode("hogwild")
variants = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,4,5,6,7,8".split(","))]
nnz = int(sys.argv[2]) if len(sys.argv) > 2 else B.NNZ
prob = B.Problem(B.N_USERS, B.N_ITEMS, nnz, B.D, seed=2, device="cuda")
itf, usf, pos = fast.CSRMatrix(prob.itf), fast.CSRMatrix(prob.usf), fast.CSRMatrix(prob.pos)
atomg = int(sys.argv[3]) if len(sys.argv) > 3 else 1
fast.set_atomic_accumulators(atomg)
for v in variants:
    fast.set_tuning(v)
    holder = prob.holder(fast)
    plan = fast.ResidentPlan("warp", itf, usf, pos, prob.row, prob.col, prob.data, prob.data, holder, 0.0, 0.0)
    for w in range(3):
        plan.epoch(seed=100 + w, num_threads=8)
    cs = [plan.epoch(seed=200 + s, num_threads=8) for s in range(5)]
    plan.close()
    ms = sum(c["train_kernel_ms"] for c in cs) / len(cs)
    ab = sum(B.algorithmic_bytes(c, B.D) for c in cs) / len(cs)
    print(json.dumps({"variant": v, "atomic_accumulators": atomg, "train_ms": round(ms, 3), "M_inter_per_s": round(cs[0]["positives"] / ms / 1e3, 1),
                      "alg_GBps": round(ab / ms / 1e6, 1), "S": round(cs[-1]["negatives_drawn"] / cs[-1]["positives"], 3),
                      "U": round(cs[-1]["updates"] / cs[-1]["positives"], 3)}), flush=True)
