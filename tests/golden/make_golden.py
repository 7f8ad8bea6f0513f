"""Generate golden input/output vectors from the REAL reference (oracle/_ref/strict).

Run where /root/reference was available to build oracle/_ref:
    python oracle/build_ref.py && python tests/golden/make_golden.py
Writes tests/golden/golden_<case>.npz: the full inputs (COO interactions, feature
CSRs, sample weights, initial state, per-epoch shuffles and rand_r seeds) and the
reference's outputs (12 state arrays after training, predictions, ranks, AUC), so
the tests depend neither on /root/reference nor on re-deriving the inputs.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers as H  # noqa: E402

CASES = {
    "warp_adagrad_identity": dict(loss="warp", schedule="adagrad", d=8, feats=False, alpha=0.0,
                                  epochs=2, users=60, items=40, nnz=600),
    "bpr_adadelta_features_l2": dict(loss="bpr", schedule="adadelta", d=8, feats=True, alpha=1e-3,
                                     epochs=2, users=50, items=45, nnz=500),
    "logistic_adagrad_weights_l2": dict(loss="logistic", schedule="adagrad", d=16, feats=True,
                                        alpha=1e-4, epochs=2, users=40, items=50, nnz=500,
                                        weights=True),
    "kos_adagrad_identity_d12": dict(loss="warp-kos", schedule="adagrad", d=12, feats=False,
                                     alpha=0.0, epochs=2, users=45, items=60, nnz=700, k=3, n=5),
    "warp_adagrad_features": dict(loss="warp", schedule="adagrad", d=32, feats=True, alpha=0.0,
                                  epochs=1, users=40, items=40, nnz=400, max_sampled=4),
}


class RecordingState(object):
    """RandomState proxy that records what run_epoch draws, so tests can replay it."""

    def __init__(self, rs):
        self.rs, self.shuffles, self.seeds = rs, [], []

    def shuffle(self, arr):
        self.rs.shuffle(arr)
        self.shuffles.append(arr.copy())

    def randint(self, *a, **k):
        out = self.rs.randint(*a, **k)
        self.seeds.append(np.asarray(out).copy())
        return out

    def rand(self, *a):
        return self.rs.rand(*a)


def csr_parts(prefix, m, out):
    m = m.tocsr()
    out[prefix + "_indptr"] = m.indptr.astype(np.int32)
    out[prefix + "_indices"] = m.indices.astype(np.int32)
    out[prefix + "_data"] = m.data.astype(np.float32)
    out[prefix + "_shape"] = np.array(m.shape, dtype=np.int64)


def main():
    ref = H.reference_native("strict")
    for ci, (name, c) in enumerate(sorted(CASES.items())):
        inter = H.synthetic_interactions(c["users"], c["items"], c["nnz"], seed=100 + ci,
                                         signed=(c["loss"] == "logistic"))
        itf = H.tag_features(c["items"], 12, 3, 200 + ci) if c["feats"] else \
            sp.identity(c["items"], dtype=np.float32, format="csr")
        usf = H.tag_features(c["users"], 9, 2, 300 + ci) if c["feats"] else \
            sp.identity(c["users"], dtype=np.float32, format="csr")
        sw = None
        if c.get("weights"):
            sw = (0.5 + np.random.default_rng(400 + ci).random(inter.nnz)).astype(np.float32)
        hp = H.Hyper(d=c["d"], schedule=c["schedule"], item_alpha=c["alpha"], user_alpha=c["alpha"],
                     k=c.get("k", 5), n=c.get("n", 10), max_sampled=c.get("max_sampled", 10))
        rs = RecordingState(np.random.RandomState(7 + ci))
        arrays = H.init_arrays(rs, itf.shape[1], usf.shape[1], c["d"], c["schedule"])
        out = {"init_" + k: v.copy() for k, v in arrays.items()}
        for _ in range(c["epochs"]):
            H.run_epoch(ref, c["loss"], inter, arrays, hp, rs, itf, usf, sw)
        for k, v in arrays.items():
            out["final_" + k] = v
        out["shuffles"] = np.stack(rs.shuffles)
        out["seeds"] = np.stack(rs.seeds).astype(np.uint32) if rs.seeds else np.zeros((0, 1), np.uint32)
        out["row"], out["col"], out["data"] = inter.row, inter.col, inter.data
        out["shape"] = np.array(inter.shape, dtype=np.int64)
        if sw is not None:
            out["sample_weight"] = sw
        csr_parts("itf", itf, out)
        csr_parts("usf", usf, out)
        out["hyper"] = np.array([c["d"], int(c["schedule"] == "adadelta"), hp.lr, hp.rho, hp.eps,
                                 hp.max_sampled, c["alpha"], hp.k, hp.n, c["epochs"]], dtype=np.float64)
        out["loss"] = np.array(c["loss"])

        # scoring outputs on the trained state
        rng = np.random.default_rng(500 + ci)
        pu = rng.integers(0, c["users"], 64).astype(np.int32)
        pi = rng.integers(0, c["items"], 64).astype(np.int32)
        pred = np.empty(64, dtype=np.float32)
        h = H.holder(ref, arrays, hp)
        ref.predict_lightfm(ref.CSRMatrix(itf), ref.CSRMatrix(usf), pu, pi, pred, h, 1)
        out["pred_users"], out["pred_items"], out["pred"] = pu, pi, pred
        # test = a fresh sample of pairs disjoint from train
        dense = inter.tocsr().astype(bool)
        cand = H.synthetic_interactions(c["users"], c["items"], c["nnz"] // 3, seed=600 + ci).tocsr().astype(bool)
        test = (cand > dense).astype(np.float32).tocsr()
        test.sort_indices()
        train = inter.tocsr().astype(np.float32)
        train.sort_indices()
        ranks = np.zeros_like(test.data)
        ref.predict_ranks(ref.CSRMatrix(itf), ref.CSRMatrix(usf), ref.CSRMatrix(test),
                          ref.CSRMatrix(train), ranks, h, 1)
        csr_parts("test", test, out)
        csr_parts("train", train, out)
        out["ranks"] = ranks.copy()
        ntp = np.asarray(train.getnnz(axis=1)).astype(np.int32)
        auc = np.zeros(test.shape[0], dtype=np.float32)
        rank_sorted = ranks.copy()
        rk = sp.csr_matrix((rank_sorted, test.indices, test.indptr), shape=test.shape)
        ref.calculate_auc_from_rank(ref.CSRMatrix(rk), ntp, rk.data, auc, 1)
        out["auc"], out["ranks_sorted"], out["num_train_positives"] = auc, rk.data.copy(), ntp
        path = os.path.join(HERE, "golden_%s.npz" % name)
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
