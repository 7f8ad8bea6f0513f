"""GPU: the multi-GPU plumbing of SURVEY 8(e) on real devices -- sub-range epochs, the fused
delta-exchange sweeps (lfm_plan_delta_*), and a 2-rank sharded fit against the single-GPU fit
(skipped when fewer than two GPUs are visible)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import scipy.sparse as sp

import helpers as H

pytestmark = pytest.mark.gpu


def _plan(loss="warp", d=32, n_users=2000, n_items=800, nnz=60_000, seed=3):
    from lightfm_b200 import _lightfm_fast as fast
    inter = H.planted_clusters(n_users, n_items, nnz, seed=seed, n_clusters=8)
    arr = H.init_arrays(np.random.RandomState(seed), n_items, n_users, d)
    hp = H.Hyper(d=d)
    holder = H.holder(fast, arr, hp)
    pos = inter.tocsr()
    pos.sort_indices()
    plan = fast.ResidentPlan(loss, fast.CSRMatrix(sp.identity(n_items, dtype=np.float32, format="csr")),
                             fast.CSRMatrix(sp.identity(n_users, dtype=np.float32, format="csr")),
                             fast.CSRMatrix(pos), inter.row, inter.col, inter.data, inter.data, holder, 0.0, 0.0)
    return fast, plan, arr, inter


def test_epoch_range_visits_exactly_the_range():
    fast, plan, arr, inter = _plan()
    fast.set_mode("hogwild")
    try:
        n = inter.nnz
        a = plan.epoch_range(seed=5, begin=0, count=n // 3)
        b = plan.epoch_range(seed=6, begin=n // 3, count=-1)
        assert a["positives"] == n // 3 and b["positives"] == n - n // 3
        # users that only occur in the second part are untouched by the first call
        plan.download()
        with pytest.raises(ValueError):
            plan.epoch_range(seed=1, begin=n - 5, count=10)
    finally:
        fast.set_mode("auto")
        plan.close()


def test_epoch_range_leaves_other_users_untouched():
    fast, plan, arr, inter = _plan()
    fast.set_mode("hogwild")
    try:
        # interactions sorted by user: the first half of the list covers a prefix of the users
        order = np.argsort(inter.row, kind="stable")
        plan.close()
        arr2 = H.init_arrays(np.random.RandomState(3), inter.shape[1], inter.shape[0], 32)
        holder = H.holder(fast, arr2, H.Hyper(d=32))
        row, col = np.ascontiguousarray(inter.row[order]), np.ascontiguousarray(inter.col[order])
        pos = inter.tocsr()
        pos.sort_indices()
        plan = fast.ResidentPlan("warp", fast.CSRMatrix(sp.identity(inter.shape[1], dtype=np.float32, format="csr")),
                                 fast.CSRMatrix(sp.identity(inter.shape[0], dtype=np.float32, format="csr")),
                                 fast.CSRMatrix(pos), row, col, inter.data, inter.data, holder, 0.0, 0.0)
        half = inter.nnz // 2
        before = arr2["user_embeddings"].copy()
        plan.epoch_range(seed=9, begin=0, count=half)
        plan.download()
        last_user = row[half - 1]
        assert np.array_equal(arr2["user_embeddings"][last_user + 1:], before[last_user + 1:])
        assert not np.array_equal(arr2["user_embeddings"][:last_user], before[:last_user])
    finally:
        fast.set_mode("auto")
        plan.close()


def test_delta_sweeps_single_rank_roundtrip():
    """begin -> epoch -> make gives D = W - W0 in one packed buffer; with nothing reduced into it
    (one rank: the sum of deltas IS the local delta), apply leaves the table as the epoch left it."""
    import torch
    from lightfm_b200.sharding import CudaArrayView
    fast, plan, arr, inter = _plan()
    fast.set_mode("hogwild")
    try:
        w0 = {k: arr[k].copy() for k in H.MODEL_ARRAYS}
        plan.delta_begin(1)
        plan.epoch(seed=4, num_threads=8)
        ptr, count, _ = plan.delta_make(1)
        n_users, d = arr["user_embeddings"].shape
        assert count == 2 * n_users * d + 2 * n_users
        delta = torch.as_tensor(CudaArrayView(ptr, count), device="cuda").cpu().numpy().copy()
        plan.delta_apply(1)
        plan.download()
        got = np.concatenate([(arr[k] - w0[k]).ravel() for k in
                              ("user_embeddings", "user_embedding_gradients", "user_biases", "user_bias_gradients")])
        assert np.allclose(delta, got, rtol=0, atol=1e-6)
        assert np.abs(delta).max() > 0
        with pytest.raises(RuntimeError):
            plan.delta_apply(1)      # state machine: make must precede apply
    finally:
        fast.set_mode("auto")
        plan.close()


WORKER = textwrap.dedent('''
    import json, os, sys
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, os.environ["LFM_ROOT"]); sys.path.insert(0, os.path.join(os.environ["LFM_ROOT"], "tests"))
    import helpers as H
    from lightfm_b200 import LightFM
    from lightfm_b200.sharding import ShardedTrainer
    rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    full = H.planted_clusters(6000, 2000, 400_000, seed=4, n_clusters=16)
    train, test = H.split(full, 5)
    users = np.arange(0, 6000, 6)
    out = {}
    for axis in ("item", "user"):
        model = LightFM(loss="warp", no_components=32, random_state=0)
        tr = ShardedTrainer(model, train, axis=axis)
        tr.fit_epochs(5)
        tr.gather()
        tr.close()
        arr = {k: getattr(model, k) for k in H.MODEL_ARRAYS}
        out[axis] = H.eval_subset(arr, train, test, users)
    # item features [I | tags] under item sharding: the tag rows are shared by every shard, so the
    # whole item-feature table joins the exchanged block (SURVEY 8(e), third bullet)
    itf = H.tag_features(2000, 100, 4, seed=2)
    model = LightFM(loss="warp", no_components=32, random_state=0)
    tr = ShardedTrainer(model, train, axis="item", item_features=itf)
    tr.fit_epochs(5)
    tr.gather()
    tr.close()
    rep = {"item_embeddings": np.asarray(itf @ model.item_embeddings),
           "item_biases": np.asarray(itf @ model.item_biases).ravel(),
           "user_embeddings": model.user_embeddings, "user_biases": model.user_biases}
    out["item_features"] = H.eval_subset(rep, train, test, users)
    if rank == 0:
        print("RESULT " + json.dumps(out))
    dist.barrier(); dist.destroy_process_group()
''')


def test_two_rank_sharded_fit_matches_single_gpu_quality(tmp_path):
    from lightfm_b200 import _lightfm_fast as fast
    if fast.device_count() < 2:
        pytest.skip("needs two GPUs")
    import json
    from lightfm_b200 import LightFM
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, LFM_ROOT=H.ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
    assert r.returncode == 0 and line, r.stdout[-3000:]
    sharded = json.loads(line[0][len("RESULT "):])
    full = H.planted_clusters(6000, 2000, 400_000, seed=4, n_clusters=16)
    train, test = H.split(full, 5)
    users = np.arange(0, 6000, 6)
    model = LightFM(loss="warp", no_components=32, random_state=0)
    model.fit(train, epochs=5, num_threads=8)
    single = H.eval_subset({k: getattr(model, k) for k in H.MODEL_ARRAYS}, train, test, users)
    itf = H.tag_features(2000, 100, 4, seed=2)
    model = LightFM(loss="warp", no_components=32, random_state=0)
    model.fit(train, item_features=itf, epochs=5, num_threads=8)
    rep = {"item_embeddings": np.asarray(itf @ model.item_embeddings),
           "item_biases": np.asarray(itf @ model.item_biases).ravel(),
           "user_embeddings": model.user_embeddings, "user_biases": model.user_biases}
    single_f = H.eval_subset(rep, train, test, users)
    print("single GPU p@10 / auc:", single, "with item features:", single_f, " 2-rank sharded:", sharded)
    for axis in ("item", "user"):
        assert abs(sharded[axis][1] - single[1]) < 0.02, (axis, sharded, single)
    assert abs(sharded["item_features"][1] - single_f[1]) < 0.02, (sharded, single_f)
