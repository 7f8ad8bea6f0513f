"""Minimal end-to-end use of lightfm_b200 on a B200 (mirrors the reference's quickstart,
doc/quickstart.rst, on synthetic data since there is no network for MovieLens).

    python examples/quickstart.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout

from lightfm_b200 import LightFM
from lightfm_b200.cross_validation import random_train_test_split
from lightfm_b200.evaluation import auc_score, precision_at_k
from lightfm_b200.synthetic import interactions

data = interactions(n_users=20_000, n_items=5_000, nnz=2_000_000, seed=0)
train, test = random_train_test_split(data, test_percentage=0.2, random_state=np.random.RandomState(7))

model = LightFM(loss="warp", no_components=64, learning_rate=0.05, random_state=0)
t0 = time.time()
model.fit(train, epochs=10, num_threads=8)          # > 1: GPU throughput kernels
print("10 epochs over %d interactions: %.2f s" % (train.nnz, time.time() - t0))

print("train precision@10 %.3f   test precision@10 %.3f" % (
    precision_at_k(model, train, k=10).mean(),
    precision_at_k(model, test, train_interactions=train, k=10).mean()))
print("test AUC %.3f" % auc_score(model, test, train_interactions=train).mean())

# scores for one user over the whole catalogue, best first
scores = model.predict(3, np.arange(5_000, dtype=np.int32))
print("top items for user 3:", np.argsort(-scores)[:5])
