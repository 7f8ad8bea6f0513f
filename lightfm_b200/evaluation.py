# coding=utf-8
"""Ranking metrics on top of ``LightFM.predict_rank``.

Same four public functions, signatures and ``preserve_rows`` semantics as the
reference's ``lightfm/evaluation.py`` (``E:``): ``precision_at_k`` E:14-87,
``recall_at_k`` E:90-166, ``auc_score`` E:169-254, ``reciprocal_rank`` E:257-327.
With this package's ``LightFM`` the per-user reductions run on the device right behind the
``predict_ranks`` kernel (``LightFM.evaluate_ranks`` -> ``lfm_evaluate_ranks``), so the
``nnz_test`` ranks never cross PCIe; the values equal what the reference's numpy post-processing
of the rank matrix gives (same dtypes).  Any other model object goes through ``predict_rank`` and
the reference's own host-side reductions.
"""
import numpy as np

from ._lightfm_fast import CSRMatrix, calculate_auc_from_rank

__all__ = ["precision_at_k", "recall_at_k", "auc_score", "reciprocal_rank"]


def _ranks(model, test_interactions, train_interactions, user_features, item_features,
           num_threads, check_intersections):
    if num_threads < 1:
        raise ValueError("Number of threads must be 1 or larger.")
    return model.predict_rank(test_interactions, train_interactions=train_interactions,
                              user_features=user_features, item_features=item_features,
                              num_threads=num_threads, check_intersections=check_intersections)


def _fused(model, test_interactions, train_interactions, user_features, item_features, num_threads,
           check_intersections, k=10, **want):
    if num_threads < 1:
        raise ValueError("Number of threads must be 1 or larger.")
    return model.evaluate_ranks(test_interactions, train_interactions=train_interactions, k=k,
                                user_features=user_features, item_features=item_features,
                                num_threads=num_threads, check_intersections=check_intersections, **want)


def _row_filter(values, test_interactions, preserve_rows):
    if preserve_rows:
        return values
    return values[test_interactions.getnnz(axis=1) > 0]


def precision_at_k(model, test_interactions, train_interactions=None, k=10, user_features=None,
                   item_features=None, preserve_rows=False, num_threads=1,
                   check_intersections=True):
    """Fraction of the top-k ranked items that are test positives, per user."""
    if hasattr(model, "evaluate_ranks"):
        hits, _, _ = _fused(model, test_interactions, train_interactions, user_features, item_features,
                            num_threads, check_intersections, k=k, hits=True, best_rank=False, auc=False)
        return _row_filter(hits.astype(np.float32) / k, test_interactions, preserve_rows)
    ranks = _ranks(model, test_interactions, train_interactions, user_features, item_features,
                   num_threads, check_intersections)
    ranks.data = np.less(ranks.data, k, ranks.data)
    precision = np.squeeze(np.array(ranks.sum(axis=1))) / k
    return _row_filter(precision, test_interactions, preserve_rows)


def recall_at_k(model, test_interactions, train_interactions=None, k=10, user_features=None,
                item_features=None, preserve_rows=False, num_threads=1,
                check_intersections=True):
    """Test positives in the top k divided by the user's number of test positives."""
    if hasattr(model, "evaluate_ranks"):
        hits, _, _ = _fused(model, test_interactions, train_interactions, user_features, item_features,
                            num_threads, check_intersections, k=k, hits=True, best_rank=False, auc=False)
        hit = hits.astype(np.float32)
        retrieved = np.squeeze(test_interactions.getnnz(axis=1))
        if not preserve_rows:
            keep = test_interactions.getnnz(axis=1) > 0
            hit, retrieved = hit[keep], retrieved[keep]
        return hit / retrieved
    ranks = _ranks(model, test_interactions, train_interactions, user_features, item_features,
                   num_threads, check_intersections)
    ranks.data = np.less(ranks.data, k, ranks.data)
    retrieved = np.squeeze(test_interactions.getnnz(axis=1))
    hit = np.squeeze(np.array(ranks.sum(axis=1)))
    if not preserve_rows:
        keep = test_interactions.getnnz(axis=1) > 0
        hit = hit[keep]
        retrieved = retrieved[keep]
    return hit / retrieved


def auc_score(model, test_interactions, train_interactions=None, user_features=None,
              item_features=None, preserve_rows=False, num_threads=1, check_intersections=True):
    """Probability that a random test positive outranks a random negative, per user."""
    if hasattr(model, "evaluate_ranks"):
        _, _, auc = _fused(model, test_interactions, train_interactions, user_features, item_features,
                           num_threads, check_intersections, hits=False, best_rank=False, auc=True)
        return _row_filter(auc, test_interactions, preserve_rows)
    ranks = _ranks(model, test_interactions, train_interactions, user_features, item_features,
                   num_threads, check_intersections)
    assert np.all(ranks.data >= 0)
    auc = np.zeros(ranks.shape[0], dtype=np.float32)
    if train_interactions is not None:
        num_train_positives = np.squeeze(
            np.array(train_interactions.getnnz(axis=1)).astype(np.int32))
    else:
        num_train_positives = np.zeros(test_interactions.shape[0], dtype=np.int32)
    # ranks.data is sorted in place per row by the native call (E:244-249).
    calculate_auc_from_rank(CSRMatrix(ranks), np.ascontiguousarray(num_train_positives),
                            ranks.data, auc, num_threads)
    return _row_filter(auc, test_interactions, preserve_rows)


def reciprocal_rank(model, test_interactions, train_interactions=None, user_features=None,
                    item_features=None, preserve_rows=False, num_threads=1,
                    check_intersections=True):
    """1 / (rank of the best-ranked test positive + 1), per user."""
    if hasattr(model, "evaluate_ranks"):
        _, best, _ = _fused(model, test_interactions, train_interactions, user_features, item_features,
                            num_threads, check_intersections, hits=False, best_rank=True, auc=False)
        den = np.where(best >= 0, best + np.float32(1.0), np.float32(1.0)).astype(np.float32)
        rr = np.where(best >= 0, np.float32(1.0) / den, np.float32(0.0)).astype(np.float32)
        return _row_filter(rr, test_interactions, preserve_rows)
    ranks = _ranks(model, test_interactions, train_interactions, user_features, item_features,
                   num_threads, check_intersections)
    ranks.data = 1.0 / (ranks.data + 1.0)
    ranks = np.squeeze(np.array(ranks.max(axis=1).todense()))
    return _row_filter(ranks, test_interactions, preserve_rows)
