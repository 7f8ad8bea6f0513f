# coding=utf-8
"""Train/test splitting of interaction matrices (reference: lightfm/cross_validation.py:18-80)."""
import numpy as np
import scipy.sparse as sp

__all__ = ["random_train_test_split"]


def random_train_test_split(interactions, test_percentage=0.2, random_state=None):
    """Randomly split the stored entries of ``interactions`` into two disjoint COO matrices.

    The entries are permuted with ``random_state.shuffle(arange(nnz))`` and cut at
    ``int((1 - test_percentage) * nnz)``; shape and dtype are preserved.
    """
    if not sp.issparse(interactions):
        raise ValueError("Interactions must be a scipy.sparse matrix.")
    if not isinstance(random_state, np.random.RandomState):
        random_state = np.random.RandomState(seed=random_state)

    coo = interactions.tocoo()
    order = np.arange(len(coo.row))
    random_state.shuffle(order)
    rows, cols, vals = coo.row[order], coo.col[order], coo.data[order]
    cutoff = int((1.0 - test_percentage) * len(rows))

    def part(sl):
        return sp.coo_matrix((vals[sl], (rows[sl], cols[sl])), shape=coo.shape, dtype=coo.dtype)

    return part(slice(None, cutoff)), part(slice(cutoff, None))
