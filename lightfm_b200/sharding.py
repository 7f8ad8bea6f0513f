"""Multi-GPU layout of the fit_partial hot path: one process per GPU.

The path shards on ONE axis and has ONE exchange per epoch:

* ``axis="item"`` (BASELINE.json north star): item(-feature) rows are hash-sharded --
  rank ``g`` owns items with ``shard_of(i) == g`` (embeddings + accumulators live only
  there) and trains exactly the interactions whose positive item it owns.  Negatives are
  drawn uniformly from the LOCAL shard (a uniform random subset of the catalogue), while
  the WARP rank estimate keeps the GLOBAL item count.  The user table is replicated.
* ``axis="user"``: the mirror image -- users sharded, the (usually much smaller) item
  table replicated, negative sampling stays global.

Exchange: the replicated table's epoch delta is all-reduced (NCCL over NVLink):
``W <- W0 + sum_g (W_g - W0)``, same for the Adagrad accumulators and biases.

Everything here is host-side index work (numpy) plus the collective; the kernels are the
single-GPU ones running on local ids.  Identity features only (feature rows shared across
shards would have to join the replicated block).
"""
import numpy as np
import scipy.sparse as sp

_HASH_MUL = np.uint64(0x9E3779B97F4A7C15)


def shard_of(ids, world):
    """Stateless hash partition of integer ids onto ``world`` ranks (uniform, deterministic)."""
    x = np.asarray(ids).astype(np.uint64)
    with np.errstate(over="ignore"):
        h = (x + np.uint64(1)) * _HASH_MUL
    return ((h >> np.uint64(33)) % np.uint64(world)).astype(np.int32)


class ShardMap(object):
    """Global <-> local id maps of the sharded axis for one rank."""

    def __init__(self, n_global, rank, world):
        self.n_global, self.rank, self.world = int(n_global), int(rank), int(world)
        owner = shard_of(np.arange(n_global, dtype=np.int64), world)
        self.global_ids = np.flatnonzero(owner == rank).astype(np.int32)   # local -> global
        self.local_of = np.full(n_global, -1, dtype=np.int32)              # global -> local
        self.local_of[self.global_ids] = np.arange(len(self.global_ids), dtype=np.int32)
        self.n_local = len(self.global_ids)


def partition(interactions, sample_weight, axis, rank, world):
    """Local training problem of ``rank``.

    Returns ``(local_coo, local_weight, positives_csr, shard_map)`` where the sharded
    axis of ``local_coo`` / ``positives_csr`` uses local ids.  The positives lookup keeps
    only the local items (item axis) or local users (user axis) -- exactly the entries the
    local kernel can ever test.
    """
    coo = interactions.tocoo()
    n_users, n_items = coo.shape
    if axis == "item":
        smap = ShardMap(n_items, rank, world)
        keep = smap.local_of[coo.col] >= 0
        rows = coo.row[keep].astype(np.int32)
        cols = smap.local_of[coo.col[keep]]
        shape = (n_users, smap.n_local)
    elif axis == "user":
        smap = ShardMap(n_users, rank, world)
        keep = smap.local_of[coo.row] >= 0
        rows = smap.local_of[coo.row[keep]]
        cols = coo.col[keep].astype(np.int32)
        shape = (smap.n_local, n_items)
    else:
        raise ValueError("axis must be 'item' or 'user'")
    data = coo.data[keep].astype(np.float32)
    local = sp.coo_matrix((data, (rows, cols)), shape=shape, dtype=np.float32)
    weight = None if sample_weight is None else np.asarray(sample_weight)[keep].astype(np.float32)
    positives = local.tocsr()
    positives.sort_indices()
    return local, weight, positives, smap


def slice_state(arrays, axis, smap):
    """Rows of the 12 state arrays this rank keeps: the sharded side is cut to the local ids,
    the other side is replicated whole."""
    side = "item" if axis == "item" else "user"
    out = {}
    for k, v in arrays.items():
        out[k] = np.ascontiguousarray(v[smap.global_ids]) if k.startswith(side) else v.copy()
    return out


def merge_state(global_arrays, local_arrays, axis, smap):
    """Write a rank's sharded rows back into the global arrays (replicated side: copy)."""
    side = "item" if axis == "item" else "user"
    for k, v in local_arrays.items():
        if k.startswith(side):
            global_arrays[k][smap.global_ids] = v
        else:
            global_arrays[k][...] = v


def allreduce_deltas(tensors, snapshots, group=None):
    """``t <- snapshot + sum_over_ranks(t - snapshot)`` in place for every tensor.

    Works on CPU tensors (gloo) and CUDA tensors (nccl); the tensors are the replicated
    table's arrays (embeddings, accumulators, biases) after a local epoch, the snapshots
    their common value before it.
    """
    import torch.distributed as dist
    for t, s in zip(tensors, snapshots):
        t.sub_(s)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.add_(s)


class CudaArrayView(object):
    """Expose a raw device pointer as ``__cuda_array_interface__`` so torch can wrap it
    (``torch.as_tensor(view, device=...)``) without copying."""

    def __init__(self, ptr, count, dtype="<f4"):
        self.__cuda_array_interface__ = {
            "shape": (int(count),), "typestr": dtype, "data": (int(ptr), False), "version": 2}
