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

Feature matrices (third bullet): with non-identity features on the SHARDED side
(item tags under ``axis="item"``) a feature row can be touched from every shard, so the whole
feature-embedding table of that side joins the replicated block -- every rank keeps all of its
rows, trains on its own entities' rows of the feature matrix, and both tables' deltas are
all-reduced (rows only one shard touches contribute zeros from the others, so the sum is exact).

Everything here is host-side index work (numpy) plus the collective; the kernels are the
single-GPU ones running on local ids.
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


def exchange_replicated(plan, side, device, group=None, world=None):
    """One delta exchange of the replicated table of a resident plan:
    ``W <- W0 + sum_g (W_g - W0)`` for w, g, b, bg of `side` (0 item, 1 user).  Call
    ``plan.delta_begin(side)`` before the local epoch and this afterwards.  The subtract and
    add-back are two sweeps of libfm_cuda's own kernels over ONE packed buffer, which is
    all-reduced in place with a single NCCL call.  Returns the device milliseconds
    ``{"make_ms", "allreduce_ms", "apply_ms"}`` (CUDA events)."""
    import torch
    import torch.distributed as dist
    ptr, count, make_ms = plan.delta_make(side)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if dist.is_initialized() and (world is None or world > 1):
        t = torch.as_tensor(CudaArrayView(ptr, count), device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    e1.record()
    torch.cuda.synchronize(device)
    apply_ms = plan.delta_apply(side)
    return {"make_ms": make_ms, "allreduce_ms": e0.elapsed_time(e1), "apply_ms": apply_ms}


class CudaArrayView(object):
    """Expose a raw device pointer as ``__cuda_array_interface__`` so torch can wrap it
    (``torch.as_tensor(view, device=...)``) without copying."""

    def __init__(self, ptr, count, dtype="<f4"):
        self.__cuda_array_interface__ = {
            "shape": (int(count),), "typestr": dtype, "data": (int(ptr), False), "version": 2}


class ShardedTrainer(object):
    """Hogwild training of one ``LightFM`` problem across the ranks of a torch.distributed
    group (one process per GPU, launched by torchrun).

    Every rank calls it with the SAME global interactions and an identically initialised
    model (same ``random_state``); the trainer keeps this rank's shard of the sharded table and
    a replica of the other table resident in HBM (``ResidentPlan``), runs local epochs with the
    single-GPU kernels and all-reduces the replicated table's delta after every epoch.

        model = LightFM(loss="warp", no_components=64, random_state=0)
        trainer = ShardedTrainer(model, interactions, axis="item")      # under torchrun
        trainer.fit_epochs(10)
        trainer.gather()          # every rank's model arrays now hold the full trained state

    ``item_features`` / ``user_features`` (optional scipy matrices, rows = global entity ids):
    a feature matrix on the sharded side is cut to this rank's entities and its embedding table
    is kept whole and exchanged like the replicated side's (module docstring).
    """

    def __init__(self, model, interactions, axis="item", sample_weight=None, group=None, device=None,
                 item_features=None, user_features=None):
        import torch
        import torch.distributed as dist
        from . import _lightfm_fast as native
        if model.learning_schedule != "adagrad":
            raise NotImplementedError("ShardedTrainer supports the adagrad schedule")
        self.model, self.axis, self.group = model, axis, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        # libfm_cuda defaults to device 0: bind it to this rank's GPU before the plan allocates
        # anything (raises if the library was already initialised on another device)
        native.set_device(self.device.index if self.device.index is not None else torch.cuda.current_device())
        coo = interactions.tocoo()
        if coo.dtype != np.float32:
            coo.data = coo.data.astype(np.float32)
        self.shape = coo.shape
        n_users, n_items = coo.shape
        feats = {"item": item_features, "user": user_features}
        for k, f in feats.items():
            if f is not None:
                f = f.tocsr().astype(np.float32)
                f.sort_indices()
                feats[k] = f
        if model.item_embeddings is None:
            model._initialize(model.no_components,
                              feats["item"].shape[1] if feats["item"] is not None else n_items,
                              feats["user"].shape[1] if feats["user"] is not None else n_users)
        local, weight, positives, self.smap = partition(coo, sample_weight, axis, self.rank, self.world)
        if weight is None:
            weight = local.data if np.array_equiv(local.data, 1.0) else np.ones_like(local.data)
        self._global = {k: getattr(model, k) for k in _STATE_NAMES}
        side_name = "item" if axis == "item" else "user"
        # a feature matrix on the sharded side: its embedding table is not cut (shared rows)
        self.sliced = feats[side_name] is None
        if self.sliced:
            self.local_state = slice_state(self._global, axis, self.smap)
        else:
            self.local_state = {k: v.copy() for k, v in self._global.items()}
        st = self.local_state
        self._holder = native.FastLightFM(
            *[st[k] for k in _STATE_NAMES], model.no_components, 0, model.learning_rate, model.rho,
            model.epsilon, model.max_sampled)
        lu, li = local.shape
        kos = model.loss == "warp-kos"
        self._keep = (local, weight, positives)
        def local_features(name, n_local):
            f = feats[name]
            if f is None:
                return sp.identity(n_local, dtype=np.float32, format="csr")
            if name == side_name:                      # rows of this rank's entities, all feature columns
                f = f[self.smap.global_ids]
                f.sort_indices()
            return f
        self._features = (local_features("item", li), local_features("user", lu))
        self.plan = native.ResidentPlan(
            model.loss, native.CSRMatrix(self._features[0]), native.CSRMatrix(self._features[1]),
            native.CSRMatrix(positives) if model.loss != "logistic" else None,
            np.ascontiguousarray(local.row), None if kos else np.ascontiguousarray(local.col),
            None if kos else local.data, None if kos else weight, self._holder,
            model.item_alpha, model.user_alpha, model.k, model.n)
        if axis == "item":
            self.plan.set_global_items(n_items)
        # lfm_plan_delta_* sides to exchange (0 item, 1 user): the replicated table, plus the sharded
        # side's whole feature table when it has shared rows
        self.exchange_sides = [1 if axis == "item" else 0] + ([] if self.sliced else [0 if axis == "item" else 1])
        self.local_interactions = local.nnz
        self.last_counters = None

    def epoch(self, seed, num_threads=8):
        """One local epoch + the delta all-reduce of the replicated table."""
        begin_ms = sum(self.plan.delta_begin(side) for side in self.exchange_sides) if self.world > 1 else 0.0
        c = self.plan.epoch(seed=(int(seed) * 977 + self.rank) & 0xFFFFFFFF, num_threads=max(2, num_threads))
        c["allreduce_ms"], c["exchange_ms"] = 0.0, begin_ms
        if self.world > 1:
            for side in self.exchange_sides:
                t = exchange_replicated(self.plan, side, self.device, self.group, self.world)
                c["allreduce_ms"] += t["allreduce_ms"]
                c["exchange_ms"] += t["make_ms"] + t["allreduce_ms"] + t["apply_ms"]
        self.last_counters = c
        return c

    def fit_epochs(self, epochs, num_threads=8):
        for _ in range(epochs):
            seed = int(self.model.random_state.randint(0, np.iinfo(np.int32).max))
            self.epoch(seed, num_threads)
        return self

    def gather(self):
        """Write the trained state back into the model's (global) numpy arrays on every rank:
        the replicated side from the local replica, the sharded side by all-gathering shards."""
        import torch
        import torch.distributed as dist
        self.plan.download()
        if not self.sliced:      # both tables are whole and identical on every rank after the exchange
            for k, v in self.local_state.items():
                self._global[k][...] = v
            return self.model
        merge_state(self._global, self.local_state, self.axis, self.smap)
        if self.world > 1:
            side = "item" if self.axis == "item" else "user"
            n_global = self.shape[1] if self.axis == "item" else self.shape[0]
            owner = shard_of(np.arange(n_global, dtype=np.int64), self.world)
            for k in _STATE_NAMES:
                if not k.startswith(side):
                    continue
                arr = self._global[k]
                t = torch.from_numpy(arr).to(self.device)
                mask = torch.from_numpy((owner == self.rank)).to(self.device)
                t = t * (mask.view(-1, *([1] * (t.dim() - 1)))).to(t.dtype)  # zero the rows of other shards
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
                arr[...] = t.cpu().numpy()
        return self.model

    def close(self):
        self.plan.close()


_STATE_NAMES = ("item_embeddings", "item_embedding_gradients", "item_embedding_momentum",
                "item_biases", "item_bias_gradients", "item_bias_momentum",
                "user_embeddings", "user_embedding_gradients", "user_embedding_momentum",
                "user_biases", "user_bias_gradients", "user_bias_momentum")
