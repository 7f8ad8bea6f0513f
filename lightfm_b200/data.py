# coding=utf-8
"""Build interaction and feature matrices from (id, id[, weight]) tuples and feature dicts.

Host-side ETL that feeds the hot path; same public surface and
behaviour as the reference's ``lightfm/data.py`` (``D:``): ``Dataset.fit`` / ``fit_partial``
D:190-257, ``build_interactions`` D:296-330, ``build_user_features`` /
``build_item_features`` D:345-422 (optional L1 row normalisation, D:124-131),
``mapping`` / ``*_shape`` / ``model_dimensions`` D:289-449.  Written from scratch: entries are
collected into Python lists and turned into scipy matrices in one step.
"""
import numpy as np
import scipy.sparse as sp
import sklearn.preprocessing

__all__ = ["Dataset"]


def _index_of(mapping, key):
    """Index of `key`, assigning the next free index on first sight."""
    return mapping.setdefault(key, len(mapping))


class Dataset(object):
    """Maps arbitrary hashable user / item / feature ids to contiguous indices and builds the
    sparse matrices ``LightFM`` consumes.

    Parameters
    ----------
    user_identity_features, item_identity_features : bool
        Give every user / item its own indicator feature in addition to the supplied ones
        (the feature matrix then starts with an identity block).
    """

    def __init__(self, user_identity_features=True, item_identity_features=True):
        self._user_identity_features = user_identity_features
        self._item_identity_features = item_identity_features
        self._reset()

    def _reset(self):
        self._user_id_mapping = {}
        self._item_id_mapping = {}
        self._user_feature_mapping = {}
        self._item_feature_mapping = {}

    def _check_fitted(self):
        if not self._user_id_mapping or not self._item_id_mapping:
            raise ValueError("You must call fit first to build the item and user id mappings.")

    # ---- id / feature vocabularies ------------------------------------------------------
    def fit(self, users, items, user_features=None, item_features=None):
        """Forget previous mappings, then ``fit_partial``."""
        self._reset()
        return self.fit_partial(users, items, user_features, item_features)

    def fit_partial(self, users=None, items=None, user_features=None, item_features=None):
        """Extend the mappings with new ids / feature names (existing indices are kept)."""
        for ids, id_map, feat_map, identity in (
                (users, self._user_id_mapping, self._user_feature_mapping, self._user_identity_features),
                (items, self._item_id_mapping, self._item_feature_mapping, self._item_identity_features)):
            if ids is None:
                continue
            for entity in ids:
                _index_of(id_map, entity)
                if identity:
                    _index_of(feat_map, entity)
        for names, feat_map in ((user_features, self._user_feature_mapping),
                                (item_features, self._item_feature_mapping)):
            if names is not None:
                for name in names:
                    _index_of(feat_map, name)

    # ---- interactions -----------------------------------------------------------------------
    def interactions_shape(self):
        """(number of users, number of items)."""
        return (len(self._user_id_mapping), len(self._item_id_mapping))

    def build_interactions(self, data):
        """``data``: iterable of (user_id, item_id) or (user_id, item_id, weight).

        Returns ``(interactions, weights)``: COO int32 matrix of ones and COO float32 matrix of
        the weights, entries in input order (so the pair is accepted by
        ``LightFM.fit(..., sample_weight=weights)``)."""
        rows, cols, weights = [], [], []
        for datum in data:
            if len(datum) == 3:
                user_id, item_id, weight = datum
            elif len(datum) == 2:
                (user_id, item_id), weight = datum, 1.0
            else:
                raise ValueError("Expecting tuples of (user_id, item_id, weight) "
                                 "or (user_id, item_id). Got {}".format(datum))
            user_idx = self._user_id_mapping.get(user_id)
            item_idx = self._item_id_mapping.get(item_id)
            if user_idx is None:
                raise ValueError("User id {} not in user id mapping. Make sure "
                                 "you call the fit method.".format(user_id))
            if item_idx is None:
                raise ValueError("Item id {} not in item id mapping. Make sure "
                                 "you call the fit method.".format(item_id))
            rows.append(user_idx)
            cols.append(item_idx)
            weights.append(weight)
        shape = self.interactions_shape()
        r = np.asarray(rows, dtype=np.int32)
        c = np.asarray(cols, dtype=np.int32)
        interactions = sp.coo_matrix((np.ones(len(rows), dtype=np.int32), (r, c)), shape=shape)
        weight_mat = sp.coo_matrix((np.asarray(weights, dtype=np.float32), (r, c)), shape=shape)
        return interactions, weight_mat

    # ---- features --------------------------------------------------------------------------------
    def user_features_shape(self):
        return (len(self._user_id_mapping), len(self._user_feature_mapping))

    def item_features_shape(self):
        return (len(self._item_id_mapping), len(self._item_feature_mapping))

    def _build_features(self, data, id_map, feat_map, identity, normalize, kind):
        rows, cols, vals = [], [], []
        if identity:
            for entity, idx in id_map.items():
                rows.append(idx)
                cols.append(feat_map[entity])
                vals.append(1.0)
        for datum in data:
            if len(datum) != 2:
                raise ValueError("Expected tuples of ({}_id, features), got {}.".format(kind, datum))
            entity, features = datum
            if entity not in id_map:
                raise ValueError("{kind} id {eid} not in {kind} id mappings.".format(kind=kind, eid=entity))
            idx = id_map[entity]
            pairs = features.items() if isinstance(features, dict) else ((name, 1.0) for name in features)
            for name, weight in pairs:
                if name not in feat_map:
                    raise ValueError("Feature {} not in feature mapping. Call fit first.".format(name))
                rows.append(idx)
                cols.append(feat_map[name])
                vals.append(weight)
        shape = (len(id_map), len(feat_map))
        mat = sp.coo_matrix((np.asarray(vals, dtype=np.float32),
                             (np.asarray(rows, dtype=np.int32), np.asarray(cols, dtype=np.int32))),
                            shape=shape).tocsr()
        if normalize:
            if np.any(mat.getnnz(1) == 0):
                raise ValueError("Cannot normalize feature matrix: some rows have zero norm. "
                                 "Ensure that features were provided for all entries.")
            sklearn.preprocessing.normalize(mat, norm="l1", copy=False)
        return mat

    def build_user_features(self, data, normalize=True):
        """``data``: iterable of (user_id, [feature names]) or (user_id, {feature: weight}).
        Returns a CSR float32 matrix [users, user features], rows L1-normalised if asked."""
        return self._build_features(data, self._user_id_mapping, self._user_feature_mapping,
                                    self._user_identity_features, normalize, "user")

    def build_item_features(self, data, normalize=True):
        """Same as ``build_user_features`` for items."""
        return self._build_features(data, self._item_id_mapping, self._item_feature_mapping,
                                    self._item_identity_features, normalize, "item")

    # ---- introspection ---------------------------------------------------------------------------------
    def model_dimensions(self):
        """(number of user features, number of item features): the embedding table heights."""
        return (len(self._user_feature_mapping), len(self._item_feature_mapping))

    def mapping(self):
        """(user id map, user feature map, item id map, item feature map)."""
        return (self._user_id_mapping, self._user_feature_mapping,
                self._item_id_mapping, self._item_feature_mapping)
