"""CPU: lightfm_b200.data.Dataset -- known answers (restating the reference's tests/test_data.py)
and, where oracle/_ref is present, equality with the reference's own Dataset on random input."""
import numpy as np
import pytest

import helpers as H
from lightfm_b200.data import Dataset


def test_fit_and_shapes():  # reference tests/test_data.py:9-31
    users, items = 10, 100
    ds = Dataset()
    ds.fit(range(users), range(items))
    assert ds.interactions_shape() == (users, items)
    assert ds.user_features_shape() == (users, users)
    assert ds.item_features_shape() == (items, items)
    assert ds.build_user_features([]).getnnz() == users
    assert ds.build_item_features([]).getnnz() == items
    ds.fit_partial(range(users, 2 * users), range(items, 2 * items))
    assert ds.interactions_shape() == (2 * users, 2 * items)
    assert ds.model_dimensions() == (2 * users, 2 * items)


def test_build_interactions_and_errors():  # tests/test_data.py:34-68
    ds = Dataset()
    ds.fit(range(5), range(7))
    inter, w = ds.build_interactions([(0, 1), (2, 3, 0.5), (4, 6)])
    assert inter.shape == w.shape == (5, 7) and inter.dtype == np.int32 and w.dtype == np.float32
    assert inter.nnz == 3 and np.array_equal(inter.data, [1, 1, 1])
    assert np.allclose(w.data, [1.0, 0.5, 1.0])
    assert np.array_equal(inter.row, w.row) and np.array_equal(inter.col, w.col)
    for bad in ([(9, 1)], [(0, 99)], [(0,)], [(0, 1, 2, 3)]):
        with pytest.raises(ValueError):
            ds.build_interactions(bad)


def test_feature_building_and_normalisation():  # tests/test_data.py:71-115
    ds = Dataset(user_identity_features=False)
    ds.fit(range(3), range(4), user_features=["a", "b"], item_features=["x"])
    assert ds.user_features_shape() == (3, 2) and ds.item_features_shape() == (4, 5)
    uf = ds.build_user_features([(0, ["a"]), (1, {"a": 1.0, "b": 3.0}), (2, ["b"])])
    assert np.allclose(uf.toarray(), [[1, 0], [0.25, 0.75], [0, 1]])
    raw = ds.build_user_features([(0, ["a"]), (1, {"a": 1.0, "b": 3.0}), (2, ["b"])], normalize=False)
    assert np.allclose(raw.toarray(), [[1, 0], [1, 3], [0, 1]])
    with pytest.raises(ValueError):  # a user without features cannot be normalised
        ds.build_user_features([(0, ["a"])])
    with pytest.raises(ValueError):
        ds.build_user_features([(0, ["nope"])], normalize=False)
    with pytest.raises(ValueError):
        ds.build_user_features([(17, ["a"])], normalize=False)
    itf = ds.build_item_features([(1, ["x"])])
    assert np.allclose(itf.toarray()[1], [0, 0.5, 0, 0, 0.5])
    assert uf.dtype == np.float32 and itf.format == "csr"


def test_equal_to_reference_dataset_on_random_input():
    import oracle
    if not oracle.reference_available("strict"):
        pytest.skip("oracle/_ref not built")
    import importlib
    oracle.load_reference("strict")
    RefDataset = importlib.import_module("lightfm.data").Dataset
    rng = np.random.default_rng(0)
    users = ["u%d" % i for i in range(40)]
    items = ["i%d" % i for i in range(30)]
    ufeat = ["f%d" % i for i in range(6)]
    ifeat = ["g%d" % i for i in range(5)]
    inter = [(users[rng.integers(40)], items[rng.integers(30)], float(rng.random())) for _ in range(300)]
    udata = [(u, {ufeat[j]: float(rng.random() + 0.1) for j in rng.choice(6, 2, replace=False)}) for u in users]
    idata = [(i, [ifeat[j] for j in rng.choice(5, 2, replace=False)]) for i in items]
    outs = []
    for cls in (RefDataset, Dataset):
        ds = cls()
        ds.fit(users, items, user_features=ufeat, item_features=ifeat)
        a, w = ds.build_interactions(inter)
        outs.append((a, w, ds.build_user_features(udata), ds.build_item_features(idata, normalize=False),
                     ds.mapping(), ds.model_dimensions()))
    r, o = outs
    for k in range(2):
        assert np.array_equal(r[k].row, o[k].row) and np.array_equal(r[k].col, o[k].col)
        assert np.array_equal(r[k].data, o[k].data) and r[k].dtype == o[k].dtype
    for k in (2, 3):
        assert np.array_equal(r[k].toarray(), o[k].toarray())
    assert r[4] == o[4] and r[5] == o[5]
