"""CPU: dataset drop-ins (reference utils/h36m_dataset_class.py / mpi_inf_3dhp_dataset_class.py) and the sharded
array loader.  The reference classes cannot be imported (they import a non-existent AAAI_Code package), so the
expected arrays are written out here with the per-pose loops the reference uses (:13-41)."""
import pickle

import numpy as np
import pytest
import torch

from links_b200.data import ArrayLoader, loader_from_dataset
from links_b200.shard import shard_bounds
from utils.h36m_dataset_class import H36M_Data, H36M_Data_Original_PCA, MPI_INF_3DHP_Dataset as MPI8
from utils.helpers import normalize_head, normalize_head_test
from utils.mpi_inf_3dhp_dataset_class import MPI_INF_3DHP_Dataset as MPI6


def make_pickle(path, subjects, key3d, seed=0):
    rng = np.random.RandomState(seed)
    data = {}
    for i, s in enumerate(subjects):
        n = 5 + i
        data[s] = {"poses_2d": rng.normal(size=(n, 17, 2)) * 100 + 500, key3d: rng.normal(size=(n, 17, 3)) * 300}
    with open(path, "wb") as f:
        pickle.dump(data, f)
    return data


def expected(data, subjects, key3d, normalize_func):
    two = np.concatenate([data[s]["poses_2d"] for s in subjects])
    three = np.concatenate([data[s][key3d] for s in subjects])
    p3 = three.transpose(0, 2, 1).reshape(-1, 51)
    if normalize_func:
        p2 = normalize_func(two.transpose(0, 2, 1).reshape(-1, 34).copy())
    else:
        rows = []
        for t in range(len(two)):               # the reference's per-pose loop
            k = two[t] - two[t][0]
            rows.append((k / np.max(abs(k))).transpose(1, 0).reshape(-1, 34))
        p2 = np.concatenate(rows)
    return p2, p3


@pytest.mark.parametrize("cls,subjects,key3d", [(H36M_Data, ['S1', 'S5', 'S7', 'S6', 'S8'], "poses_3d"),
                                                (H36M_Data_Original_PCA, ['S1', 'S5', 'S7', 'S6', 'S8'], "poses_3d"),
                                                (MPI6, ['S1', 'S2', 'S3', 'S4', 'S5', 'S6'], "poses_3d_univ")])
@pytest.mark.parametrize("norm", [None, normalize_head, normalize_head_test])
def test_dataset_dropins(tmp_path, cls, subjects, key3d, norm):
    path = str(tmp_path / "d.pkl")
    data = make_pickle(path, subjects, key3d)
    ds = cls(path, train=True, normalize_func=norm, get_2dgt=True)
    p2, p3 = expected(data, subjects, key3d, norm)
    assert len(ds) == p3.shape[0]
    np.testing.assert_array_equal(ds.data["poses_3d"], p3)
    np.testing.assert_allclose(ds.data["poses_2d"], p2, rtol=1e-6 if cls is MPI6 and norm else 1e-12)
    s = ds[3]
    assert set(s) == {"p2d_gt", "poses_3d"} and s["p2d_gt"].shape == (34,) and s["poses_3d"].shape == (51,)
    assert "p2d_pred" in cls(path, normalize_func=norm, get_2dgt=False)[0]


def test_mpi_class_of_h36m_file_only_flattens(tmp_path):
    subjects = ['S1', 'S2', 'S3', 'S4', 'S5', 'S6', 'S7', 'S8']
    path = str(tmp_path / "m.pkl")
    data = make_pickle(path, subjects, "poses_3d_univ")
    ds = MPI8(path)
    two = np.concatenate([data[s]["poses_2d"] for s in subjects])
    np.testing.assert_array_equal(ds.data["poses_2d"], two.transpose(0, 2, 1).reshape(-1, 34))


def test_pca_option(tmp_path):
    path = str(tmp_path / "d.pkl")
    make_pickle(path, ['S1', 'S5', 'S7', 'S6', 'S8'], "poses_3d")
    ds = H36M_Data(path, normalize_func=normalize_head, get_pca=True)
    assert ds.left_pca.components_.shape[1] == 22 and ds.right_pca.components_.shape[1] == 22


def test_array_loader_shards_and_batches(tmp_path):
    path = str(tmp_path / "d.pkl")
    make_pickle(path, ['S1', 'S5', 'S7', 'S6', 'S8'], "poses_3d")
    ds = H36M_Data(path, normalize_func=normalize_head, get_2dgt=True)
    n = len(ds)                                             # 35 poses
    seen = []
    for rank in range(2):
        ld = loader_from_dataset(ds, 8, rank, 2, seed=3)
        b, e = shard_bounds(n, rank, 2)
        assert ld.batch == 4 and len(ld) == (e - b) // 4
        rows = torch.cat(list(ld))
        assert rows.shape == (len(ld) * 4, 34) and rows.dtype == torch.float32
        shard = torch.as_tensor(ds.data["poses_2d"][b:e], dtype=torch.float32)
        # every yielded row is a row of this rank's shard, none twice within an epoch
        idx = [(shard == r).all(1).nonzero().item() for r in rows]
        assert len(set(idx)) == len(idx)
        seen.append(set(i + b for i in idx))
    assert not (seen[0] & seen[1])
    with pytest.raises(ValueError):
        ArrayLoader(ds.data["poses_2d"], None, 6, 0, 2)       # odd per-rank batch
