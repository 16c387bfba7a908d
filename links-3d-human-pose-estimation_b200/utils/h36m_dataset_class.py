"""Drop-in for reference ``utils/h36m_dataset_class.py``: same class names, constructor arguments, ``__len__`` /
``__getitem__`` sample keys ('p2d_gt' | 'p2d_pred', 'poses_3d') and array layouts ([n, 2*J] = J x then J y; [n, 3*J]
= x, y, z blocks).  The pickle is ``{subject: {'poses_2d': [n, J, 2], 'poses_3d': [n, J, 3]}}``
(reference :13-25).  Flattening and both normalisation modes are vectorised (the reference loops over poses in
Python for the per-pose max normalisation, :27-39); ``links_b200.data`` streams the resulting arrays to the GPU."""
import pickle

import numpy as np
import torch
from torch.utils.data import Dataset

from .helpers import split_data_left_right


def flatten_joints(a):
    """[n, J, d] -> [n, d*J]: coordinate-major rows (all x, then all y(, then all z)), the layout of every script."""
    return np.ascontiguousarray(a.transpose(0, 2, 1)).reshape(a.shape[0], -1)


def max_normalise(two_d):
    """[n, J, 2] -> [n, 2*J]: root-centre each pose and divide by its largest absolute coordinate (reference :27-39)."""
    k = two_d - two_d[:, :1, :]
    k = k / np.abs(k).reshape(k.shape[0], -1).max(axis=1)[:, None, None]
    return flatten_joints(k)


def read_subjects(file_name, subjects, key_2d="poses_2d", key_3d="poses_3d"):
    with open(file_name, "rb") as f:
        data = pickle.load(f)
    two_d = np.concatenate([data[s][key_2d] for s in subjects])
    three_d = np.concatenate([data[s][key_3d] for s in subjects])
    return two_d, three_d


class _PoseData(Dataset):
    KEY_3D = "poses_3d"
    CAST_F32 = False                 # the MPI class stores normalised 2D poses as float32 (reference mpi class :29)

    def __init__(self, file_name, train=False, joints=17, get_pca=False, normalize_func=None, get_2dgt=False,
                 subjects=None):
        self.train = train
        self.get_2dgt = get_2dgt
        two_d, three_d = read_subjects(file_name, subjects, key_3d=self.KEY_3D)
        self.data = {"poses_3d": flatten_joints(three_d).reshape(-1, 3 * joints)}
        if normalize_func:
            p2d = normalize_func(flatten_joints(two_d).reshape(-1, 2 * joints))
            self.data["poses_2d"] = p2d.astype(np.float32) if self.CAST_F32 else p2d
        else:
            self.data["poses_2d"] = self._unnormalised(two_d, joints)
        if get_pca:
            self._fit_pca()

    def _unnormalised(self, two_d, joints):
        return max_normalise(two_d).reshape(-1, 2 * joints)

    def _fit_pca(self):
        from sklearn.decomposition import PCA
        self.pca = PCA()
        self.pca.fit(self.data["poses_2d"])

    def __len__(self):
        return self.data["poses_3d"].shape[0]

    def __getitem__(self, idx):
        if torch.is_tensor(idx):
            idx = idx.tolist()
        return {"p2d_gt" if self.get_2dgt else "p2d_pred": self.data["poses_2d"][idx],
                "poses_3d": self.data["poses_3d"][idx]}


class H36M_Data(_PoseData):
    def __init__(self, file_name, train=False, joints=17, get_pca=False, normalize_func=None, get_2dgt=False,
                 subjects=['S1', 'S5', 'S7', 'S6', 'S8']):
        super().__init__(file_name, train, joints, get_pca, normalize_func, get_2dgt, subjects)

    def _fit_pca(self):                      # left / right part PCAs (reference :43-48)
        from sklearn.decomposition import PCA
        self.left_pca, self.right_pca = PCA(), PCA()
        left, right = split_data_left_right(torch.tensor(self.data["poses_2d"]))
        self.left_pca.fit(left.numpy())
        self.right_pca.fit(right.numpy())


class MPI_INF_3DHP_Dataset(_PoseData):
    KEY_3D = "poses_3d_univ"
    CAST_F32 = True

    def __init__(self, file_name, train=False, joints=17, get_pca=False, normalize_func=None, get_2dgt=False,
                 subjects=['S1', 'S2', 'S3', 'S4', 'S5', 'S6', 'S7', 'S8']):
        super().__init__(file_name, train, joints, get_pca, normalize_func, get_2dgt, subjects)

    def _unnormalised(self, two_d, joints):  # this file's MPI class only flattens (reference :30-32)
        return flatten_joints(two_d).reshape(-1, 2 * joints)


class H36M_Data_Original_PCA(_PoseData):
    def __init__(self, file_name, train=False, joints=17, get_pca=False, normalize_func=None, get_2dgt=False,
                 subjects=['S1', 'S5', 'S7', 'S6', 'S8']):
        super().__init__(file_name, train, joints, get_pca, normalize_func, get_2dgt, subjects)
