"""Drop-in for reference ``utils/mpi_inf_3dhp_dataset_class.py`` (six training subjects, 'poses_3d_univ' targets,
per-pose max normalisation when no normalize_func is given; reference :8-44)."""
from .h36m_dataset_class import _PoseData


class MPI_INF_3DHP_Dataset(_PoseData):
    KEY_3D = "poses_3d_univ"
    CAST_F32 = True

    def __init__(self, file_name, train=False, joints=17, get_pca=False, normalize_func=None, get_2dgt=False,
                 subjects=['S1', 'S2', 'S3', 'S4', 'S5', 'S6']):
        super().__init__(file_name, train, joints, get_pca, normalize_func, get_2dgt, subjects)
