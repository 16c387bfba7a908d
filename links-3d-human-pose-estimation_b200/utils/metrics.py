"""Drop-in for reference ``utils/metrics.py``: the per-pose NumPy metrics (a MATLAB-procrustes port) that the
scripts' validation loops call one pose at a time (eval_h36m.py:86-91).  NumPy in / NumPy out, like the reference.

For whole batches use ``Metrics().pmpjpe_batch(gt, pred)`` (or ``utils.metrics_batch.Metrics().pmpjpe_best``):
the same 'best'-reflection PA-MPJPE for every pose at once on the B200 (in-register 3x3 Jacobi SVD kernel).
"""
import numpy as np


def _as_3xJ(p):
    return p.reshape(3, p.shape[1] // 3) if p.shape[0] == 1 else p


class Metrics:
    def __init__(self, init=0):
        self.init = init

    def mpjpe(self, p_ref, p, scale=False, mean_align=False):
        p, p_ref = _as_3xJ(p), _as_3xJ(p_ref)
        if mean_align:
            p = p - p.mean(axis=1, keepdims=True)
            p_ref = p_ref - p_ref.mean(axis=1, keepdims=True)
        if scale:
            p = p * (np.linalg.norm(p_ref.reshape(-1, 1), ord=2) / np.linalg.norm(p.reshape(-1, 1), ord=2))
        return np.linalg.norm(p - p_ref, axis=0).sum() / p.shape[1]

    def pmpjpe(self, p_ref, p, reflection=False):
        p, p_ref = _as_3xJ(p), _as_3xJ(p_ref)
        _, Z, _ = self.procrustes(p_ref.T, p.T, reflection=reflection)
        return self.mpjpe(p_ref, Z.T)

    def PCK(self, p_ref, p, reflection=False):
        return self.pmpjpe(p_ref, p, reflection=reflection)

    def procrustes(self, X, Y, scaling=True, reflection='best'):
        """Least-squares similarity transform of Y onto X.  Returns (d, Z, tform) like MATLAB's procrustes."""
        n, m = X.shape
        _, my = Y.shape
        muX, muY = X.mean(0), Y.mean(0)
        X0, Y0 = X - muX, Y - muY
        ssX, ssY = (X0 ** 2.).sum(), (Y0 ** 2.).sum()
        normX, normY = np.sqrt(ssX), np.sqrt(ssY)
        X0, Y0 = X0 / normX, Y0 / normY
        if my < m:
            Y0 = np.concatenate((Y0, np.zeros((n, m - my))), 1)
        U, s, Vt = np.linalg.svd(X0.T @ Y0, full_matrices=False)
        V = Vt.T
        T = V @ U.T
        if isinstance(reflection, str) and reflection == 'best':
            pass
        else:
            if bool(reflection) != (np.linalg.det(T) < 0):
                V[:, -1] *= -1
                s[-1] *= -1
                T = V @ U.T
        trace = s.sum()
        if scaling:
            b = trace * normX / normY
            d = 1 - trace ** 2
            Z = normX * trace * (Y0 @ T) + muX
        else:
            b = 1
            d = 1 + ssY / ssX - 2 * trace * normY / normX
            Z = normY * (Y0 @ T) + muX
        if my < m:
            T = T[:my, :]
        c = muX - b * (muY @ T)
        return d, Z, {'rotation': T, 'scale': b, 'translation': c}

    def pmpjpe_batch(self, p_ref, p, num_joints=17):
        """Batched GPU equivalent of looping ``pmpjpe(..., reflection='best')``: arrays [M, 3*J] -> [M]."""
        import torch
        from utils.metrics_batch import Metrics as _MB
        dev = torch.device("cuda")
        a = torch.as_tensor(np.ascontiguousarray(p_ref), dtype=torch.float32, device=dev)
        b = torch.as_tensor(np.ascontiguousarray(p), dtype=torch.float32, device=dev)
        return _MB().pmpjpe_best(a, b, num_joints=num_joints).cpu().numpy()
