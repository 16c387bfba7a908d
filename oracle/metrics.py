"""Oracle (test infrastructure): pose metrics.

Restates reference ``utils/metrics_batch.py`` (torch, batched; M1-M4 in SURVEY 8a) and
``utils/metrics.py`` (numpy per-pose MATLAB-procrustes port; M5 -- the PA-MPJPE the
scripts actually report, eval_h36m.py:86-91).
"""
import numpy as np
import torch


def _normalise(p_ref, p, use_scaling, root_joint, num_joints):
    # metrics_batch.py:10-20 (identical prologue in mpjpe/PCK/AUC/get_all)
    p = p.reshape(-1, 3, num_joints)
    p_ref = p_ref.reshape(-1, 3, num_joints)
    p = p - p[:, :, root_joint:root_joint + 1]
    p_ref = p_ref - p_ref[:, :, root_joint:root_joint + 1]
    if use_scaling:
        scale_p = p.reshape(-1, 3 * num_joints).norm(p=2, dim=1, keepdim=True)
        scale_p_ref = p_ref.reshape(-1, 3 * num_joints).norm(p=2, dim=1, keepdim=True)
        p = (p.reshape(-1, 3 * num_joints) * (scale_p_ref / scale_p)).reshape(-1, 3, num_joints)
    return p_ref, p


def mpjpe(p_ref, p, use_scaling=True, root_joint=6, num_joints=16):
    """metrics_batch.py:8-24 -> [M]."""
    p_ref, p = _normalise(p_ref, p, use_scaling, root_joint, num_joints)
    return (p - p_ref).norm(p=2, dim=1).mean(axis=1)


def pck(p_ref, p, use_scaling=True, root_joint=6, num_joints=16, thresh=150.0):
    """metrics_batch.py:26-42."""
    p_ref, p = _normalise(p_ref, p, use_scaling, root_joint, num_joints)
    return ((p - p_ref).norm(dim=1) < thresh).sum() / (p_ref.shape[0] * num_joints) * 100


def auc(p_ref, p, use_scaling=True, root_joint=6, num_joints=16):
    """metrics_batch.py:44-64."""
    p_ref, p = _normalise(p_ref, p, use_scaling, root_joint, num_joints)
    d = (p - p_ref).norm(dim=1)
    err = 0
    for t in torch.linspace(0, 150, 150):
        err = err + (d < t).sum() / (d.shape[0] * d.shape[1] * 150)
    return err


def get_all(p_ref, p, use_scaling=True, root_joint=0, num_joints=17):
    """metrics_batch.py:66-102."""
    p_ref, p = _normalise(p_ref, p, use_scaling, root_joint, num_joints)
    d = (p - p_ref).norm(dim=1)
    out = {"MPJPE": d.mean(), "PCK": (d < 150.0).sum() / (p_ref.shape[0] * num_joints) * 100}
    a = 0
    for t in torch.linspace(0, 150, 31):
        a = a + (d < t).sum() / (d.shape[0] * d.shape[1] * 31)
    out["AUC"] = a * 100
    cp = [((d > t).sum(axis=1) < 1).sum() / d.shape[0] for t in torch.linspace(0, 300, 301)]
    out["CPS"] = torch.Tensor(cp).sum()
    return out


def procrustes_batch(poses_inp, template_poses, use_reflection=False, use_scaling=True):
    """metrics_batch.py:116-159.  [M,3,J] each."""
    nj = int(poses_inp.shape[-1])
    mu_t = template_poses.mean(axis=2, keepdims=True)
    tc = template_poses - mu_t
    scale_t = torch.sqrt((tc ** 2).sum(axis=[1, 2], keepdim=True) / (3 * nj))
    ts = tc / scale_t
    mu_p = poses_inp.mean(axis=2, keepdims=True)
    pc = poses_inp - mu_p
    scale_p = torch.sqrt((pc ** 2).sum(axis=[1, 2], keepdim=True) / (3 * nj))
    ps = pc / scale_p
    U, S, Vh = torch.linalg.svd(torch.matmul(ts, ps.transpose(2, 1)))
    R = torch.matmul(U, Vh)                       # U @ V^T
    if not use_reflection:
        Z = torch.eye(3, dtype=R.dtype).repeat(R.shape[0], 1, 1)
        Z[:, -1, -1] *= torch.linalg.det(R)
        R = Z.matmul(R)                           # left-multiplied, as in the reference (:147)
    out = torch.matmul(R, ps)
    if use_scaling:
        out = out * scale_t
    return out + mu_t


def pmpjpe_batch(p_ref, p, use_reflection=False, num_joints=16):
    """metrics_batch.py:104-114 (ignores its own use_reflection arg, like the reference)."""
    p = p.reshape(-1, 3, num_joints)
    p_ref = p_ref.reshape(-1, 3, num_joints)
    aligned = procrustes_batch(p, p_ref)
    return (p_ref - aligned).norm(p=2, dim=1).mean(axis=1)


def procrustes_np(X, Y, scaling=True, reflection="best"):
    """metrics.py:62-171 (X, Y: [J,3] numpy).  Returns (d, Z, tform)."""
    muX, muY = X.mean(0), Y.mean(0)
    X0, Y0 = X - muX, Y - muY
    ssX, ssY = (X0 ** 2.).sum(), (Y0 ** 2.).sum()
    normX, normY = np.sqrt(ssX), np.sqrt(ssY)
    X0 = X0 / normX
    Y0 = Y0 / normY
    A = np.dot(X0.T, Y0)
    U, s, Vt = np.linalg.svd(A, full_matrices=False)
    V = Vt.T
    T = np.dot(V, U.T)
    if reflection != "best":
        have_reflection = np.linalg.det(T) < 0
        if reflection != have_reflection:
            V[:, -1] *= -1
            s[-1] *= -1
            T = np.dot(V, U.T)
    traceTA = s.sum()
    if scaling:
        b = traceTA * normX / normY
        d = 1 - traceTA ** 2
        Z = normX * traceTA * np.dot(Y0, T) + muX
    else:
        b = 1
        d = 1 + ssY / ssX - 2 * traceTA * normY / normX
        Z = normY * np.dot(Y0, T) + muX
    c = muX - b * np.dot(muY, T)
    return d, Z, {"rotation": T, "scale": b, "translation": c}


def pmpjpe_best_np(p_ref, p, reflection="best"):
    """metrics.py:35-46 + :8-33 for one pose given as [1,3J] or [3,J] numpy arrays."""
    if p.shape[0] == 1:
        p = p.reshape(3, p.shape[1] // 3)
    if p_ref.shape[0] == 1:
        p_ref = p_ref.reshape(3, p_ref.shape[1] // 3)
    _, Z, _ = procrustes_np(p_ref.T, p.T, reflection=reflection)
    Zt = Z.T
    return float(np.linalg.norm(Zt - p_ref, axis=0).sum() / p_ref.shape[1])


def pmpjpe_best_batch(p_ref, p, num_joints=17):
    """Vectorised fp64 equivalent of looping ``pmpjpe_best_np`` (eval_h36m.py:83-93) -> [M]."""
    X = np.asarray(p_ref, dtype=np.float64).reshape(-1, 3, num_joints).transpose(0, 2, 1)
    Y = np.asarray(p, dtype=np.float64).reshape(-1, 3, num_joints).transpose(0, 2, 1)
    muX = X.mean(1, keepdims=True)
    X0 = X - muX
    Y0 = Y - Y.mean(1, keepdims=True)
    normX = np.sqrt((X0 ** 2).sum((1, 2), keepdims=True))
    normY = np.sqrt((Y0 ** 2).sum((1, 2), keepdims=True))
    X0 = X0 / normX
    Y0 = Y0 / normY
    A = np.matmul(X0.transpose(0, 2, 1), Y0)
    U, s, Vt = np.linalg.svd(A)
    T = np.matmul(Vt.transpose(0, 2, 1), U.transpose(0, 2, 1))
    Z = normX * s.sum(1)[:, None, None] * np.matmul(Y0, T) + muX
    return np.linalg.norm(Z - X, axis=2).mean(1)
