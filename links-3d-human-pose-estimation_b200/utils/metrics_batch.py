"""Drop-in for reference ``utils/metrics_batch.py`` on the sm_100a metric kernels (csrc/metrics.cuh): same method
names, argument order ``(p_ref, p, ...)`` and defaults (``root_joint=6, num_joints=16`` etc.).  Tensors in,
tensors out, same device.  CUDA only (no CPU fallback)."""
import torch

from links_b200 import _cabi


def _prep(t, num_joints):
    if not t.is_cuda:
        raise _cabi.LinksError("links_b200 metrics run on a B200 only (no CPU fallback)")
    t = t.reshape(-1, 3 * num_joints).contiguous().float()
    return t.clone() if t.data_ptr() % 16 else t      # the kernels stream 16-byte aligned chunks (views may start anywhere)


def _st():
    return torch.cuda.current_stream().cuda_stream


class Metrics:
    def __init__(self, init=0):
        self.init = init

    # ---- shared: per-joint distances after root-centring / optional scale matching (:10-22)
    def _dist(self, p_ref, p, use_scaling, root_joint, num_joints, want_dist=True):
        r, q = _prep(p_ref, num_joints), _prep(p, num_joints)
        M = r.shape[0]
        per = torch.empty(M, device=r.device)
        mx = torch.empty(M, device=r.device)
        dist = torch.empty(M, num_joints, device=r.device) if want_dist else None
        _cabi.check(_cabi.lib().links_mpjpe(r.data_ptr(), q.data_ptr(), M, num_joints, root_joint, 1 if use_scaling else 0,
                                            per.data_ptr(), mx.data_ptr(), dist.data_ptr() if want_dist else None, None,
                                            _st()), "links_mpjpe")
        return per, mx, dist

    @staticmethod
    def _counts(values, thresholds, strict=True):
        thr = thresholds.to(values.device, torch.float32).contiguous()
        cnt = torch.zeros(thr.numel(), dtype=torch.int64, device=values.device)
        v = values.contiguous()
        for i in range(0, thr.numel(), 512):
            n = min(512, thr.numel() - i)
            _cabi.check(_cabi.lib().links_threshold_counts(v.data_ptr(), v.numel(), thr[i:].data_ptr(), n, 1 if strict else 0,
                                                           cnt[i:].data_ptr(), _st()), "links_threshold_counts")
        return cnt

    def mpjpe(self, p_ref, p, use_scaling=True, root_joint=6, num_joints=16):
        return self._dist(p_ref, p, use_scaling, root_joint, num_joints, want_dist=False)[0]

    def PCK(self, p_ref, p, use_scaling=True, root_joint=6, num_joints=16, thresh=150.0):
        _, _, d = self._dist(p_ref, p, use_scaling, root_joint, num_joints)
        c = self._counts(d, torch.tensor([thresh]))
        return c[0] / (d.shape[0] * num_joints) * 100

    def AUC(self, p_ref, p, use_scaling=True, root_joint=6, num_joints=16):
        _, _, d = self._dist(p_ref, p, use_scaling, root_joint, num_joints)
        c = self._counts(d, torch.linspace(0, 150, 150))
        return (c / (d.shape[0] * d.shape[1] * 150)).sum()

    def get_all(self, p_ref, p, use_scaling=True, root_joint=0, num_joints=17):
        per, mx, d = self._dist(p_ref, p, use_scaling, root_joint, num_joints)
        n = d.shape[0]
        out = {'MPJPE': d.mean(), 'PCK': self._counts(d, torch.tensor([150.0]))[0] / (n * num_joints) * 100}
        out['AUC'] = (self._counts(d, torch.linspace(0, 150, 31)) / (n * num_joints * 31)).sum() * 100
        # CPS as the reference computes it (marked "not correct" there): poses whose every joint is within d
        out['CPS'] = (self._counts(mx, torch.linspace(0, 300, 301), strict=False) / n).sum()
        return out

    def _pa(self, p_ref, p, num_joints, mode, want_aligned=False):
        r, q = _prep(p_ref, num_joints), _prep(p, num_joints)
        M = r.shape[0]
        per = torch.empty(M, device=r.device)
        al = torch.empty(M, 3 * num_joints, device=r.device) if want_aligned else None
        _cabi.check(_cabi.lib().links_pmpjpe(r.data_ptr(), q.data_ptr(), M, num_joints, mode, per.data_ptr(),
                                             al.data_ptr() if want_aligned else None, None, _st()), "links_pmpjpe")
        return per, al

    def pmpjpe(self, p_ref, p, use_reflection=False, num_joints=16):
        """RMS-scale-matched, rotation-only Procrustes error (:104-159); like the reference, `use_reflection`
        is accepted and ignored."""
        return self._pa(p_ref, p, num_joints, 0)[0]

    def pmpjpe_best(self, p_ref, p, num_joints=17):
        """PA-MPJPE with optimal scale and reflection allowed == reference utils/metrics.py pmpjpe(reflection='best')."""
        return self._pa(p_ref, p, num_joints, 1)[0]

    def procrustes(self, poses_inp, template_poses, use_reflection=False, use_scaling=True):
        if use_reflection or not use_scaling:
            raise NotImplementedError("the accelerated path implements the defaults every caller uses "
                                      "(use_reflection=False, use_scaling=True)")
        nj = int(poses_inp.shape[-1])
        _, al = self._pa(template_poses, poses_inp, nj, 0, want_aligned=True)
        return al.reshape(-1, 3, nj)
