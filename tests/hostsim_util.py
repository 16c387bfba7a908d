"""TEST INFRASTRUCTURE: build + bind tests/hostsim/_hostsim.so (the product's .cuh kernels compiled for the CPU).

Same argument lists as links_b200._cabi.SIGNATURES without the stream; all pointers are HOST pointers
(numpy arrays).  Never used by the product.
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SIMDIR = os.path.join(HERE, "hostsim")
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "links-3d-human-pose-estimation_b200", "csrc")
SO = os.path.join(SIMDIR, "_hostsim.so")
_lib = None


def _digest():
    h = hashlib.sha256()
    files = [os.path.join(SIMDIR, f) for f in ("cuda_shim.h", "hostsim_api.cpp")]
    files += [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cuh")]
    files.append(os.path.join(ROOT, "include", "links_b200.h"))
    for f in files:
        h.update(open(f, "rb").read())
    return h.hexdigest()


def sim():
    global _lib
    if _lib is not None:
        return _lib
    from links_b200 import _cabi
    stamp = SO + ".sha256"
    dig = _digest()
    if not (os.path.exists(SO) and os.path.exists(stamp) and open(stamp).read() == dig):
        subprocess.check_call(["g++", "-std=c++20", "-O2", "-pthread", "-shared", "-fPIC", "-fvisibility=hidden",
                               "-I" + SIMDIR, "-o", SO, os.path.join(SIMDIR, "hostsim_api.cpp")])
        open(stamp, "w").write(dig)
    L = C.CDLL(SO)
    for name, (res, args) in _cabi.SIGNATURES.items():
        if name in ("links_flow_nll_train", "links_flow_vjp_train", "links_adam_zero", "links_peer_barrier"):      # tensor-core-only / multi-GPU entry points: no CPU build
            continue
        fn = getattr(L, name.replace("links_", "sim_"))
        fn.restype, fn.argtypes = res, list(args)
    L.sim_flow_packed_floats.restype = C.c_size_t
    L.sim_flow_packed_floats.argtypes = [C.c_int, C.c_int]
    _lib = L
    return L


def ptr(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class View:
    """Element `offset` of `base` as a pointer argument: column-offset views into one packed buffer (the way the lifter
    engine hands the geometry kernels the four head outputs of a pass, all living in one [N, 32] row)."""

    def __init__(self, base, offset):
        assert base.flags["C_CONTIGUOUS"]
        self.base, self.offset = base, int(offset)


def bf16_to_f32(a_u16):
    return (a_u16.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16(a):
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = u + 0x7FFF + ((u >> 16) & 1)
    return (u >> 16).astype(np.uint16)


def pack_flow(backend, params, C_, n_blocks=8):
    """FrEIA-layout param dict (torch tensors) -> packed float32 numpy buffer via the pack kernel of `backend`."""
    names = ("subnet.0.weight", "subnet.0.bias", "subnet.2.weight", "subnet.2.bias", "global_scale",
             "global_offset", "w_perm", "w_perm_inv")
    host = {n: [np.ascontiguousarray(params["module_list.%d.%s" % (k, n)].detach().cpu().numpy(), dtype=np.float32)
                for k in range(n_blocks)] for n in names}
    packed = np.zeros(backend.packed_floats(C_, n_blocks), dtype=np.float32)
    if backend.name == "sim":
        tables = [(C.c_void_p * n_blocks)(*[a.ctypes.data for a in host[n]]) for n in names]
        rc = backend.L.sim_flow_pack(C_, n_blocks, *tables, ptr(packed))
    else:
        import torch
        dev = {n: [torch.from_numpy(a).cuda() for a in host[n]] for n in names}
        tables = [(C.c_void_p * n_blocks)(*[t.data_ptr() for t in dev[n]]) for n in names]
        pk = torch.from_numpy(packed).cuda()
        rc = backend.L.links_flow_pack(C_, n_blocks, *tables, pk.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        packed[...] = pk.cpu().numpy()
    assert rc == 0
    return packed


# ------------------------------------------------------------------------------------------------
# Backends: the same kernel tests run against the CPU simulation ("sim", default suite) and against
# the real CUDA library through the C ABI ("cuda", `-m gpu`).
# ------------------------------------------------------------------------------------------------
class SimBackend:
    name = "sim"

    def __init__(self):
        self.L = sim()

    def call(self, fn, *args):
        conv = []
        for a in args:
            if isinstance(a, View):
                conv.append(C.c_void_p(a.base.ctypes.data + a.offset * a.base.itemsize))
            else:
                conv.append(ptr(a) if isinstance(a, np.ndarray) else a)
        return getattr(self.L, "sim_" + fn)(*conv)

    def packed_floats(self, C_, nb):
        return self.L.sim_flow_packed_floats(C_, nb)


class CudaBackend:
    """Marshals numpy arrays through device memory and calls liblinks_b200.so on the current stream."""
    name = "cuda"

    def __init__(self):
        from links_b200 import _cabi
        self.L = _cabi.lib()

    def call(self, fn, *args):
        import torch
        dev, conv, up = [], [], {}
        for a in args:
            base = a.base if isinstance(a, View) else a
            if isinstance(base, np.ndarray):
                if id(base) not in up:                      # views into one buffer share one device copy
                    up[id(base)] = torch.from_numpy(base).cuda()
                    dev.append((base, up[id(base)]))
                t = up[id(base)]
                conv.append(t.data_ptr() + (a.offset * base.itemsize if isinstance(a, View) else 0))
            else:
                conv.append(a)
        stream = torch.cuda.current_stream().cuda_stream
        rc = getattr(self.L, "links_" + fn)(*conv, stream)
        torch.cuda.synchronize()
        for a, t in dev:
            a[...] = t.cpu().numpy()
        return rc

    def packed_floats(self, C_, nb):
        return self.L.links_flow_packed_floats(C_, nb)


def _make_backend(kind):
    return SimBackend() if kind == "sim" else CudaBackend()


import pytest  # noqa: E402

_backend_cache = {}


@pytest.fixture
def backend(request):
    kind = request.param
    if kind not in _backend_cache:
        _backend_cache[kind] = _make_backend(kind)
    return _backend_cache[kind]


backend_params = pytest.mark.parametrize(
    "backend", ["sim", pytest.param("cuda", marks=pytest.mark.gpu)], indirect=True)
