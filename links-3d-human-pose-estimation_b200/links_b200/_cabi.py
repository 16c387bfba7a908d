"""ctypes binding of liblinks_b200.so (include/links_b200.h).

The product path has no CPU fallback: ``lib()`` raises if the CUDA library is missing or the
device is not a compute-capability-10.x GPU.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LINKS_B200_LIB", os.path.join(_HERE, "_lib", "liblinks_b200.so"))  # env: A/B builds only

HEAD_LD = 32
KPAD = 64
MAX_GEMM_PROBLEMS = 8

EPI_LEAKY_PRE, EPI_LEAKY_POST, EPI_RELU_PRE, EPI_ACCUM_F32 = 1, 2, 4, 8
GEMM_A_MN, GEMM_B_MN = 16, 32
EPI_YMASK_ZERO = 64

vp = C.c_void_p
ci = C.c_int
cf = C.c_float
sz = C.c_size_t


class GemmProblem(C.Structure):
    _fields_ = [("A", vp), ("B", vp), ("M", ci), ("N", ci), ("K", ci), ("lda", ci), ("ldb", ci),
                ("flags", C.c_uint32), ("bias", vp),
                ("add0", vp), ("ld_add0", ci), ("add1", vp), ("ld_add1", ci),
                ("ymask", vp), ("ld_ymask", ci), ("bits", vp), ("ld_bits", ci),
                ("sign_out", vp), ("ld_sign", ci), ("mid", vp), ("ld_mid", ci),
                ("out", vp), ("ld_out", ci), ("out_f32", vp), ("ld_f32", ci),
                ("adam_p", vp), ("adam_m", vp), ("adam_v", vp), ("adam_shadow", vp), ("ld_shadow", ci), ("adam_hyper", vp),
                ("push", vp * 8), ("push_rows", ci), ("ld_push", ci)]


class ChainProblem(C.Structure):
    """LinksChainProblem: one GEMM of a chain launch + its level and producer indices [A operand, add0, add1]."""
    _fields_ = [("g", GemmProblem), ("level", ci), ("dep", ci * 3), ("dep_all_rows", ci)]


class ChainPlan(C.Structure):
    """LinksGemmChainPlan (filled by links_gemm_chain_build, consumed by links_gemm_chain_run)."""
    _fields_ = [("ws", vp), ("probs", vp), ("sched", vp), ("sched_cnt", vp), ("counters", vp),
                ("grid", ci), ("n_counters", ci), ("sched_ld", ci), ("n_problems", ci), ("total_tiles", ci),
                ("sim_units", cf), ("ideal_units", cf)]


MAX_CHAIN_PROBLEMS = 512


class AdamZeroLayer(C.Structure):
    """LinksAdamZeroLayer (links_adam_zero)."""
    _fields_ = [("master_off", C.c_ulonglong), ("stage_off", C.c_ulonglong), ("shadow", vp * 8)]


class ColsumItem(C.Structure):
    _fields_ = [("G", vp), ("out", vp), ("ldg", ci), ("M", ci), ("N", ci), ("accumulate", ci)]


MAX_COLSUM_ITEMS = 96


class CastItem(C.Structure):
    _fields_ = [("W", vp), ("Wb", vp), ("N", ci), ("K", ci), ("ldw", ci), ("pad_", ci)]


MAX_CAST_ITEMS = 128


class GeomMaps(C.Structure):
    _fields_ = [("V", ci), ("n_joints", ci * 2), ("src_net", (ci * 17) * 2), ("col", ci * 17),
                ("part_net", (ci * 17) * 2), ("part_idx", (ci * 17) * 2), ("bone_rel", cf * 16),
                ("depth", cf), ("w_likeli", cf), ("w_2d", cf), ("w_3d", cf), ("w_vel", cf), ("w_bl", cf)]


PP = C.POINTER(vp)
# name -> (restype, argtypes WITHOUT the trailing stream argument)
SIGNATURES = {
    "links_pack_rows": (ci, [vp, ci, ci, vp, ci, ci, vp, vp, ci, ci]),
    "links_colsum_bf16": (ci, [vp, ci, ci, ci, vp, ci]),
    "links_colsum_bf16_batched": (ci, [C.POINTER(ColsumItem), ci]),
    "links_cast_weight": (ci, [vp, ci, ci, vp, ci, vp, ci]),
    "links_cast_weight_batched": (ci, [C.POINTER(CastItem), ci]),
    "links_adam_step": (ci, [vp, vp, vp, vp, sz, cf, cf, cf, cf, cf, ci, vp, cf, vp]),
    "links_adam_step_g16": (ci, [vp, vp, vp, vp, sz, cf, cf, cf, cf, cf, ci, vp, cf, vp]),
    "links_grad_compress_bf16": (ci, [vp, vp, sz]),
    "links_small_matvec": (ci, [vp, vp, ci, ci, vp]),
    "links_normalize_head": (ci, [vp, ci, ci, ci, cf, vp, vp]),
    "links_peer_barrier": (ci, [vp, ci, ci, ci]),
    "links_adam_zero": (ci, [vp, vp, vp, vp, sz, vp, ci, ci, ci, ci, ci, vp]),
    "links_adam_prepare": (ci, [vp, vp, cf, cf, cf, cf, cf, cf, vp]),
    "links_elev_stats": (ci, [vp, vp, ci, vp]),
    "links_geom_forward": (ci, [C.POINTER(GeomMaps)] + [vp] * 8 + [ci] + [vp] * 4),
    "links_geom_loss": (ci, [C.POINTER(GeomMaps)] + [vp] * 10 + [ci] + [vp] * 5 + [ci, ci]),
    "links_geom_backward": (ci, [C.POINTER(GeomMaps)] + [vp] * 14 + [ci] + [vp] * 4 + [ci, ci] + [vp] * 3),
    "links_geom_backward_angles": (ci, [vp] * 6 + [ci] + [vp] * 4 + [ci, ci, ci]),
    "links_elev_sums": (ci, [vp, vp, ci, vp]),
    "links_elev_finalize": (ci, [vp, ci, vp]),
    "links_flow_pack": (ci, [ci, ci] + [PP] * 8 + [vp]),
    "links_flow_apply": (ci, [vp, ci, ci, vp, ci, ci, vp, vp]),
    "links_flow_nll_fwdbwd": (ci, [vp, ci, ci, vp, ci, cf, vp, vp, vp]),
    "links_flow_sample": (ci, [vp, ci, vp, vp, ci, vp]),
    "links_flow_vjp": (ci, [vp, ci, ci, vp, ci, vp, vp, vp]),
    "links_flow_nll_train": (ci, [vp, ci, ci, vp, ci, cf, vp, vp, vp, vp, vp, vp, vp]),
    "links_flow_vjp_train": (ci, [vp, ci, ci, vp, ci, vp, vp, vp, vp, vp, vp, vp, vp]),
    "links_mpjpe": (ci, [vp, vp, ci, ci, ci, ci, vp, vp, vp, vp]),
    "links_threshold_counts": (ci, [vp, sz, vp, ci, ci, vp]),
    "links_pmpjpe": (ci, [vp, vp, ci, ci, ci, vp, vp, vp]),
    "links_eval_lift_score": (ci, [vp, vp, ci, vp, ci, cf, vp]),
    "links_occ_lift": (ci, [vp, vp, vp, ci, cf, vp]),
    "links_occ_rotate_y": (ci, [vp, vp, ci, vp]),
    "links_occ_mse": (ci, [vp, ci, vp, vp, ci, ci, cf, vp, vp, vp, ci, ci]),
}
# entry points without a stream argument
PLAIN = {
    "links_abi_version": (ci, []),
    "links_device_ok": (ci, []),
    "links_flow_packed_floats": (sz, [ci, ci]),
    "links_flow_stash_floats": (sz, [ci, ci, ci]),
    "links_flow_set_simt_only": (ci, [ci]),
    "links_gemm_set_max_ctas": (ci, [ci]),
    "links_gemm_launch_count": (sz, []),
    "links_launch_count": (sz, []),
    "links_gemm_chain_ws_bytes": (sz, [C.POINTER(ChainProblem), ci]),
}
GEMM = {"links_gemm_grouped": (ci, [C.POINTER(GemmProblem), ci, vp]),
        "links_gemm_chain_build": (ci, [C.POINTER(ChainProblem), ci, vp, sz, C.POINTER(ChainPlan), vp]),
        "links_gemm_chain_run": (ci, [C.POINTER(ChainPlan), vp])}

ALL_SYMBOLS = sorted(list(SIGNATURES) + list(PLAIN) + list(GEMM))

_lib = None


class LinksError(RuntimeError):
    pass


def load_library(path=LIB_PATH):
    """dlopen + declare prototypes (no device needed)."""
    if not os.path.exists(path):
        raise LinksError("CUDA library %s not built: run `python build.py` (there is no CPU fallback)" % path)
    L = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, list(args) + [vp]
    for name, (res, args) in list(PLAIN.items()) + list(GEMM.items()):
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, list(args)
    return L


def lib():
    """The loaded library, verified against a Blackwell device."""
    global _lib
    if _lib is None:
        L = load_library()
        import torch
        if not torch.cuda.is_available():
            raise LinksError("links_b200 needs a CUDA device (sm_100a); no CPU fallback exists")
        if not L.links_device_ok():
            raise LinksError("links_b200 is built for sm_100a (B200) only; current device is %s"
                             % torch.cuda.get_device_name())
        _lib = L
    return _lib


_ERR = {-1: "null pointer / bad size", -2: "pointer or leading dimension not 16-byte aligned",
        -3: "value outside the supported range", -4: "cuTensorMapEncodeTiled unavailable or failed"}


def check(rc, what):
    if rc == 0:
        return
    if rc < 0:
        raise ValueError("%s: %s (LINKS_E %d)" % (what, _ERR.get(rc, "argument error"), rc))
    raise LinksError("%s: CUDA error %d" % (what, rc))


class _LaunchCounter:
    """Kernel launches issued through the C ABI between construction and stop() (bench.py `gpu_launches`): counted
    inside the library (links_launch_count / links_gemm_launch_count), so cached plans and captured function pointers
    cannot hide launches from it."""

    def __init__(self, L):
        self.L = L
        self.k0, self.g0 = L.links_launch_count(), L.links_gemm_launch_count()
        self.gemm = 0

    def stop(self):
        self.gemm = self.L.links_gemm_launch_count() - self.g0
        return self.L.links_launch_count() - self.k0


def install_launch_counter():
    return _LaunchCounter(lib())
