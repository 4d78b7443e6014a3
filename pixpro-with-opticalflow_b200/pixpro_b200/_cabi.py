"""ctypes binding of libpixpro_b200.so — the C ABI declared in include/pixpro_b200.h.

There is NO fallback: if the library is missing or a call fails, an exception is raised.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpixpro_b200.so")
CSRC = os.path.join(os.path.dirname(_HERE), "csrc")

_vp = ctypes.c_void_p
_i = ctypes.c_int
_l = ctypes.c_int64
_d = ctypes.c_double

# name -> (restype, argtypes); mirrors include/pixpro_b200.h one to one
SIGNATURES = {
    "pp_abi_version": (_i, []),
    "pp_last_error": (ctypes.c_char_p, []),
    "pp_launch_count": (_l, []),
    "pp_fb_redo_count": (_l, [_i]),
    "pp_chain_slow_count": (_l, [_i]),
    "pp_profile_enable": (_i, [_i]),
    "pp_profile_num_kernels": (_i, []),
    "pp_profile_get": (_i, [_i, ctypes.c_char_p, _i, ctypes.POINTER(_l), ctypes.POINTER(_d)]),
    "pp_mt_chunk_elems": (_i, []),
    "pp_ema_update": (_i, [_vp, _vp, _i, _d, _d, _vp]),
    "pp_lars_workspace": (_l, [_i, _i]),
    "pp_lars_sgd_step": (_i, [_vp, _i, _vp, _vp, _i, _d, _d, _vp, _vp]),
    "pp_upflow8": (_i, [_vp, _l, _i, _i, _vp, _vp]),
    "pp_normalize": (_i, [_vp, _l, _i, _i, _i, _i, _vp, _vp]),
    "pp_concat_flow": (_i, [_vp, _i, _l, _i, _i, _l, _l, _i, _i, _vp, _vp]),
    "pp_fb_consistency": (_i, [_vp, _vp, _l, _i, _i, _d, _d, _i, _i, _vp, _vp, _vp, _vp]),
    "pp_fb_masks": (_i, [_vp, _vp, _l, _i, _i, _d, _d, _i, _i, _vp, _vp, _vp]),
    "pp_flow_stage_workspace": (_l, [_l, _i, _i, _i, _i]),
    "pp_flow_stage": (_i, [_vp, _vp, _l, _i, _i, _i, _i, _i, _d, _d, _i, _i, _vp, _vp, _vp, _vp, _vp, _l, _vp]),
    "pp_calc_mask_ratio": (_i, [_vp, _l, _i, _i, _vp, _vp]),
    "pp_add_optical_flow": (_i, [_vp, _l, _i, _i, _vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "pp_regression_loss_workspace": (_l, [_l, _i, _i]),
    "pp_regression_loss": (_i, [_vp, _vp, _l, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _i, _i, _d, _i,
                                _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pp_regression_loss_pair": (_i, [_vp, _vp, _l, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _i, _i, _d, _i,
                                     _vp, _vp, _vp, _vp, _vp, _vp]),
    "pp_sparse_corr": (_i, [_vp, _vp, _l, _i, _i, _i, _i, _i, _d, _d, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "pp_regression_loss_warped": (_i, [_vp, _vp, _l, _i, _i, _vp, _vp, _vp, _i, _i, _d, _i,
                                       _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pp_regression_loss_pair_warped": (_i, [_vp, _vp, _l, _i, _i, _vp, _vp, _vp, _i, _i, _d, _i,
                                            _vp, _vp, _vp, _vp, _vp, _vp]),
    "pp_ppm_saved_bytes": (_l, [_l, _i, _i]),
    "pp_ppm_fwd": (_i, [_vp, _vp, _l, _i, _i, _d, _d, _i, _vp, _vp, _vp]),
    "pp_ppm_bwd_workspace": (_l, [_l, _i, _i]),
    "pp_ppm_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _l, _i, _i, _d, _d, _i, _vp, _vp, _vp, _vp]),
    "pp_conv1x1_fwd_workspace": (_l, [_l, _i, _i, _i]),
    "pp_conv1x1_fwd": (_i, [_vp, _vp, _vp, _l, _i, _i, _i, _vp, _vp, _vp]),
    "pp_conv1x1_bwd_workspace": (_l, [_l, _i, _i, _i]),
    "pp_conv1x1_bwd": (_i, [_vp, _vp, _vp, _l, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pp_tc_gemm_nt": (_i, [_vp, _vp, _vp, _l, _i, _i, _i, _vp]),
    "pp_tc_gemm_nt_workspace": (_l, [_l, _i, _i, _i]),
    "pp_tc_gemm_nt_ws": (_i, [_vp, _vp, _vp, _l, _i, _i, _i, _vp, _vp]),
    "pp_bn_workspace": (_l, [_i]),
    "pp_bn_stats": (_i, [_vp, _l, _i, _i, _i, _i, _vp, _vp, _vp]),
    "pp_bn_apply": (_i, [_vp, _vp, _l, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _d, _d, _vp, _vp, _vp]),
    "pp_bn_bwd_stats": (_i, [_vp, _vp, _l, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pp_bn_bwd_apply": (_i, [_vp, _vp, _vp, _l, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pp_tc_gemm_ws": (_i, [_vp, _vp, _vp, _l, _i, _i, _i, _i, _i, _vp]),
    "pp_tc_gemm_ex": (_i, [_vp, _vp, _vp, _l, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "pp_corr_volume": (_i, [_vp, _vp, _l, _i, _i, _i, _vp, _vp]),
    "pp_corr_pool": (_i, [_vp, _l, _i, _i, _vp, _vp]),
    "pp_corr_lookup": (_i, [_vp, _i, _vp, _l, _i, _i, _i, _i, _vp, _vp]),
}

_lib = None


class PixProB200Error(RuntimeError):
    pass


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into libpixpro_b200.so (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-C", CSRC] + ([] if verbose else ["-s"]))
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PixProB200Error(
                f"{LIB_PATH} is missing: build it with `make -C {CSRC}` (or __graft_entry__.build()). "
                "There is no CPU or PyTorch fallback for the pixel-pretext hot path.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().pp_last_error()
        raise PixProB200Error(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def launch_count():
    return int(lib().pp_launch_count())


def fb_redo_count(reset=False):
    """Pixels the TMA-staged FB-mask kernel recomputed from global memory (see pp_fb_redo_count)."""
    return int(lib().pp_fb_redo_count(int(bool(reset))))


def chain_slow_count(reset=False):
    """Pixel-links the fused up-sampling chain kernel evaluated on its direct (slow) path (see pp_chain_slow_count)."""
    return int(lib().pp_chain_slow_count(int(bool(reset))))


def profile_enable(on=True):
    """Bracket every kernel launch with CUDA events (tracing aid; see pp_profile_enable)."""
    check(lib().pp_profile_enable(int(bool(on))), "pp_profile_enable")


def profile_report():
    """{kernel name: (launches, total device ms)} for launches since profile_enable(True)."""
    L = lib()
    out = {}
    for i in range(L.pp_profile_num_kernels()):
        name = ctypes.create_string_buffer(128)
        n = _l(0)
        ms = _d(0.0)
        check(L.pp_profile_get(i, name, 128, ctypes.byref(n), ctypes.byref(ms)), "pp_profile_get")
        out[name.value.decode()] = (int(n.value), float(ms.value))
    return out
