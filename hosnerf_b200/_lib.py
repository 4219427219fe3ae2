"""ctypes binding of libhosnerf_b200.so (the C ABI declared in include/hosnerf_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call
returns a non-zero status, a RuntimeError is raised.  Build the library with
``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C hosnerf_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhosnerf_b200.so")

c_f = C.c_void_p      # device pointers travel as integers
c_i = C.c_int
c_l = C.c_int64
c_fl = C.c_float
c_hp = C.POINTER(C.c_float)   # HOST float pointer (small parameter vectors)


class MlpLayer(C.Structure):
    _fields_ = [("out_dim", c_i), ("in_h", c_i), ("in_x", c_i), ("x_first", c_i),
                ("relu", c_i), ("rowbias", c_i), ("head", c_i)]


class MlpHead(C.Structure):
    _fields_ = [("out_dim", c_i), ("post", c_i), ("shift", c_fl), ("out_slot", c_i)]


BKG_MAX_LEVELS = 4


class BkgLevel(C.Structure):      # hos_bkg_level
    _fields_ = [("mlp", C.c_void_p), ("n_samples", c_i), ("dilate", c_i), ("dilation", c_fl), ("u_base", c_f), ("jitter", c_f), ("jitter_cols", c_i),
                ("max_jitter", c_fl), ("view_W", c_f), ("view_b", c_f), ("view_dim", c_i)]


class BkgConfig(C.Structure):     # hos_bkg_config
    _fields_ = [("n_levels", c_i), ("levels", BkgLevel * BKG_MAX_LEVELS), ("s_near", c_fl), ("s_far", c_fl),
                ("dom_lo", c_fl), ("dom_hi", c_fl), ("anneal", c_fl), ("resample_padding", c_fl),
                ("opaque_background", c_i), ("bg", c_fl), ("deg_view", c_i), ("basis_host", c_hp),
                ("mlp_events", C.POINTER(C.c_void_p))]


class HumanConfig(C.Structure):       # hos_human_config
    _fields_ = [("nr_mlp", C.c_void_p), ("cnl_mlp", C.c_void_p), ("n_samples", c_i), ("t_lin", c_f), ("jitter", c_f),
                ("R", c_f), ("T", c_f), ("vol", c_f), ("bones", c_i), ("grid", c_i), ("bbox_min_host", c_hp), ("bbox_scale_host", c_hp),
                ("nr_freqs", c_i), ("hann_w", c_f), ("hann_w_host", c_hp), ("cnl_freqs", c_i), ("stage2", c_i), ("bgcolor_host", c_hp)]


class GemmTmaDesc(C.Structure):      # hos_gemm_tma_desc
    _fields_ = [("mode", c_i), ("rows", c_l),
                ("a0_hi", c_f), ("a0_lo", c_f), ("k0", c_i), ("lda0", c_i),
                ("a1_hi", c_f), ("a1_lo", c_f), ("k1", c_i), ("lda1", c_i),
                ("w0_hi", c_f), ("w0_lo", c_f), ("ldw0", c_i),
                ("w1_hi", c_f), ("w1_lo", c_f), ("ldw1", c_i),
                ("n", c_i), ("bias", c_f), ("rowbias", c_f), ("rowbias_div", c_i), ("relu", c_i), ("mask", c_f), ("ld_mask", c_i),
                ("y_hi", c_f), ("y_lo", c_f), ("ldy", c_i), ("y_f32", c_f), ("ldy32", c_i),
                ("hn", c_i), ("head_w", c_f), ("head_b", c_f), ("head_post", c_i), ("head_shift", c_fl), ("head_out", c_f)]


# name -> (restype, argtypes); mirrors include/hosnerf_b200.h one to one
SIGNATURES = {
    "hos_last_error": (C.c_char_p, []),
    "hos_version": (c_i, []),
    "hos_device_check": (c_i, [c_i]),
    "hos_max_dilate": (c_i, [c_f, c_f, c_i, c_i, c_fl, c_fl, c_fl, c_f, c_f, c_f]),
    "hos_sample_intervals": (c_i, [c_f, c_f, c_f, c_f, c_i, c_fl, c_i, c_i, c_i, c_fl, c_fl, c_f, c_f, c_f, c_f]),
    "hos_invert_cdf": (c_i, [c_f, c_f, c_f, c_f, c_i, c_fl, c_i, c_i, c_i, c_fl, c_fl, c_f, c_f, c_f, c_f]),
    "hos_resample_level": (c_i, [c_f, c_f, c_i, c_i, c_i, c_fl, c_fl, c_fl, c_f, c_f, c_i, c_fl, c_i, c_fl,
                                 c_fl, c_fl, c_fl, c_f, c_f, c_f]),
    "hos_human_samples": (c_i, [c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_f, c_f, c_f]),
    "hos_ipe_features": (c_i, [c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_i, c_f, c_f, c_f]),
    "hos_ipe_from_gaussians": (c_i, [c_f, c_f, c_f, c_l, c_i, c_i, c_i, c_f, c_i, c_f]),
    "hos_pos_enc": (c_i, [c_f, c_i, c_i, c_i, c_i, c_f, c_f]),
    "hos_fourier_embed": (c_i, [c_f, c_l, c_i, c_i, c_f, c_f, c_i, c_i, c_f]),
    "hos_lbs_warp": (c_i, [c_f, c_f, c_f, c_f, c_hp, c_hp, c_l, c_i, c_i, c_f, c_f, c_f]),
    "hos_lbs_forward": (c_i, [c_f, c_f, c_f, c_f, c_hp, c_hp, c_l, c_i, c_i, c_f, c_f, c_f]),
    "hos_lbs_warp_backward": (c_i, [c_f, c_f, c_f, c_f, c_hp, c_hp, c_l, c_i, c_i, c_f, c_f, c_f, c_f, c_f, c_f]),
    "hos_lbs_forward_backward": (c_i, [c_f, c_f, c_f, c_f, c_hp, c_hp, c_l, c_i, c_i, c_f, c_f, c_f, c_f, c_f, c_f]),
    "hos_linear_f32": (c_i, [c_f, c_i, c_i, c_f, c_i, c_i, c_f, c_f, c_l, c_i, c_i, c_f, c_i, c_f]),
    "hos_linear_f32_ex": (c_i, [c_f, c_i, c_i, c_f, c_i, c_i, c_i, c_f, c_f, c_l, c_i, c_i, c_f, c_i, c_f]),
    "hos_head_f32": (c_i, [c_f, c_i, c_i, c_f, c_f, c_l, c_i, c_i, c_fl, c_f, c_f, c_i, c_f]),
    "hos_mlp_create": (C.c_void_p, [c_i, c_i, C.POINTER(MlpLayer), c_i, C.POINTER(MlpHead)]),
    "hos_mlp_destroy": (None, [C.c_void_p]),
    "hos_mlp_set_layer": (c_i, [C.c_void_p, c_i, c_f, c_f, c_f]),
    "hos_mlp_set_bias": (c_i, [C.c_void_p, c_i, c_f, c_f]),
    "hos_mlp_set_head": (c_i, [C.c_void_p, c_i, c_f, c_f, c_f]),
    "hos_mlp_in_kblocks": (c_i, [C.c_void_p]),
    "hos_mlp_forward": (c_i, [C.c_void_p, c_f, c_l, c_f, c_i, c_f, c_f, c_f, c_f]),
    "hos_mlp_set_ipe_input": (c_i, [C.c_void_p, c_i]),
    "hos_mlp_forward_ipe": (c_i, [C.c_void_p, c_f, c_f, c_f, c_f, c_hp, c_i, c_i, c_f, c_i, c_f, c_f, c_f]),
    "hos_mlp_fourier_supported": (c_i, [C.c_void_p, c_l]),
    "hos_mlp_forward_fourier": (c_i, [C.c_void_p, c_f, c_l, c_i, c_i, c_hp, c_f, c_f, c_f, c_f]),
    "hos_mlp_debug_timeline": (c_i, [C.c_void_p, c_f]),
    "hos_mlp_set_variant": (c_i, [C.c_void_p, c_i]),
    "hos_pack_rows_f16": (c_i, [c_f, c_l, c_i, c_i, c_f, c_f]),
    "hos_gemm_create": (C.c_void_p, [c_i, c_i, c_i, c_i, c_i]),
    "hos_ipe_features_fast": (c_i, [c_f, c_f, c_f, c_f, c_hp, c_i, c_i, c_f, c_f]),
    "hos_gemm_destroy": (None, [C.c_void_p]),
    "hos_gemm_set_cluster": (c_i, [C.c_void_p, c_i]),
    "hos_gemm_set_weight": (c_i, [C.c_void_p, c_f, c_f, c_f]),
    "hos_gemm_set_head": (c_i, [C.c_void_p, c_i, c_f, c_f, c_f]),
    "hos_gemm_forward": (c_i, [C.c_void_p, c_f, c_f, c_l, c_i, c_f, c_f, c_i, c_fl, c_f]),
    "hos_gemm_tma": (c_i, [C.POINTER(GemmTmaDesc), c_f]),
    "hos_wgrad_tma": (c_i, [c_f, c_i, c_i, c_f, c_i, c_i, c_l, c_f, c_i, c_i, c_f, c_f]),
    "hos_colsum_f16": (c_i, [c_f, c_l, c_i, c_i, c_f, c_i, c_f, c_i, c_f]),
    "hos_head_dgrad": (c_i, [c_f, c_i, c_f, c_i, c_f, c_i, c_f, c_i, c_l, c_i, c_f, c_i, c_f]),
    "hos_rays_from_krt": (c_i, [c_i, c_i, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), c_i, c_i,
                                c_f, c_f, c_f, c_f, c_f]),
    "hos_rays_intersect_bbox": (c_i, [C.POINTER(C.c_double), C.POINTER(C.c_double), c_f, c_f, c_l, c_i, c_f, c_f, c_f, c_f]),
    "hos_composite_mip360": (c_i, [c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_fl, c_f, c_f, c_f]),
    "hos_composite_mip360_backward": (c_i, [c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_fl, c_f, c_f, c_f]),
    "hos_composite_nerf": (c_i, [c_f, c_f, c_f, c_f, c_hp, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_f]),
    "hos_composite_s3": (c_i, [c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_hp, c_f, c_f, c_i, c_i, c_i, c_fl,
                               c_f, c_f, c_f, c_f, c_f]),
    "hos_render_bkg_workspace": (c_i, [C.POINTER(BkgConfig), c_i, C.POINTER(C.c_size_t)]),
    "hos_render_bkg": (c_i, [C.POINTER(BkgConfig), c_f, c_f, c_f, c_f, c_i, c_f, C.c_size_t, c_f, c_f, c_f, c_f]),
    "hos_render_human_workspace": (c_i, [C.POINTER(HumanConfig), c_i, C.POINTER(C.c_size_t)]),
    "hos_render_human": (c_i, [C.POINTER(HumanConfig), c_f, c_f, c_f, c_f, c_i, c_f, C.c_size_t, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_f]),
    "hos_lossfun_distortion": (c_i, [c_f, c_f, c_i, c_i, c_f, c_f]),
    "hos_lossfun_outer": (c_i, [c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_f, c_f, c_f]),
    "hos_lossfun_distortion_backward": (c_i, [c_f, c_f, c_f, c_fl, c_i, c_i, c_f, c_f]),
    "hos_lossfun_outer_backward": (c_i, [c_f, c_f, c_f, c_f, c_fl, c_i, c_i, c_i, c_f, c_f]),
    "hos_reduce_scaled": (c_i, [c_f, c_f, c_l, C.c_double, c_f, c_f]),
}

_lib = None
LAUNCHES = 0          # kernels of this library launched so far (bench.py reports it per timed region)
_KERNELS_PER_CALL = {"hos_composite_s3": 2, "hos_mlp_set_layer": 2, "hos_mlp_set_bias": 1, "hos_mlp_set_head": 0,
                     "hos_mlp_set_ipe_input": 0, "hos_mlp_set_variant": 0, "hos_gemm_set_head": 0,
                     "hos_render_bkg_workspace": 0, "hos_render_bkg": 0, "hos_render_human_workspace": 0, "hos_render_human": 7}      # hos_render_bkg: counted by its caller


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is not built and hosnerf_b200 has no CPU "
            "fallback.  Run `python -c 'import __graft_entry__ as g; g.build()'` in the repo root.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = ""):
    if status != 0:
        msg = load().hos_last_error().decode(errors="replace")
        raise RuntimeError(f"libhosnerf_b200 {what} failed (status {status}): {msg}")


def call(name: str, *args):
    global LAUNCHES
    check(getattr(load(), name)(*args), name)
    LAUNCHES += _KERNELS_PER_CALL.get(name, 1)


def call_unless_empty(n, name: str, *args):
    """torch gives empty tensors a NULL data pointer; an empty batch is a no-op by contract."""
    if n != 0:
        call(name, *args)
    # (LAUNCHES is only incremented by real launches)
