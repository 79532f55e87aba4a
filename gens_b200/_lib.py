"""ctypes binding of libgens_b200.so -- the only door from Python into the CUDA kernels.

There is no CPU fallback: if the shared library is missing or a call fails this raises.
torch is used only for device memory and the current stream.
"""
from __future__ import annotations

import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libgens_b200.so")
ABI_VERSION = 3
DIV_TRUE, DIV_RECIP = 0, 1

_lib = None

_vp, _i, _ll, _f = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float

MAX_PEERS = 8


class VolumeScale(ctypes.Structure):
    """Mirror of gens_volume_scale_t (include/gens_b200.h)."""
    _fields_ = [
        ("feat_padded", _vp), ("H", _i), ("W", _i), ("D", _i), ("a0", _i), ("a1", _i), ("a_base", _i),
        ("channel_stride", _ll), ("k_row_scale", _f), ("grid", _vp), ("volume", _vp), ("mask_volume", _vp),
        ("n_peers", _i), ("peer_volume", _vp * MAX_PEERS), ("peer_mask", _vp * MAX_PEERS), ("self_peer", _i),
        ("cam_slot", _i),
    ]


MAX_SCALES = 8


class Pyramid(ctypes.Structure):
    """Mirror of gens_pyramid_t."""
    _fields_ = [("vol", _vp * MAX_SCALES), ("dim", _i * MAX_SCALES), ("n_scales", _i)]


def make_pyramid(tensors, dims):
    p = Pyramid()
    if len(tensors) > MAX_SCALES:
        raise RuntimeError(f"gens_b200 supports at most {MAX_SCALES} scales, got {len(tensors)}")
    for i, (t, d) in enumerate(zip(tensors, dims)):
        p.vol[i] = t.data_ptr() if t is not None else None
        p.dim[i] = int(d)
    p.n_scales = len(tensors)
    return p


class ImagePyramid(ctypes.Structure):
    """Mirror of gens_image_pyramid_t."""
    _fields_ = [("map", _vp * MAX_SCALES), ("h", _i * MAX_SCALES), ("w", _i * MAX_SCALES), ("n_scales", _i)]


def make_image_pyramid(tensors, sizes):
    p = ImagePyramid()
    for i, (t, (h, w)) in enumerate(zip(tensors, sizes)):
        p.map[i] = t.data_ptr() if t is not None else None
        p.h[i], p.w[i] = int(h), int(w)
    p.n_scales = len(tensors)
    return p


class CompositeArgs(ctypes.Structure):
    """Mirror of gens_composite_args_t."""
    _fields_ = ([("n_rays", _i), ("n_samples", _i), ("n_src", _i), ("cos_anneal_ratio", _f), ("sample_dist", _f)]
                + [(k, _vp) for k in (
                    "rays_o", "rays_d", "z_vals", "pts", "sdf_raw", "grad_raw", "smooth_raw", "colour_raw",
                    "voxel_mask", "evaluated", "mask_views", "inv_s", "z_max", "rot",
                    "weights_out", "weight_sum_out", "weight_max_out", "depth_out", "color_out", "normal_out",
                    "inside_out", "valid_out", "sdf_out", "gradients_out", "mid_inside_out", "sdf_depth_out",
                    "pts_sdf0_out", "ge_num_out", "ge_den_out", "smooth_norm_out")])


_PP = ctypes.POINTER(Pyramid)
_IP = ctypes.POINTER(ImagePyramid)

_SIGNATURES = {
    "gens_abi_version": ([], _i),
    "gens_error_string": ([_i], ctypes.c_char_p),
    "gens_pack_feature_maps": ([_vp, _vp, _i, _i, _i, _vp], _i),
    "gens_pack_feature_maps_multi": ([_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _i, _vp], _i),
    "gens_invert_poses": ([_vp, _i, _vp, _vp], _i),
    "gens_unpack_feature_grads": ([_vp, _vp, _i, _i, _i, _vp], _i),
    "gens_stage_cameras": ([_vp, _vp, _i, _vp, _i, _vp, _vp], _i),
    "gens_volume_agg_fwd_multi": ([ctypes.POINTER(VolumeScale), _i, _i, _vp, _vp, _i, _i, _vp], _i),
    "gens_volume_build": ([_vp, _vp, _vp, _vp, ctypes.POINTER(VolumeScale), _i, _i, _vp, _vp, _vp, _i, _i, _vp], _i),
    "gens_volume_agg_fwd": ([_vp, _i, _i, _i, _vp, _vp, _f, _vp, _i, _i, _i, _i, _ll, _i, _i, _vp, _vp, _vp], _i),
    "gens_unpack_slabs": ([_vp, _i, _ll, _ll, _i, _vp, _vp, _vp], _i),
    "gens_volume_project_debug": ([_i, _i, _i, _vp, _vp, _f, _vp, _i, _i, _vp, _vp, _vp, _vp], _i),
    "gens_volume_agg_bwd": ([_vp, _i, _i, _i, _vp, _vp, _f, _vp, _i, _i, _i, _i, _ll, _i, _vp, _vp, _vp], _i),
    "gens_pack_volume": ([_vp, _vp, _i, _vp], _i),
    "gens_unpack_volume": ([_vp, _vp, _i, _vp], _i),
    "gens_mask_nearest": ([_vp, _ll, _PP, _i, _vp, _vp, _vp], _i),
    "gens_trilinear_fwd": ([_vp, _ll, _PP, _vp, _vp], _i),
    "gens_trilinear_bwd": ([_vp, _ll, _PP, _vp, _vp, _PP, _vp], _i),
    "gens_trilinear_bwd2": ([_vp, _ll, _PP, _vp, _vp, _vp, _vp, _PP, _vp], _i),
    "gens_trilinear_fwd_jvp": ([_vp, _ll, _PP, _vp, _vp, _vp, _vp], _i),
    "gens_trilinear_vjp2": ([_vp, _ll, _PP, _vp, _vp, _vp, _vp, _vp, _vp], _i),
    "gens_sdf_encode": ([_vp, _vp, _vp, _ll, _f, _vp, _i, _i, _i, _vp, _vp, _vp], _i),
    "gens_sdf_act_fwd": ([_vp, _vp, _i, _vp, _ll, _i, _f, _f, _vp, _i, _vp, _vp, _vp], _i),
    "gens_copy_scaled": ([_vp, _i, _ll, _f, _vp, _i, _i, _vp], _i),
    "gens_sdf_act_bwd": ([_vp, _i, _f, _vp, _vp, _ll, _i, _vp, _i, _vp], _i),
    "gens_sdf_decode": ([_vp, _vp, _vp, _vp, _vp, _ll, _f, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp], _i),
    "gens_sdf_mlp_value_tc": ([_vp, _vp, _ll, _vp, _vp, _i, _vp, _i, _f, _i, _vp, _vp], _i),
    "gens_sdf_mlp_jvp_tc": ([_vp, _vp, _ll, _vp, _vp, _i, _vp, _i, _f, _i, _vp, _vp, _vp], _i),
    "gens_sdf_mlp_rev_tc": ([_vp, _ll, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp], _i),
    "gens_pack_nhwc4": ([_vp, _vp, _i, _i, _i, _i, _vp], _i),
    "gens_unpack_nhwc4": ([_vp, _vp, _i, _i, _i, _i, _vp], _i),
    "gens_lookup_feature_fwd": ([_vp, _ll, _i, _vp, _vp, _vp, _vp, _IP, _vp, _i, _vp, _vp, _vp, _vp], _i),
    "gens_lookup_feature_bwd": ([_vp, _ll, _i, _vp, _vp, _IP, _i, _vp, _IP, _vp], _i),
    "gens_upsample_rays": ([_vp, _vp, _vp, _vp, _i, _i, _PP, _i, _f, _i, _vp, _vp], _i),
    "gens_merge_samples": ([_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp], _i),
    "gens_composite_rays": ([ctypes.POINTER(CompositeArgs), _vp], _i),
    "gens_blend_weight_floats": ([], _i),
    "gens_blend_colour": ([_vp, _vp, _vp, _ll, _i, _vp, _vp, _vp], _i),
    "gens_patch_warp": ([_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp], _i),
    "gens_lncc_fwd": ([_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp], _i),
    "gens_lncc_bwd": ([_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp], _i),
    "gens_mc_classify": ([_vp, _i, _i, _i, _f, _vp, _vp, _vp, _vp], _i),
    "gens_mc_vertices": ([_vp, _i, _i, _i, _f, _vp, _vp, _vp, _ll, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                          _vp, _vp], _i),
    "gens_mc_triangles": ([_vp, _i, _i, _i, _f, _vp, _vp, _ll, _vp, _vp, _ll, _vp, _vp, _vp, _i, ctypes.c_char_p, _ll,
                           _vp, _vp], _i),
    "gens_conv3d_k3": ([_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp], _i),
    "gens_conv3d_k3s2": ([_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp], _i),
    "gens_deconv3d_k3s2": ([_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp], _i),
    "gens_instnorm_relu": ([_vp, _vp, _i, _ll, ctypes.c_double, _f, _vp, _vp], _i),
    "gens_tv_reduce": ([_PP, _PP, _i, _i, _vp, _vp], _i),
    "gens_debug_set_variant": ([_i], _i),
    "gens_debug_set_tc_terms": ([_i], _i),
    "gens_debug_blend_const": ([_i], _i),
    "gens_debug_conv_td8": ([_i], _i),
    "gens_debug_tc_profile": ([_vp, _i], _i),
    "gens_tf32_mma_peak": ([_i, _vp, _vp], _i),
    "gens_selftest_division": ([_i, ctypes.c_ulonglong, _vp, _vp], _i),
}


def exported_symbols():
    """Names every build of the library must export (mirrors include/gens_b200.h)."""
    return list(_SIGNATURES)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m gens_b200.build` "
                "(gens_b200 has no CPU or PyTorch fallback)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so is stale
            fn.argtypes = argtypes
            fn.restype = restype
        got = handle.gens_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError(f"libgens_b200.so ABI {got} != expected {ABI_VERSION}; rebuild")
        _lib = handle
    return _lib


LAUNCHES = 0  # C-ABI calls made so far (each enqueues at least one of our kernels); bench.py reads it


def check(code: int, what: str):
    global LAUNCHES
    LAUNCHES += 1
    if code != 0:
        msg = lib().gens_error_string(code)
        raise RuntimeError(f"{what} failed ({code}): {msg.decode() if msg else '?'}")


def ptr(t: torch.Tensor):
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    """Every tensor on the CURRENT CUDA device.  The C entries launch on the current device with the stream handed to
    them: a tensor living elsewhere would otherwise surface as an invalid-resource-handle error from the launch (one
    process drives one GPU here -- call torch.cuda.set_device / use `with torch.cuda.device(...)` first)."""
    cur = None
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("gens_b200 kernels need CUDA tensors (no CPU fallback); got a "
                               f"{t.device} tensor")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise RuntimeError(f"gens_b200: tensor on {t.device} but the current CUDA device is cuda:{cur}; "
                               "make the tensors' device current (torch.cuda.set_device) before calling")


def f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def invert_poses(t: torch.Tensor) -> torch.Tensor:
    """inverse of (n,4,4) fp32 CUDA matrices in ONE launch, bit-identical to torch.inverse on CUDA
    (gens_invert_poses); the reference's torch.inverse(c2ws) costs 11 library launches."""
    require_cuda(t)
    if t.dim() != 3 or t.shape[1:] != (4, 4):
        raise RuntimeError(f"invert_poses expects (n,4,4), got {tuple(t.shape)}")
    src = f32c(t)
    out = torch.empty_like(src)
    check(lib().gens_invert_poses(ptr(src), src.shape[0], ptr(out), stream_ptr(t.device)), "gens_invert_poses")
    return out


def inverse(t: torch.Tensor) -> torch.Tensor:
    """torch.inverse without its host-synchronising singularity check: same LU kernels (bit-identical
    result for invertible input), but the caller's stream keeps running."""
    return torch.linalg.inv_ex(t, check_errors=False)[0]
