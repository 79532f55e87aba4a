"""Volume / feature look-ups of the ray marcher -- drop-ins for the reference's
models/modules/projector.py (lookup_volume :217-245, lookup_feature :294-349, surface_patch_warp
:353-419) and the autograd triple of models/modules/grid_sample_cuda/cuda_gridsample.py.

All scales of a pyramid are sampled by ONE kernel launch (csrc/sampling.cu); re-laid-out copies of
the volumes (channels-last) are owned here, keyed on the tensor object and its _version so that volumes
optimised in place during fine-tuning are re-packed when they change.
"""
from __future__ import annotations

import ctypes
import weakref
from typing import Dict, List, Sequence, Tuple, Union

import torch

from . import _lib

# 1 = un-normalise nearest look-ups the way ATen's CUDA kernels do (fused multiply-subtract);
# masks are then bit-identical to the reference running on the same GPU.  0 = ATen CPU flavour.
ATEN_CUDA_FLAVOUR = 1

# id(tensor) -> (weakref to the tensor, its _version when packed, packed copy).  The weakref guards
# against id / storage reuse after the source tensor died; _version catches in-place updates.
_PACK_CACHE: Dict[int, Tuple["weakref.ref", int, torch.Tensor]] = {}
_PACK_CACHE_MAX = 32


def _as_list(volume) -> List[torch.Tensor]:
    return [volume] if isinstance(volume, torch.Tensor) else list(volume)


def packed_volume(vol: torch.Tensor) -> torch.Tensor:
    """Channels-last (D,D,D,4) copy of a (1,4,D,D,D) volume, cached per (storage, version)."""
    if vol.dim() != 5 or vol.shape[0] != 1 or vol.shape[1] != 4 or not (vol.shape[2] == vol.shape[3] == vol.shape[4]):
        raise RuntimeError(f"gens_b200 trilinear kernels need (1,4,D,D,D) volumes, got {tuple(vol.shape)}")
    key = id(vol)
    hit = _PACK_CACHE.get(key)
    if hit is not None and hit[0]() is vol and hit[1] == vol._version and hit[2].device == vol.device:
        return hit[2]
    src = _lib.f32c(vol.detach())
    d = vol.shape[2]
    out = torch.empty((d, d, d, 4), device=vol.device, dtype=torch.float32)
    _lib.check(_lib.lib().gens_pack_volume(_lib.ptr(src), _lib.ptr(out), d, _lib.stream_ptr(vol.device)),
               "gens_pack_volume")
    for k in [k for k, (ref, _, _) in _PACK_CACHE.items() if ref() is None]:
        del _PACK_CACHE[k]
    if len(_PACK_CACHE) >= _PACK_CACHE_MAX:
        _PACK_CACHE.pop(next(iter(_PACK_CACHE)))
    _PACK_CACHE[key] = (weakref.ref(vol), vol._version, out)
    return out


def clear_caches():
    _PACK_CACHE.clear()
    _TV_CACHE.clear()


def _pts(pts: torch.Tensor) -> torch.Tensor:
    _lib.require_cuda(pts)
    return _lib.f32c(pts.reshape(-1, 3))


def _unpack_grad(g_cl: torch.Tensor) -> torch.Tensor:
    d = g_cl.shape[0]
    out = torch.empty((1, 4, d, d, d), device=g_cl.device, dtype=torch.float32)
    _lib.check(_lib.lib().gens_unpack_volume(_lib.ptr(g_cl), _lib.ptr(out), d, _lib.stream_ptr(g_cl.device)),
               "gens_unpack_volume")
    return out


class _Trilinear(torch.autograd.Function):
    """feats(n, 4S) = trilinear look-up of S volumes; three differentiation levels like the reference's
    _GridSample3dForward / _GridSample3dBackward / grad2_3d (cuda_gridsample.py:71-123)."""

    @staticmethod
    def forward(ctx, pts, *vols):
        packed = [packed_volume(v) for v in vols]
        dims = [v.shape[2] for v in vols]
        n = pts.shape[0]
        out = torch.empty((n, 4 * len(vols)), device=pts.device, dtype=torch.float32)
        pyr = _lib.make_pyramid(packed, dims)
        _lib.check(_lib.lib().gens_trilinear_fwd(_lib.ptr(pts), n, pyr, _lib.ptr(out), _lib.stream_ptr(pts.device)),
                   "gens_trilinear_fwd")
        ctx.save_for_backward(pts, *vols)
        return out

    @staticmethod
    def backward(ctx, g_out):
        pts, *vols = ctx.saved_tensors
        res = _TrilinearBackward.apply(g_out, pts, ctx.needs_input_grad[0], tuple(ctx.needs_input_grad[1:]), *vols)
        return (res[0], *res[1:])


class _TrilinearBackward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g_out, pts, need_pts, need_vols, *vols):
        packed = [packed_volume(v) for v in vols]
        dims = [v.shape[2] for v in vols]
        n = pts.shape[0]
        dev = pts.device
        g_out = _lib.f32c(g_out)
        g_pts = torch.empty((n, 3), device=dev, dtype=torch.float32) if need_pts else None
        g_cl = [torch.zeros_like(p) if need else None for p, need in zip(packed, need_vols)]
        pyr = _lib.make_pyramid(packed, dims)
        gpyr = _lib.make_pyramid(g_cl, dims) if any(need_vols) else None
        _lib.check(_lib.lib().gens_trilinear_bwd(_lib.ptr(pts), n, pyr, _lib.ptr(g_out),
                                                 _lib.ptr(g_pts) if need_pts else None, gpyr,
                                                 _lib.stream_ptr(dev)), "gens_trilinear_bwd")
        ctx.save_for_backward(g_out, pts, *vols)
        ctx.need_vols = need_vols
        g_vols = [(_unpack_grad(g) if g is not None else None) for g in g_cl]
        return (g_pts, *g_vols)

    @staticmethod
    def backward(ctx, gg_pts, *gg_vols):
        g_out, pts, *vols = ctx.saved_tensors
        if any(g is not None and bool(g.any()) for g in gg_vols):
            raise RuntimeError("gens_b200: second-order gradients THROUGH the volume gradient are not implemented "
                               "(the reference passes zeros here, cuda_gridsample.py:113-114)")
        if gg_pts is None:
            return (None, None, None, None, *([None] * len(vols)))
        packed = [packed_volume(v) for v in vols]
        dims = [v.shape[2] for v in vols]
        n = pts.shape[0]
        dev = pts.device
        gg_pts = _lib.f32c(gg_pts)
        gg_out = torch.empty_like(g_out)
        g2_pts = torch.empty((n, 3), device=dev, dtype=torch.float32)
        need_vols = tuple(ctx.needs_input_grad[4:])
        g_cl = [torch.zeros_like(p) if need else None for p, need in zip(packed, need_vols)]
        pyr = _lib.make_pyramid(packed, dims)
        gpyr = _lib.make_pyramid(g_cl, dims) if any(need_vols) else None
        _lib.check(_lib.lib().gens_trilinear_bwd2(_lib.ptr(pts), n, pyr, _lib.ptr(g_out), _lib.ptr(gg_pts),
                                                  _lib.ptr(gg_out), _lib.ptr(g2_pts), gpyr, _lib.stream_ptr(dev)),
                   "gens_trilinear_bwd2")
        g_vols = [(_unpack_grad(g) if g is not None else None) for g in g_cl]
        # like the reference's grad2_3d, the results carry no grad_fn: third-order terms are dropped
        return (gg_out, g2_pts, None, None, *g_vols)


def mask_nearest(pts: torch.Tensor, masks, want_each: bool = False):
    """Nearest look-up of the mask pyramid.  Returns the (n,) bool `any` reduction every caller of the
    reference applies, or the raw (n,S) float samples when want_each."""
    masks = _as_list(masks)
    p = _pts(pts)
    n = p.shape[0]
    ms = [_lib.f32c(m) for m in masks]
    for m in ms:
        _lib.require_cuda(m)
        if m.dim() != 5 or m.shape[0] != 1 or m.shape[1] != 1:
            raise RuntimeError(f"mask volumes must be (1,1,D,D,D), got {tuple(m.shape)}")
    pyr = _lib.make_pyramid(ms, [m.shape[2] for m in ms])
    if want_each:
        each = torch.empty((n, len(ms)), device=p.device, dtype=torch.float32)
        _lib.check(_lib.lib().gens_mask_nearest(_lib.ptr(p), n, pyr, ATEN_CUDA_FLAVOUR, None, _lib.ptr(each),
                                                _lib.stream_ptr(p.device)), "gens_mask_nearest")
        return each
    any_ = torch.empty((n,), device=p.device, dtype=torch.uint8)
    _lib.check(_lib.lib().gens_mask_nearest(_lib.ptr(p), n, pyr, ATEN_CUDA_FLAVOUR, _lib.ptr(any_), None,
                                            _lib.stream_ptr(p.device)), "gens_mask_nearest")
    return any_.bool()


def lookup_volume(pts: torch.Tensor, volume: Union[torch.Tensor, Sequence[torch.Tensor]], sample_mode: str = "grad"):
    """Same contract as the reference (projector.py:217-245): (n,3) points -> (n, sum C)."""
    vols = _as_list(volume)
    p = _pts(pts)
    if sample_mode == "grad":
        return _Trilinear.apply(p, *vols)
    if sample_mode == "nearest":
        if all(v.shape[1] == 1 for v in vols):
            return mask_nearest(p, vols, want_each=True)
        raise RuntimeError("gens_b200: nearest look-up is implemented for 1-channel (mask) volumes")
    raise RuntimeError(f"gens_b200: unsupported sample_mode {sample_mode!r}")


# ---- source-view reprojection (reference projector.py:278-349): K6 -----------------------------
def _pack_nhwc4(t: torch.Tensor) -> torch.Tensor:
    """(n,c,h,w), c in {3,4} -> channels-last (n,h,w,4)."""
    t = _lib.f32c(t)
    n, c, h, w = t.shape
    out = torch.empty((n, h, w, 4), device=t.device, dtype=torch.float32)
    _lib.check(_lib.lib().gens_pack_nhwc4(_lib.ptr(t), _lib.ptr(out), n, c, h, w, _lib.stream_ptr(t.device)),
               "gens_pack_nhwc4")
    return out


class _LookupFeature(torch.autograd.Function):
    """One launch for all source views and scales; differentiable w.r.t. the feature maps only (the
    sampling grid is under no_grad in the reference, projector.py:318-334)."""

    @staticmethod
    def forward(ctx, pts, w2c_src, k_src, c2w_ref, c2w_src, imgs_src, *feats):
        n, ns = pts.shape[0], k_src.shape[0]
        dev = pts.device
        packed = [_pack_nhwc4(f[1:]) for f in feats]
        rgb = _pack_nhwc4(imgs_src)
        sizes = [(f.shape[2], f.shape[3]) for f in feats]
        pyr = _lib.make_image_pyramid(packed, sizes)
        width = 3 + 4 * len(feats)
        out = torch.empty((n, ns, width), device=dev, dtype=torch.float32)
        ray_diff = torch.empty((n, ns, 4), device=dev, dtype=torch.float32)
        mask = torch.empty((n, ns), device=dev, dtype=torch.uint8)
        _lib.check(_lib.lib().gens_lookup_feature_fwd(
            _lib.ptr(pts), n, ns, _lib.ptr(w2c_src), _lib.ptr(k_src), _lib.ptr(c2w_ref), _lib.ptr(c2w_src), pyr,
            _lib.ptr(rgb), ATEN_CUDA_FLAVOUR, _lib.ptr(out), _lib.ptr(ray_diff), _lib.ptr(mask),
            _lib.stream_ptr(dev)), "gens_lookup_feature_fwd")
        ctx.save_for_backward(pts, w2c_src, k_src, *packed)
        ctx.shapes = [tuple(f.shape) for f in feats]
        ctx.flavour = ATEN_CUDA_FLAVOUR
        mask = mask.bool()
        ctx.mark_non_differentiable(ray_diff, mask)
        return out, ray_diff, mask

    @staticmethod
    def backward(ctx, g_out, _g_rd, _g_mask):
        pts, w2c_src, k_src, *packed = ctx.saved_tensors
        n, ns = pts.shape[0], k_src.shape[0]
        dev = pts.device
        g_out = _lib.f32c(g_out)
        need = ctx.needs_input_grad[6:]
        g_cl = [torch.zeros_like(p) if nd else None for p, nd in zip(packed, need)]
        sizes = [(sh[2], sh[3]) for sh in ctx.shapes]
        if any(need):
            _lib.check(_lib.lib().gens_lookup_feature_bwd(
                _lib.ptr(pts), n, ns, _lib.ptr(w2c_src), _lib.ptr(k_src), _lib.make_image_pyramid(packed, sizes),
                ctx.flavour, _lib.ptr(g_out), _lib.make_image_pyramid(g_cl, sizes), _lib.stream_ptr(dev)),
                "gens_lookup_feature_bwd")
        grads = []
        for g, sh in zip(g_cl, ctx.shapes):
            if g is None:
                grads.append(None)
                continue
            full = torch.zeros(sh, device=dev, dtype=torch.float32)  # the reference view (index 0) gets no gradient
            src = torch.empty((sh[0] - 1,) + sh[1:], device=dev, dtype=torch.float32)
            _lib.check(_lib.lib().gens_unpack_nhwc4(_lib.ptr(g), _lib.ptr(src), sh[0] - 1, sh[1], sh[2], sh[3],
                                                    _lib.stream_ptr(dev)), "gens_unpack_nhwc4")
            full[1:] = src
            grads.append(full)
        return (None, None, None, None, None, None, *grads)


def lookup_feature(pts, imgs, intrs, c2ws, features):
    """Project points into every source view at every scale, sample RGB (scale 0) and features.
    Returns ((n,ns,3+sum c), (n,ns,4), (n,ns) bool) exactly as the reference (projector.py:294-349):
    align-corners normalisation, align_corners=False sampling, no epsilon in the perspective divide."""
    if not isinstance(features, (list, tuple)):
        features = [features]
    _lib.require_cuda(pts, imgs, intrs, c2ws, *features)
    if any(f.shape[1] != 4 for f in features) or imgs.shape[1] != 3:
        raise RuntimeError("gens_b200.lookup_feature is built for 4-channel feature maps and RGB images")
    p = _lib.f32c(pts.reshape(-1, 3))
    w2c_src = _lib.invert_poses(c2ws[1:])  # torch.inverse(c2ws[1:]) of the reference (projector.py:322), one launch
    k_src = _lib.f32c(intrs[1:])
    return _LookupFeature.apply(p, w2c_src, k_src, _lib.f32c(c2ws[0]), _lib.f32c(c2ws[1:]), imgs[1:], *features)


# ---- feature-metric consistency patches (reference projector.py:353-437) ------------------------
def surface_patch_warp(pts_sdf0, gradients_sdf0, images, intrinsics, poses, patch_size=11):
    """Plane-induced homography warp of a patch_size^2 pixel patch around each surface point from the
    reference view into the sources.  Returns ((1,B,p*p,c), (ns,B,p*p,c)); differentiable w.r.t. pts."""
    import torch.nn.functional as F
    b = pts_sdf0.shape[0]
    if pts_sdf0.is_cuda and not (torch.is_grad_enabled() and (pts_sdf0.requires_grad or images.requires_grad)):
        # K8 (csrc/patch_warp.cu): one launch for the homographies and every sample of every view
        _lib.require_cuda(gradients_sdf0, images, intrinsics, poses)
        nv, c, h, w = images.shape
        pp = patch_size * patch_size
        imgs, k, p = _lib.f32c(images.detach()), _lib.f32c(intrinsics), _lib.f32c(poses)
        k_inv = _lib.invert_poses(k[:1])
        ref = torch.empty((1, b, pp, c), device=images.device, dtype=torch.float32)
        src = torch.empty((nv - 1, b, pp, c), device=images.device, dtype=torch.float32)
        # converted inputs stay referenced until the launch is enqueued (f32c may return a temporary)
        pts_c = _lib.f32c(pts_sdf0.detach().reshape(-1, 3))
        nrm_c = _lib.f32c(gradients_sdf0.detach().reshape(-1, 3))
        _lib.check(_lib.lib().gens_patch_warp(
            _lib.ptr(pts_c), _lib.ptr(nrm_c),
            _lib.ptr(imgs), _lib.ptr(k), _lib.ptr(p), _lib.ptr(k_inv), b, nv, c, h, w, int(patch_size), _lib.ptr(ref),
            _lib.ptr(src) if nv > 1 else None, _lib.stream_ptr(images.device)), "gens_patch_warp")
        del pts_c, nrm_c
        return ref, src
    r0, c0 = poses[0, :3, :3], poses[0, :3, 3]
    k0 = intrinsics[0, :3, :3]
    k0_inv = _lib.inverse(intrinsics)[0, :3, :3]
    x_ref = pts_sdf0 @ r0 - (c0 @ r0)[None, None, :]              # (B,1,3) point in the reference camera
    proj = x_ref @ k0.t()                                         # (B,1,3)
    disp = (gradients_sdf0 * x_ref).sum(-1, keepdim=True)         # n . X  (B,1,1)

    k_src = intrinsics[1:, :3, :3]
    ns = k_src.shape[0]
    r_src_t = poses[1:, :3, :3].transpose(1, 2)                   # (ns,3,3) world -> source camera
    r_rel = r_src_t @ r0                                          # (ns,3,3)
    t_rel = (r_src_t @ (c0[None, :] - poses[1:, :3, 3])[..., None])  # (ns,3,1)
    plane = t_rel[None] @ gradients_sdf0[:, None, :, :].expand(b, ns, 1, 3)   # (B,ns,3,3) = t n^T
    hom = k_src[None] @ (r_rel[None] + plane / (disp[:, None] + 1e-10)) @ k0_inv[None, None]

    centre = torch.stack([proj[:, 0, 0] / (proj[:, 0, 2] + 1e-8), proj[:, 0, 1] / (proj[:, 0, 2] + 1e-8)], -1).float()
    half = patch_size // 2
    r = torch.arange(-half, half + 1, device=centre.device, dtype=centre.dtype)
    off = torch.stack(torch.meshgrid(r, r, indexing="ij")[::-1], dim=-1).reshape(1, -1, 2)  # x fastest
    patch = centre[:, None, :] + off                               # (B,p*p,2)
    h, w = images.shape[-2:]
    npx = patch.shape[1]

    uv1 = torch.cat([patch, torch.ones_like(patch[..., :1])], dim=-1)  # (B,p*p,3)
    warped = torch.einsum("bsij,bpj->sbpi", hom, uv1).reshape(ns, -1, 3)
    g = warped[..., :2] / (warped[..., 2:] + 1e-8)
    gx = 2 * g[:, :, 0] / (w - 1) - 1.0
    gy = 2 * g[:, :, 1] / (h - 1) - 1.0
    src = F.grid_sample(images[1:], torch.stack([gx, gy], -1).view(ns, -1, 1, 2), align_corners=True)
    sampled = src.view(ns, -1, b, npx).permute(0, 2, 3, 1).contiguous()
    px = 2 * patch[..., 0] / (w - 1) - 1.0
    py = 2 * patch[..., 1] / (h - 1) - 1.0
    ref = F.grid_sample(images[:1], torch.stack([px, py], -1).detach().view(1, -1, 1, 2), align_corners=True)
    ref = ref.view(1, -1, b, npx).permute(0, 2, 3, 1).contiguous()
    return ref, sampled


# ---- K5: warp-per-ray hierarchical up-sampling (reference implicit_surface.py:60-133) --------------
def upsample_rays(rays_o, rays_d, z_vals, sdf, mask_volumes, inv_s: float, n_new: int):
    """up_sample + sample_pdf(det=True) fused, one warp per ray: (B,M) sorted depths/SDF -> (B,n_new)."""
    _lib.require_cuda(rays_o, rays_d, z_vals, sdf)
    b, m = z_vals.shape
    ms = [_lib.f32c(v) for v in _as_list(mask_volumes)]
    pyr = _lib.make_pyramid(ms, [v.shape[2] for v in ms])
    out = torch.empty((b, n_new), device=z_vals.device, dtype=torch.float32)
    # every converted input is held in `keep` until the launch is enqueued: f32c returns a temporary for
    # non-contiguous / non-fp32 callers (e.g. an expanded rays_o), which must not be freed and recycled before
    keep = [_lib.f32c(rays_o), _lib.f32c(rays_d), _lib.f32c(z_vals), _lib.f32c(sdf.reshape(b, m))]
    _lib.check(_lib.lib().gens_upsample_rays(
        _lib.ptr(keep[0]), _lib.ptr(keep[1]), _lib.ptr(keep[2]), _lib.ptr(keep[3]), b, m, pyr, ATEN_CUDA_FLAVOUR,
        float(inv_s), int(n_new), _lib.ptr(out), _lib.stream_ptr(z_vals.device)), "gens_upsample_rays")
    del keep
    return out


def merge_samples(z_vals, sdf, new_z, new_sdf=None):
    """Merge ascending new depths into the sorted ray (the sort/gather of cat_z_vals); SDF follows if given."""
    _lib.require_cuda(z_vals, new_z)
    b, m = z_vals.shape
    k = new_z.shape[1]
    dev = z_vals.device
    z_out = torch.empty((b, m + k), device=dev, dtype=torch.float32)
    with_sdf = new_sdf is not None
    sdf_out = torch.empty((b, m + k), device=dev, dtype=torch.float32) if with_sdf else None
    sdf_c = _lib.f32c(sdf.reshape(b, m)) if with_sdf else None
    nsdf_c = _lib.f32c(new_sdf.reshape(b, k)) if with_sdf else None
    z_c, nz_c = _lib.f32c(z_vals), _lib.f32c(new_z)  # held until the launch is enqueued (see upsample_rays)
    _lib.check(_lib.lib().gens_merge_samples(
        _lib.ptr(z_c), _lib.ptr(sdf_c) if with_sdf else None, _lib.ptr(nz_c),
        _lib.ptr(nsdf_c) if with_sdf else None, b, m, k, _lib.ptr(z_out), _lib.ptr(sdf_out) if with_sdf else None,
        _lib.stream_ptr(dev)), "gens_merge_samples")
    del z_c, nz_c, sdf_c, nsdf_c
    return z_out, sdf_out


# ---- K9: masked total variation of the pyramid (reference implicit_surface.py:135-150) --------------
# (ids, versions) of the volumes / masks the cached value was computed from -> (weakrefs, value)
_TV_CACHE: Dict[tuple, tuple] = {}
_TV_BLOCKS = 148 * 8


def tv_regularization(volume_feat_cas, volume_mask_cas=None) -> torch.Tensor:
    """sum_i 0.5^i * sqrt((tx_i + ty_i + tz_i) / (mx_i.sum() + 1e-8)) over the pyramid, no gradient: ONE
    launch visits every voxel once (csrc/tv_reg.cu) instead of ~12 full-volume ATen passes per scale.
    The value depends only on the volumes and masks, which do not change between the ray chunks of one
    image, so it is kept per (tensor objects, _version) -- render_core asks for it on every chunk."""
    vols = _as_list(volume_feat_cas)
    masks = None if volume_mask_cas is None else _as_list(volume_mask_cas)
    _lib.require_cuda(*vols, *(masks or []))
    every = vols + (masks or [])
    key = tuple((id(t), t._version) for t in every)
    hit = _TV_CACHE.get(key)
    if hit is not None and all(r() is t for r, t in zip(hit[0], every)):
        return hit[1]
    dev = vols[0].device
    for v in vols:
        if v.dim() != 5 or v.shape[0] != 1 or not (v.shape[2] == v.shape[3] == v.shape[4]):
            raise RuntimeError(f"tv_regularization needs (1,C,D,D,D) volumes, got {tuple(v.shape)}")
    channels = vols[0].shape[1]
    if any(v.shape[1] != channels for v in vols):
        raise RuntimeError("tv_regularization: all scales must have the same channel count")
    vs = [_lib.f32c(v.detach()) for v in vols]
    dims = [v.shape[2] for v in vs]
    ms = None
    if masks is not None:
        ms = [_lib.f32c(m.detach()) for m in masks]
        if [m.shape[2] for m in ms] != dims:
            raise RuntimeError("tv_regularization: mask / volume pyramid shapes differ")
    sums = torch.zeros((len(vs), 4), device=dev, dtype=torch.float64)
    _lib.check(_lib.lib().gens_tv_reduce(
        ctypes.byref(_lib.make_pyramid(vs, dims)), ctypes.byref(_lib.make_pyramid(ms, dims)) if ms else None,
        channels, _TV_BLOCKS, _lib.ptr(sums), _lib.stream_ptr(dev)), "gens_tv_reduce")
    decay = torch.tensor([0.5 ** i for i in range(len(vs))], device=dev, dtype=torch.float64)
    total = (torch.sqrt(sums[:, :3].sum(1) / (sums[:, 3] + 1e-8)) * decay).sum().float()
    _TV_CACHE.clear()  # one pyramid at a time
    _TV_CACHE[key] = ([weakref.ref(t) for t in every], total)
    return total


# ---- K7: alpha compositing, one warp per ray (reference implicit_surface.py:179-326) ----------------
def composite_rays(rays_o, rays_d, z_vals, pts, sdf_raw, grad_raw, smooth_raw, colour_raw, voxel_mask, evaluated,
                   mask_views, inv_s, rot, cos_anneal_ratio: float, sample_dist: float) -> Dict[str, torch.Tensor]:
    """Everything render_core computes between the network evaluations and its output dictionary, for
    inference (no autograd graph), in one launch (csrc/composite.cu).  voxel_mask / evaluated (n,) and
    mask_views (n,ns) are bool; inv_s is the (1,1) un-clipped deviation-network output; rot (3,3)."""
    _lib.require_cuda(rays_o, rays_d, z_vals, pts, sdf_raw, grad_raw, smooth_raw, colour_raw, voxel_mask, evaluated,
                      mask_views, inv_s, rot)
    b, n = z_vals.shape
    dev = z_vals.device
    f, u8 = _lib.f32c, lambda t: t.contiguous().view(torch.uint8) if t.dtype == torch.bool else t.to(torch.uint8).contiguous()
    keep = [f(rays_o), f(rays_d), f(z_vals), f(pts.reshape(-1, 3)), f(sdf_raw.reshape(-1)), f(grad_raw.reshape(-1, 3)),
            f(smooth_raw.reshape(-1, 3)), f(colour_raw.reshape(-1, 3)), u8(voxel_mask.reshape(-1)),
            u8(evaluated.reshape(-1)), u8(mask_views.reshape(b * n, -1)), f(inv_s.reshape(-1)),
            f(z_vals.max().reshape(1)), f(rot)]
    new = lambda *shape, dtype=torch.float32: torch.empty(shape, device=dev, dtype=dtype)
    out = {"weights": new(b, n), "weight_sum": new(b, 1), "weight_max": new(b, 1), "render_depth": new(b),
           "color_fine": new(b, 3), "normal": new(b, 3), "inside_sphere": new(b, n),
           "valid_mask": new(b, 1, dtype=torch.uint8), "sdf": new(b * n, 1), "gradients": new(b, n, 3),
           "mid_inside_sphere": new(b, 1), "sdf_depth": new(b, 1), "pts_sdf0": new(b, 1, 3), "ge_num": new(b),
           "ge_den": new(b), "smooth_norm": new(b)}
    a = _lib.CompositeArgs()
    a.n_rays, a.n_samples, a.n_src = b, n, keep[10].shape[1]
    a.cos_anneal_ratio, a.sample_dist = float(cos_anneal_ratio), float(sample_dist)
    for name, t in zip(("rays_o", "rays_d", "z_vals", "pts", "sdf_raw", "grad_raw", "smooth_raw", "colour_raw",
                        "voxel_mask", "evaluated", "mask_views", "inv_s", "z_max", "rot"), keep):
        setattr(a, name, t.data_ptr())
    for name, key in (("weights_out", "weights"), ("weight_sum_out", "weight_sum"), ("weight_max_out", "weight_max"),
                      ("depth_out", "render_depth"), ("color_out", "color_fine"), ("normal_out", "normal"),
                      ("inside_out", "inside_sphere"), ("valid_out", "valid_mask"), ("sdf_out", "sdf"),
                      ("gradients_out", "gradients"), ("mid_inside_out", "mid_inside_sphere"),
                      ("sdf_depth_out", "sdf_depth"), ("pts_sdf0_out", "pts_sdf0"), ("ge_num_out", "ge_num"),
                      ("ge_den_out", "ge_den"), ("smooth_norm_out", "smooth_norm")):
        setattr(a, name, out[key].data_ptr())
    _lib.check(_lib.lib().gens_composite_rays(ctypes.byref(a), _lib.stream_ptr(dev)), "gens_composite_rays")
    out["valid_mask"] = out["valid_mask"].bool()
    return out
