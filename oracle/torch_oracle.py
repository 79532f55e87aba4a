"""ATen-op restatement of the GenS hot path -- TEST INFRASTRUCTURE and CPU/GPU baseline arm.

The reference's hot path is thin Python over PyTorch ATen ops (the third-party dependency
that holds the arithmetic: torch==1.13.1 pinned by the reference, torch 2.11 here).  This
module restates the same op sequence, device-agnostic, so that
  * on the GPU box it gives the reference-path numerics of ATen's CUDA kernels (mask parity
    in GENS_DIV_RECIP mode) and the "reference ops on the same B200" timing, and
  * on host cores it is the multi-threaded CPU baseline of bench.py (`--impl reference`).
It is pinned against the golden fixtures produced by the real reference
(tests/test_oracle_golden.py).  Never imported by gens_b200/.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def voxel_centres(d: int, like: torch.Tensor) -> torch.Tensor:
    """(3, d^3) world coordinates; flat index a*d*d + b*d + c <-> (g[a], g[b], g[c]).
    Reference: volume.py:28-31."""
    g = torch.linspace(-1, 1, d).type_as(like)
    return torch.stack(torch.meshgrid(g, g, g, indexing="ij")).reshape(3, -1)


def project_voxels(intrs, c2ws, scale: int, d: int, hw):
    """Normalised sampling grid (nv, d^3, 2) and per-view validity (nv, d^3).
    Reference: volume.py:24-25, :32-43."""
    h, w = hw
    nv = intrs.shape[0]
    k = intrs.clone()
    k[:, :2] *= 0.5 ** scale
    xyz = voxel_centres(d, k)
    homo = torch.cat([xyz, torch.ones_like(xyz[:1])], 0).unsqueeze(0).repeat(nv, 1, 1)
    img = torch.matmul(k, torch.matmul(torch.inverse(c2ws), homo))[:, :3]
    xy = img[:, :2] / (img[:, 2:] + 1e-8)
    nx = xy[:, 0] / ((w - 1) / 2) - 1
    ny = xy[:, 1] / ((h - 1) / 2) - 1
    valid = (nx.abs() <= 1) & (ny.abs() <= 1) & (img[:, 2] > 0)
    return torch.stack([nx, ny], -1), valid


def agg_mean_var_scale(feat, intrs, c2ws, scale: int, d: int, min_vis_view: int = 1):
    """One scale of Volume.agg_mean_var (reference volume.py:21-58) -> (1,2c,d,d,d), (1,1,d,d,d)."""
    nv, c, h, w = feat.shape
    with torch.no_grad():
        grid, valid = project_voxels(intrs, c2ws, scale, d, (h, w))
    m = valid.unsqueeze(1)
    warped = F.grid_sample(feat, grid.unsqueeze(1), padding_mode="zeros", align_corners=True).squeeze(2) * m
    total, total_sq, n = warped.sum(0), (warped ** 2).sum(0), m.sum(0)
    n_safe = torch.where(n <= 0, torch.ones_like(n) * 1e-8, n)
    mean = total / n_safe
    var = total_sq / n_safe - mean ** 2
    vol = torch.cat([mean, var], 0).reshape(1, 2 * c, d, d, d)
    return vol, (n > min_vis_view).float().reshape(1, 1, d, d, d)


def agg_mean_var(features, intrs, c2ws, dims, min_vis_view: int = 1):
    out = [agg_mean_var_scale(features[i], intrs, c2ws, i, d, min_vis_view) for i, d in enumerate(dims)]
    return [o[0] for o in out], [o[1] for o in out]


def corner_indices(grid, hw):
    """floor() corner of the bilinear footprint, as ATen's align_corners=True un-normalise gives it."""
    h, w = hw
    ix = ((grid[..., 0] + 1) / 2) * (w - 1)
    iy = ((grid[..., 1] + 1) / 2) * (h - 1)
    return torch.floor(ix).to(torch.int32), torch.floor(iy).to(torch.int32)


# ---- lookup_volume (reference projector.py:217-245) -------------------------------------------
def trilinear_dd(vol, pts):
    """Double-differentiable trilinear look-up of a (1,C,D,D,D) volume at (n,3) points, built from
    elementary torch ops (ATen's grid_sampler_3d_backward has no derivative of its own; the reference
    needs its CUDA-only grad2 kernel for that).  p=(p0,p1,p2) addresses tensor dims (2,3,4);
    align_corners=True, zeros padding.  Returns (n,C)."""
    _, c, d, _, _ = vol.shape
    u = ((pts + 1) / 2) * (d - 1)
    i0 = torch.floor(u).detach()
    t = u - i0
    flat = vol.reshape(c, -1)
    out = 0
    for da in (0, 1):
        for db in (0, 1):
            for dc in (0, 1):
                a, b, cc = i0[:, 0] + da, i0[:, 1] + db, i0[:, 2] + dc
                w = (t[:, 0] if da else 1 - t[:, 0]) * (t[:, 1] if db else 1 - t[:, 1]) * (t[:, 2] if dc else 1 - t[:, 2])
                ok = (a >= 0) & (a < d) & (b >= 0) & (b < d) & (cc >= 0) & (cc < d)
                idx = (a.clamp(0, d - 1) * d + b.clamp(0, d - 1)) * d + cc.clamp(0, d - 1)
                out = out + (flat[:, idx.long()] * ok.to(vol.dtype) * w).t()
    return out


def lookup_volume(pts, volumes, sample_mode="grad"):
    """ATen-op restatement of projector.lookup_volume for a list of volumes -> (n, sum C)."""
    grid = pts.reshape(1, 1, 1, -1, 3).flip(-1)
    outs = []
    for v in volumes:
        if sample_mode == "grad":
            o = F.grid_sample(v, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
        else:
            o = F.grid_sample(v, grid, mode="nearest", padding_mode="zeros", align_corners=False)
        outs.append(o.reshape(v.shape[1], -1).t())
    return torch.cat(outs, -1)


# ---- the reference's three-level autograd of the volume look-up, in ATen ops --------------------
def _cat_dd(vols, pts):
    return torch.cat([trilinear_dd(v, pts) for v in vols], -1)


class _RefLookup(torch.autograd.Function):
    """Forward of cug.grid_sample_3d over a list of volumes (cuda_gridsample.py:71-91)."""

    @staticmethod
    def forward(ctx, pts, *vols):
        ctx.save_for_backward(pts, *vols)
        with torch.no_grad():
            return _cat_dd(vols, pts)

    @staticmethod
    def backward(ctx, g_out):
        pts, *vols = ctx.saved_tensors
        res = _RefLookupBackward.apply(g_out, pts, *vols)
        return (res[0], *res[1:])


class _RefLookupBackward(torch.autograd.Function):
    """aten::grid_sampler_3d_backward, whose own backward is the reference's grad2_3d: its results
    carry NO grad_fn (cuda_gridsample.py:110-123), i.e. third-order terms are dropped.  The oracle must
    drop them too, otherwise training gradients would differ from the reference by those terms."""

    @staticmethod
    def forward(ctx, g_out, pts, *vols):
        ctx.save_for_backward(g_out, pts, *vols)
        with torch.enable_grad():
            p = pts.detach().requires_grad_(True)
            vs = [v.detach().requires_grad_(True) for v in vols]
            y = _cat_dd(vs, p)
            grads = torch.autograd.grad(y, [p] + vs, g_out.detach(), allow_unused=True)
        return tuple(g.detach() if g is not None else torch.zeros_like(t) for g, t in zip(grads, [pts] + list(vols)))

    @staticmethod
    def backward(ctx, gg_pts, *gg_vols):
        g_out, pts, *vols = ctx.saved_tensors
        with torch.enable_grad():
            go = g_out.detach().requires_grad_(True)
            p = pts.detach().requires_grad_(True)
            vs = [v.detach().requires_grad_(True) for v in vols]
            y = _cat_dd(vs, p)
            (g_pts,) = torch.autograd.grad(y, p, go, create_graph=True)
            res = torch.autograd.grad((g_pts * gg_pts.detach()).sum(), [go, p] + vs, allow_unused=True)
        return tuple(r.detach() if r is not None else None for r in res)


# ---- ATen-on-CPU provider of the ray marcher's look-up ops ------------------------------------
class CpuOps:
    """Drop-in for gens_b200.projector as the `ops` of gens_b200.implicit_surface.ImplicitSurface:
    the same four look-ups expressed with ATen ops (device-agnostic, no CUDA extension).  Used (a)
    by the CPU tests that pin the product's HOST logic to the reference's golden vectors and (b) as
    the CPU baseline of bench.py's render metric.  `nearest_fused` is irrelevant here: ATen itself
    picks the flavour of the device it runs on."""

    @staticmethod
    def lookup_volume(pts, volume, sample_mode="grad"):
        vols = [volume] if isinstance(volume, torch.Tensor) else list(volume)
        pts = pts.reshape(-1, 3)
        if sample_mode == "grad":
            if torch.is_grad_enabled() and (pts.requires_grad or any(v.requires_grad for v in vols)):
                # the reference's three-level autograd (its CUDA-only grad2 op expressed in ATen ops)
                return _RefLookup.apply(pts, *vols)
            return lookup_volume(pts, vols, "grad")
        return lookup_volume(pts, vols, "nearest")

    @staticmethod
    def mask_nearest(pts, masks, want_each=False):
        each = CpuOps.lookup_volume(pts, masks, "nearest")
        return each if want_each else each.any(dim=-1)

    @staticmethod
    def lookup_feature(pts, imgs, intrs, c2ws, features):
        """Reference projector.py:278-349 restated (see gens_b200.projector.lookup_feature for the contract)."""
        if not isinstance(features, (list, tuple)):
            features = [features]
        src_k, src_c2w, ref_c = intrs[1:], c2ws[1:], c2ws[0, :3, 3]
        a = ref_c[None, None] - pts[None]
        a = a / (torch.norm(a, dim=-1, keepdim=True) + 1e-6)
        b = src_c2w[:, :3, 3][:, None] - pts[None]
        b = b / (torch.norm(b, dim=-1, keepdim=True) + 1e-6)
        d = a - b
        ray_diff = torch.cat([d / torch.clamp(torch.norm(d, dim=-1, keepdim=True), min=1e-6),
                              (a * b).sum(-1, keepdim=True)], -1).permute(1, 0, 2).contiguous()
        ns, n = src_k.shape[0], pts.shape[0]
        homo = torch.cat([pts.t(), pts.new_ones(1, n)], 0)
        w2c = torch.inverse(src_c2w)
        out, masks, rgb = [], [], None
        for i, feat in enumerate(features):
            with torch.no_grad():
                k = src_k.clone()
                k[:, :2] = k[:, :2] * (0.5 ** i)
                h, w = feat.shape[-2:]
                img = torch.matmul(k[:, :3, :3], torch.matmul(w2c, homo[None])[:, :3])
                xy = img[:, :2] / img[:, 2:]
                nx, ny = xy[:, 0] / ((w - 1) / 2) - 1, xy[:, 1] / ((h - 1) / 2) - 1
                masks.append(((img[:, 2] > 0) & (xy[:, 0] >= 0) & (xy[:, 0] < w) & (xy[:, 1] >= 0) & (xy[:, 1] < h)).t())
                grid = torch.stack([nx, ny], -1).unsqueeze(2)
            out.append(F.grid_sample(feat[1:], grid, mode="bilinear", padding_mode="zeros", align_corners=False)
                       .reshape(ns, feat.shape[1], n).permute(2, 0, 1))
            if i == 0:
                rgb = F.grid_sample(imgs[1:], grid, mode="bilinear", padding_mode="zeros", align_corners=False) \
                    .reshape(ns, 3, n).permute(2, 0, 1)
        return (torch.cat([rgb] + out, 2).float().contiguous(), ray_diff,
                torch.stack(masks, -1).all(-1).contiguous())

    @staticmethod
    def surface_patch_warp(pts_sdf0, normals, images, intrinsics, poses, patch_size=11):
        """Reference projector.py:353-437 restated: plane-induced homography patches."""
        b = pts_sdf0.shape[0]
        r0, c0 = poses[0, :3, :3], poses[0, :3, 3]
        k0, k0_inv = intrinsics[0, :3, :3], torch.inverse(intrinsics)[0, :3, :3]
        x_ref = pts_sdf0 @ r0 - (c0 @ r0)[None, None]
        proj = x_ref @ k0.t()
        disp = (normals * x_ref).sum(-1, keepdim=True)
        ns = intrinsics.shape[0] - 1
        r_src_t = poses[1:, :3, :3].transpose(1, 2)
        t_rel = r_src_t @ (c0[None] - poses[1:, :3, 3])[..., None]
        hom = intrinsics[1:, :3, :3][None] @ (
            (r_src_t @ r0)[None] + (t_rel[None] @ normals[:, None].expand(b, ns, 1, 3)) / (disp[:, None] + 1e-10)
        ) @ k0_inv[None, None]
        centre = torch.stack([proj[:, 0, 0] / (proj[:, 0, 2] + 1e-8), proj[:, 0, 1] / (proj[:, 0, 2] + 1e-8)], -1)
        half = patch_size // 2
        r = torch.arange(-half, half + 1, device=centre.device, dtype=centre.dtype)
        off = torch.stack(torch.meshgrid(r, r, indexing="ij")[::-1], -1).reshape(1, -1, 2)
        patch = centre[:, None] + off
        h, w = images.shape[-2:]
        npx = patch.shape[1]
        warped = torch.einsum("bsij,bpj->sbpi", hom, torch.cat([patch, torch.ones_like(patch[..., :1])], -1))
        warped = warped.reshape(ns, -1, 3)
        g = warped[..., :2] / (warped[..., 2:] + 1e-8)
        grid = torch.stack([2 * g[..., 0] / (w - 1) - 1, 2 * g[..., 1] / (h - 1) - 1], -1).view(ns, -1, 1, 2)
        src = F.grid_sample(images[1:], grid, align_corners=True).view(ns, -1, b, npx).permute(0, 2, 3, 1)
        pgrid = torch.stack([2 * patch[..., 0] / (w - 1) - 1, 2 * patch[..., 1] / (h - 1) - 1], -1)
        ref = F.grid_sample(images[:1], pgrid.detach().view(1, -1, 1, 2), align_corners=True)
        return ref.view(1, -1, b, npx).permute(0, 2, 3, 1).contiguous(), src.contiguous()


# ---- loss-side consumers of the hot path's outputs (reference models/losses/ncc.py:7-50, loss.py:23-84) ----------
def compute_lncc(ref_gray, src_grays):
    """Patch NCC score, restating compute_LNCC (ncc.py:7-50).  ref_gray (1,B,P,C), src_grays (S,B,P,C) with
    P = patch^2 samples.  The reference runs five grouped patch x patch convolutions over zero-padded patches only to
    read the centre pixel, i.e. plain sums over the P samples; the sums are taken directly here.  Returns (B,1)."""
    ref = ref_gray.permute(1, 0, 3, 2)   # (B,1,C,P)
    src = src_grays.permute(1, 0, 3, 2)  # (B,S,C,P)
    npatch = src.shape[-1]
    ref_sum, src_sum = ref.sum(-1), src.sum(-1)
    ref_sq_sum, src_sq_sum = ref.pow(2).sum(-1), src.pow(2).sum(-1)
    ref_src_sum = (ref * src).sum(-1)
    u_ref, u_src = ref_sum / npatch, src_sum / npatch
    cross = ref_src_sum - u_src * ref_sum - u_ref * src_sum + u_ref * u_src * npatch
    ref_var = ref_sq_sum - 2 * u_ref * ref_sum + u_ref * u_ref * npatch
    src_var = src_sq_sum - 2 * u_src * src_sum + u_src * u_src * npatch
    cc = cross * cross / (ref_var * src_var + 1e-5)
    ncc = torch.clamp(1 - cc, 0.0, 2.0).mean(dim=2)          # (B,S)
    ncc, _ = torch.topk(ncc, 2, dim=1, largest=False)
    return ncc.mean(dim=1, keepdim=True)


def loss_forward(preds, targets, w):
    """Loss.forward (loss.py:23-84) for the keys the synthetic fixtures provide; `w` = the train.loss conf block."""
    valid = preds["valid_mask"].float()
    color = (F.l1_loss(preds["color_fine"], targets["color"], reduction="none") * valid).sum() / (valid.sum() + 1e-5)
    eikonal = preds["gradient_error"].mean()
    sparse = torch.exp(-torch.abs(preds["sparse_sdf"]) * w["sparse_scale_factor"]).mean()
    smooth = preds["smooth_error"].mean()
    tv = preds["tv_reg"].mean()
    ncc = compute_lncc(preds["ref_gray_val"], preds["sampled_gray_val"])
    ncc_mask = valid * preds["mid_inside_sphere"]
    mfc = 0.5 * ((ncc * ncc_mask).sum(dim=0) / (ncc_mask.sum(dim=0) + 1e-8)).squeeze(-1)
    pseudo = torch.abs(preds["pseudo_sdf"]).mean() if "pseudo_sdf" in preds else torch.zeros((), device=mfc.device)
    loss = (color * w["color_weight"] + eikonal * w["igr_weight"] + sparse * w["sparse_weight"] + mfc * w["mfc_weight"]
            + smooth * w["smooth_weight"] + tv * w["tv_weight"] + pseudo * w["pseudo_sdf_weight"])
    return {"loss": loss, "color_loss": color, "eikonal_loss": eikonal, "sparse_loss": sparse, "mfc_loss": mfc,
            "smooth_loss": smooth, "tv_loss": tv, "pseudo_sdf_loss": pseudo}
