"""The small MLPs on the ray-marching path, state_dict-compatible with the reference so its
checkpoints load unchanged:

  SDFNetwork             reference models/modules/sdf_network.py:28-153   (keys lin{l}.weight_g/weight_v/bias)
  BlendingNetwork        reference models/modules/blending_network.py:22-117
  SingleVarianceNetwork  reference models/modules/variance_network.py:5-11

The dense layers stay on cuBLAS (plain library GEMMs, fp32: the 1e-4 parity budget rules out TF32);
what changes is everything around them -- the multi-scale volume look-up is one fused launch
(projector.lookup_volume) and the inference path (`SDFNetwork.sdf_nograd`) folds the weight
normalisation once per call and never concatenates activations with the encoded features.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import projector as _projector


def positional_encoding(x: torch.Tensor, n_freqs: int) -> torch.Tensor:
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)] (embedder.py:11-36)."""
    if n_freqs <= 0:
        return x
    parts = [x]
    for k in range(n_freqs):
        f = float(2.0 ** k)
        parts.append(torch.sin(x * f))
        parts.append(torch.cos(x * f))
    return torch.cat(parts, dim=-1)


def encoded_width(d: int, n_freqs: int) -> int:
    return d * (1 + 2 * n_freqs) if n_freqs > 0 else d


class SDFNetwork(nn.Module):
    def __init__(self, d_in, d_out, d_hidden, n_layers, skip_in=(4,), multires=0, bias=0.5, scale=1,
                 geometric_init=True, weight_norm=True, inside_outside=False, feat_channels=32, feat_multires=2,
                 lookup=None):
        super().__init__()
        self._lookup = _projector.lookup_volume if lookup is None else lookup
        self.multires, self.feat_multires = int(multires), int(feat_multires)
        self.init_feat_channels = feat_channels
        self.scale = scale
        self.skip_in = tuple(skip_in)
        pe_in = encoded_width(d_in, self.multires)
        pe_feat = encoded_width(feat_channels, self.feat_multires)
        self.pe_in, self.pe_feat = pe_in, pe_feat
        widths = [pe_in] + [d_hidden + pe_feat] * n_layers + [d_out]
        self.num_layers = len(widths)
        last = self.num_layers - 2
        for l in range(self.num_layers - 1):
            fan_out = widths[l + 1] - (widths[0] if (l + 1) in self.skip_in else 0)
            if l < last:
                fan_out -= pe_feat  # the encoded features are re-attached in front of every hidden layer
            lin = nn.Linear(widths[l], fan_out)
            if geometric_init:
                self._geometric_init(lin, l, last, widths, fan_out, bias, inside_outside, pe_feat)
            if weight_norm:
                lin = nn.utils.weight_norm(lin)
            setattr(self, f"lin{l}", lin)
        self.activation = nn.Softplus(beta=100)

    def _geometric_init(self, lin, l, last, widths, fan_out, bias, inside_outside, pe_feat):
        """Sphere initialisation of IDR/NeuS as the reference applies it (sdf_network.py:63-88): the
        columns fed by positional-encoding terms and by the volume features start at zero."""
        with torch.no_grad():
            if l == last:
                sign = -1.0 if inside_outside else 1.0
                nn.init.normal_(lin.weight, mean=sign * math.sqrt(math.pi) / math.sqrt(widths[l]), std=1e-4)
                nn.init.constant_(lin.bias, -sign * bias)
                lin.weight[:, -pe_feat:] = 0.0
                lin.bias[-pe_feat:] = 0.0
                return
            nn.init.constant_(lin.bias, 0.0)
            std = math.sqrt(2) / math.sqrt(fan_out)
            if self.multires > 0 and l == 0:
                lin.weight.zero_()
                nn.init.normal_(lin.weight[:, :3], 0.0, std)
            elif self.multires > 0 and l in self.skip_in:
                nn.init.normal_(lin.weight, 0.0, std)
                lin.weight[:, -(widths[0] - 3 + pe_feat):] = 0.0
            else:
                nn.init.normal_(lin.weight, 0.0, std)
                lin.weight[:, -pe_feat:] = 0.0

    # -- reference-shaped path (autograd at op granularity; training and .gradient) -------------
    def forward(self, inputs, volumes):
        feats = positional_encoding(self._lookup(inputs.clone(), volumes), self.feat_multires)
        pos = positional_encoding(inputs * self.scale, self.multires)
        x = pos
        for l in range(self.num_layers - 1):
            if l in self.skip_in:
                x = torch.cat([x, pos], -1) / math.sqrt(2)
            if 0 < l < self.num_layers - 1:
                x = torch.cat([x, feats], -1)
            x = getattr(self, f"lin{l}")(x)
            if l < self.num_layers - 2:
                x = self.activation(x)
        return torch.cat([x[:, :1] / self.scale, x[:, 1:]], dim=-1)

    def sdf(self, x, volumes):
        return self.forward(x, volumes)[:, :1]

    def sdf_hidden_appearance(self, x, volumes):
        return self.forward(x, volumes)

    @torch.enable_grad()
    def gradient(self, x, volumes):
        """(grad sdf, d/dx sum_k grad_k) with the graph kept, as sdf_network.py:131-153."""
        x.requires_grad_(True)
        y = self.sdf(x, volumes)
        (g,) = torch.autograd.grad(y, x, torch.ones_like(y), create_graph=True, retain_graph=True)
        (h,) = torch.autograd.grad(g, x, torch.ones_like(g), create_graph=True, retain_graph=True)
        return g, h

    # -- inference path ---------------------------------------------------------------------------
    def value_grad_smooth_nograd(self, pts, volumes, need_smooth=True, folded=None):
        """(sdf, grad, smooth) without an autograd graph (CUDA only; see gens_b200/sdf_analytic.py).  `folded` = a
        FoldedSDF of the current weights (ImplicitSurface keeps one per parameter version)."""
        from . import sdf_analytic
        fw = folded if isinstance(folded, sdf_analytic.FoldedSDF) else None
        return sdf_analytic.value_grad_smooth(self, pts, volumes, fw, need_smooth)

    def folded_weights(self):
        """Effective (weight, bias) per layer with the weight normalisation folded in."""
        out = []
        for l in range(self.num_layers - 1):
            lin = getattr(self, f"lin{l}")
            if hasattr(lin, "weight_g"):
                w = torch._weight_norm(lin.weight_v, lin.weight_g, 0)
            else:
                w = lin.weight
            out.append((w, lin.bias))
        return out

    @torch.no_grad()
    def sdf_nograd(self, pts, volumes, folded=None):
        """SDF values (n,1) without autograd bookkeeping: same arithmetic as forward()[:, :1], but the
        encoded volume features enter every layer through ONE (n,100)x(100,sum fan_out) GEMM instead of
        six concatenations, and only column 0 of the output layer is evaluated."""
        if self._lookup is _projector.lookup_volume and pts.is_cuda:
            from . import sdf_analytic  # fused CUDA stages (csrc/sdf_glue.cu); `folded` may be a FoldedSDF
            fw = folded if isinstance(folded, sdf_analytic.FoldedSDF) else None
            return sdf_analytic.value_only(self, pts, volumes, fw)
        folded = self.folded_weights() if folded is None or not isinstance(folded, list) else folded
        feats = positional_encoding(self._lookup(pts, volumes), self.feat_multires)
        pos = positional_encoding(pts * self.scale, self.multires)
        last = self.num_layers - 2
        # feature columns of layers 1..last, stacked
        wf = torch.cat([folded[l][0][: (1 if l == last else None), -self.pe_feat:] for l in range(1, last + 1)], 0)
        feat_part = feats @ wf.t()
        off = 0
        x = F.softplus(F.linear(pos, folded[0][0], folded[0][1]), beta=100)
        for l in range(1, last + 1):
            w, b = folded[l]
            rows = 1 if l == last else w.shape[0]
            if l in self.skip_in:
                x = torch.cat([x, pos], -1) / math.sqrt(2)
            y = torch.addmm(feat_part[:, off:off + rows] + b[:rows], x, w[:rows, : x.shape[1]].t())
            off += rows
            x = y if l == last else F.softplus(y, beta=100)
        return x / self.scale


def _elu_mlp(sizes: Sequence[int], final_act: bool = True, sigmoid: bool = False) -> nn.Sequential:
    """Linear/ELU stack whose module indices match the reference's nn.Sequential layouts."""
    layers: List[nn.Module] = []
    for i in range(len(sizes) - 1):
        layers.append(nn.Linear(sizes[i], sizes[i + 1]))
        if i < len(sizes) - 2 or final_act:
            layers.append(nn.ELU(inplace=True))
    if sigmoid:
        layers.append(nn.Sigmoid())
    return nn.Sequential(*layers)


def _kaiming(m):
    if isinstance(m, nn.Linear):
        nn.init.kaiming_normal_(m.weight.data)
        if m.bias is not None:
            nn.init.zeros_(m.bias.data)


class BlendingNetwork(nn.Module):
    """IBRNet-style colour blending over the source views (blending_network.py:22-117)."""

    def __init__(self, d_feature=16, anti_alias_pooling=True):
        super().__init__()
        self.anti_alias_pooling = anti_alias_pooling
        if anti_alias_pooling:
            self.s = nn.Parameter(torch.tensor(0.2), requires_grad=True)
        c = d_feature + 3
        self.ray_dir_fc = _elu_mlp([4, 16, c])
        self.base_fc = _elu_mlp([3 * c, 64, 32])
        self.vis_fc = _elu_mlp([32, 32, 33])
        self.vis_fc2 = _elu_mlp([32, 32, 1], final_act=False, sigmoid=True)
        self.rgb_fc = _elu_mlp([32 + 1 + 4, 16, 8, 1], final_act=False)
        for net in (self.base_fc, self.vis_fc2, self.vis_fc, self.rgb_fc):
            net.apply(_kaiming)

    # -- inference path: the whole network as one kernel (K10, csrc/blend.cu) -----------------------------
    def packed_weights(self) -> torch.Tensor:
        """The eleven Linear layers re-ordered into the shared-memory image of blend_kernel: "A" layers as
        [out/4][in][4], the "B" layer that follows as [in/4][ceil(out/2)][4 inputs][2 outputs] (an odd output count is
        padded with a zero row), biases padded to float4 (offsets: the constexpr table at the top of csrc/blend.cu).
        Cached per parameter versions."""
        params = list(self.parameters())
        key = tuple((p.data_ptr(), p._version) for p in params)
        if getattr(self, "_packed_key", None) == key:
            return self._packed
        dev = params[0].device

        def a_layout(w):                      # (out, in) -> [out/4][in][4]
            out, inp = w.shape
            return w.reshape(out // 4, 4, inp).permute(0, 2, 1).reshape(-1)

        def b_layout(w):                      # (out, in) -> [in/4][out/2][4][2]
            w = F.pad(w, (0, 0, 0, w.shape[0] % 2))
            out, inp = w.shape
            return w.reshape(out // 2, 2, inp // 4, 4).permute(2, 0, 3, 1).reshape(-1)

        def pad4(v):
            v = v.reshape(-1)
            return F.pad(v, (0, (-v.numel()) % 4))

        rd, base, vis, vis2, rgb = self.ray_dir_fc, self.base_fc, self.vis_fc, self.vis_fc2, self.rgb_fc
        c = rd[2].weight.shape[0]
        if c != 23 or base[0].weight.shape != (64, 3 * c) or rgb[0].weight.shape != (16, 37) or \
                not self.anti_alias_pooling:
            raise RuntimeError("gens_b200 blend kernel is built for d_feature = 20 with anti-alias pooling")
        with torch.no_grad():
            w3 = base[0].weight
            pieces = [a_layout(rd[0].weight), rd[0].bias, b_layout(rd[2].weight), pad4(rd[2].bias),
                      a_layout(w3[:, : 2 * c]), base[0].bias, a_layout(w3[:, 2 * c:]),
                      b_layout(base[2].weight), base[2].bias,
                      a_layout(vis[0].weight), vis[0].bias, b_layout(vis[2].weight), pad4(vis[2].bias),
                      a_layout(vis2[0].weight), vis2[0].bias, b_layout(vis2[2].weight), pad4(vis2[2].bias),
                      a_layout(rgb[0].weight), rgb[0].bias, b_layout(rgb[2].weight), rgb[2].bias,
                      rgb[4].weight.reshape(-1), rgb[4].bias.reshape(-1), torch.abs(self.s).reshape(1),
                      torch.zeros(2, device=dev)]
            flat = torch.cat([p.reshape(-1).float() for p in pieces]).contiguous()
        from . import _lib
        if flat.numel() != _lib.lib().gens_blend_weight_floats():
            raise RuntimeError(f"packed blending weights: {flat.numel()} floats, the kernel expects "
                               f"{_lib.lib().gens_blend_weight_floats()}")
        self._packed_key, self._packed = key, flat
        return flat

    @torch.no_grad()
    def blend_nograd(self, rgb_feat, ray_diff, mask):
        """forward() without autograd, one launch (CUDA only)."""
        from . import _lib
        _lib.require_cuda(rgb_feat, ray_diff, mask)
        n, ns = mask.shape
        rf, rdiff = _lib.f32c(rgb_feat), _lib.f32c(ray_diff)
        m = mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8).contiguous()
        w = self.packed_weights()
        out = torch.empty((n, 3), device=rf.device, dtype=torch.float32)
        _lib.check(_lib.lib().gens_blend_colour(_lib.ptr(rf), _lib.ptr(rdiff), _lib.ptr(m), n, ns, _lib.ptr(w),
                                                _lib.ptr(out), _lib.stream_ptr(rf.device)), "gens_blend_colour")
        return out

    def forward(self, rgb_feat, ray_diff, mask):
        """rgb_feat (n,ns,3+c), ray_diff (n,ns,4), mask (n,ns) -> rgb (n,3)."""
        m = mask[:, :, None]
        ns = rgb_feat.shape[1]
        rgb_in = rgb_feat[..., :3]
        feat = rgb_feat + self.ray_dir_fc(ray_diff)
        if self.anti_alias_pooling:
            dot = ray_diff[..., 3:4]
            e = torch.exp(torch.abs(self.s) * (dot - 1))
            w = (e - e.min(dim=1, keepdim=True)[0]) * m
            w = w / (w.sum(dim=1, keepdim=True) + 1e-8)
        else:
            w = m / (m.sum(dim=1, keepdim=True) + 1e-8)
        mean = (feat * w).sum(dim=1, keepdim=True)
        var = (w * (feat - mean) ** 2).sum(dim=1, keepdim=True)
        x = torch.cat([torch.cat([mean, var], -1).expand(-1, ns, -1), feat], dim=-1)
        x = self.base_fc(x)
        x_vis = self.vis_fc(x * w)
        x_res, vis = x_vis[..., :-1], x_vis[..., -1:]
        vis = torch.sigmoid(vis) * m
        x = x + x_res
        vis = self.vis_fc2(x * vis) * m
        logits = self.rgb_fc(torch.cat([x, vis, ray_diff], dim=-1)).masked_fill(m == 0, -1e9)
        return (rgb_in * F.softmax(logits, dim=1)).sum(dim=1)


class SingleVarianceNetwork(nn.Module):
    def __init__(self, init_val):
        super().__init__()
        self.register_parameter("variance", nn.Parameter(torch.tensor(init_val)))

    def forward(self, x):
        return torch.ones([len(x), 1]).type_as(x) * torch.exp(self.variance * 10.0)
