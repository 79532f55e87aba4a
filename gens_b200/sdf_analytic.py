"""Analytic value / gradient / second-order pass of the SDF network, without an autograd graph.

What `SDFNetwork.forward` + `SDFNetwork.gradient` (reference sdf_network.py:98-153) deliver through two
nested `torch.autograd.grad(create_graph=True)` calls, computed forward-over-reverse along the fixed
direction u = (1,1,1) (the reference's `d_output2 = ones`):

    sdf(p),   grad = d sdf / dp,   smooth = d/dp ( sum_k grad_k ) = H(p) . u

Work matrices carry 2n rows (primal on top, tangent below) so every layer is one SGEMM over 2n rows in
each direction (cuBLAS, plain library GEMM, fp32); the encoded volume features enter all layers through a
single (2n,100)x(100,614) GEMM and leave through its transpose.  Everything between the GEMMs is fused in
csrc/sdf_glue.cu; the two volume look-up passes (value+JVP, reverse through value and tangent) are K3
kernels of csrc/sampling.cu.  GEMM work is 4x a forward pass instead of the ~7x of the autograd graph,
and no activation is kept beyond sp'(a), sp''(a)da.
"""
from __future__ import annotations

import ctypes
import math
from typing import List, Optional, Tuple

import torch

from . import _lib
from .projector import packed_volume

_U = (ctypes.c_float * 3)(1.0, 1.0, 1.0)

# value-only SDF evaluations run the whole MLP as one tcgen05 kernel (csrc/sdf_mlp_tc.cu, 3xTF32 with fp32
# accumulation, ~1e-5 of the fp32 path); False = per-layer cuBLAS SGEMMs + fused bias/softplus kernels.
USE_TC = True


class FoldedSDF:
    """Weight-normalised layers folded and split once per call: x-part / feature-part per layer."""

    def __init__(self, net):
        self.n_layers = net.num_layers - 1          # 7 linear layers
        self.skip_in = net.skip_in
        self.scale = float(net.scale)
        self.multires, self.feat_multires = net.multires, net.feat_multires
        self.pe_in, self.pe_feat = net.pe_in, net.pe_feat
        self.n_feat = net.init_feat_channels
        last = self.n_layers - 1
        folded = net.folded_weights()
        self.wx: List[torch.Tensor] = []            # (fan_out_l, K_l) x-part, contiguous
        self.bias: List[torch.Tensor] = []
        wf = []
        self.fo: List[int] = []
        for l, (w, b) in enumerate(folded):
            rows = 1 if l == last else w.shape[0]   # only the SDF row of the output layer matters
            w, b = w[:rows], b[:rows]
            if l == 0:
                self.wx.append(w.contiguous())
            else:
                self.wx.append(w[:, : w.shape[1] - self.pe_feat].contiguous())
                wf.append(w[:, w.shape[1] - self.pe_feat:])
            self.bias.append(b.contiguous())
            self.fo.append(rows)
        self.wf = torch.cat(wf, 0).contiguous()     # (sum fan_out_{1..last}, pe_feat)
        self.wf_t = self.wf.t().contiguous()
        self.wx_t = [w.t().contiguous() for w in self.wx]
        self.off = [0]
        for l in range(1, self.n_layers):
            self.off.append(self.off[-1] + self.fo[l])  # off[l-1] = column of layer l in the feature part
        self._packed = None
        self._packed_rev = None
        self._tc_ok = None

    def tc_supported(self) -> bool:
        """True when the tensor-core kernels take this network (layers up to 128 wide, encodings that fit the
        resident operands); otherwise the cuBLAS fp32 chain of the same functions is used."""
        if self._tc_ok is None:
            try:
                self.packed()
                self.packed_rev()
                self._tc_ok = True
            except RuntimeError:
                self._tc_ok = False
        return self._tc_ok

    def packed_rev(self):
        """The transposed network for the tensor-core reverse sweep (built on first use)."""
        if self._packed_rev is None:
            from .mlp_tc import PackedSDFReverse
            self._packed_rev = PackedSDFReverse(self)
        return self._packed_rev

    def packed(self):
        """The same weights in the streaming format of the tensor-core kernel (built on first use)."""
        if self._packed is None:
            from .mlp_tc import PackedSDF
            self._packed = PackedSDF(self)
        return self._packed


def _c(code, what):
    _lib.check(code, what)


@torch.no_grad()
def value_grad_smooth(net, pts: torch.Tensor, volumes, folded: Optional[FoldedSDF] = None,
                      need_smooth: bool = True) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """(sdf (n,1), grad (n,3), smooth (n,3)) at `pts` (n,3); no gradient graph is built."""
    _lib.require_cuda(pts)
    L = _lib.lib()
    fw = FoldedSDF(net) if folded is None else folded
    pts = _lib.f32c(pts.reshape(-1, 3))
    n = pts.shape[0]
    dev = pts.device
    st = _lib.stream_ptr(dev)
    vols = [volumes] if isinstance(volumes, torch.Tensor) else list(volumes)
    packed = [packed_volume(v) for v in vols]
    pyr = _lib.make_pyramid(packed, [v.shape[2] for v in vols])
    nf = 4 * len(vols)
    if nf != fw.n_feat:
        raise RuntimeError(f"SDF network expects {fw.n_feat} volume features, the pyramid provides {nf}")
    new = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
    P = _lib.ptr

    feats, dfeats = new(n, nf), new(n, nf)
    _c(L.gens_trilinear_fwd_jvp(P(pts), n, pyr, _U, P(feats), P(dfeats), st), "gens_trilinear_fwd_jvp")
    pos, fe = new(2 * n, fw.pe_in), new(2 * n, fw.pe_feat)
    _c(L.gens_sdf_encode(P(pts), P(feats), P(dfeats), n, fw.scale, _U, fw.multires, fw.feat_multires, nf, P(pos),
                         P(fe), st), "gens_sdf_encode")
    if USE_TC and fw.tc_supported():
        # whole MLP on the tensor cores: one persistent kernel forward (value + tangent), one in reverse
        from . import mlp_tc
        sdf, tape = mlp_tc.sdf_jvp(fw.packed(), pos, fe, n)
        g_pos, g_fe = mlp_tc.sdf_reverse(fw.packed_rev(), tape, n)
        del tape
        grad, smooth = new(n, 3), new(n, 3)
        g_f, dg_f = new(n, nf), new(n, nf)
        _c(L.gens_sdf_decode(P(pts), P(feats), P(dfeats), P(g_pos), P(g_fe), n, fw.scale, _U, fw.multires,
                             fw.feat_multires, nf, P(g_f), P(dg_f), P(grad), P(smooth), st), "gens_sdf_decode")
        _c(L.gens_trilinear_vjp2(P(pts), n, pyr, _U, P(g_f), P(dg_f), P(grad), P(smooth) if need_smooth else None, st),
           "gens_trilinear_vjp2")
        return sdf, grad, (smooth if need_smooth else None)
    featpart = fe @ fw.wf_t                                   # (2n, sum fan_out)
    ldfp = featpart.shape[1]
    last = fw.n_layers - 1
    inv_sqrt2 = 1.0 / math.sqrt(2.0)

    # ---- forward (primal + tangent) ------------------------------------------------------------
    sp1, sp2 = [], []
    x = pos
    for l in range(last):
        fo = fw.fo[l]
        y = x @ fw.wx_t[l]                                    # (2n, fo)
        nxt_skip = (l + 1) in fw.skip_in
        width = fo + (fw.pe_in if nxt_skip else 0)
        x_next = new(2 * n, width)
        s1, t2 = new(n, fo), new(n, fo)
        fp_ptr = None if l == 0 else ctypes.c_void_p(featpart.data_ptr() + 4 * fw.off[l - 1])
        _c(L.gens_sdf_act_fwd(P(y), fp_ptr, ldfp, P(fw.bias[l]), n, fo, 100.0, inv_sqrt2 if nxt_skip else 1.0,
                              P(x_next), width, P(s1), P(t2), st), "gens_sdf_act_fwd")
        if nxt_skip:
            _c(L.gens_copy_scaled(P(pos), fw.pe_in, 2 * n, inv_sqrt2, P(x_next), width, fo, st), "gens_copy_scaled")
        sp1.append(s1)
        sp2.append(t2)
        x = x_next
    y_last = x[:n] @ fw.wx_t[last]                            # (n,1): only the primal SDF is needed
    sdf = (y_last + featpart[:n, fw.off[last - 1]: fw.off[last - 1] + 1] + fw.bias[last]) / fw.scale

    # ---- reverse (cotangent + its tangent) ---------------------------------------------------------
    gfp = new(2 * n, ldfp)                                    # cotangents of the feature part, per layer slice
    gfp[:n, fw.off[last - 1]] = 1.0 / fw.scale
    gfp[n:, fw.off[last - 1]] = 0.0
    g = new(2 * n, fw.wx[last].shape[1])                      # [g_h; dg_h] of the last hidden layer
    g[:n] = fw.wx[last][0] / fw.scale
    g[n:] = 0.0
    g_pos = torch.zeros((2 * n, fw.pe_in), device=dev, dtype=torch.float32)
    in_scale = 1.0
    for l in range(last - 1, -1, -1):
        fo = fw.fo[l]
        if l == 0:
            ga = new(2 * n, fo)
            ga_ptr, ldga = P(ga), fo
        else:
            ga = gfp[:, fw.off[l - 1]: fw.off[l - 1] + fo]     # write straight into the feature-part slice
            ga_ptr, ldga = ctypes.c_void_p(gfp.data_ptr() + 4 * fw.off[l - 1]), ldfp
        _c(L.gens_sdf_act_bwd(P(g), g.shape[1], in_scale, P(sp1[l]), P(sp2[l]), n, fo, ga_ptr, ldga, st),
           "gens_sdf_act_bwd")
        gx = ga @ fw.wx[l]                                    # (2n, K_l)
        if l == 0:
            g_pos += gx
        elif l in fw.skip_in:
            # input of this layer was [h_{l-1}, pos] / sqrt(2)
            g_pos += gx[:, fw.fo[l - 1]:] * inv_sqrt2
            g, in_scale = gx, inv_sqrt2                       # first fo[l-1] columns, leading dim K_l
        else:
            g, in_scale = gx, 1.0
    g_fe = gfp @ fw.wf                                        # (2n, pe_feat)

    grad, smooth = new(n, 3), new(n, 3)
    g_f, dg_f = new(n, nf), new(n, nf)
    _c(L.gens_sdf_decode(P(pts), P(feats), P(dfeats), P(g_pos), P(g_fe), n, fw.scale, _U, fw.multires,
                         fw.feat_multires, nf, P(g_f), P(dg_f), P(grad), P(smooth), st), "gens_sdf_decode")
    _c(L.gens_trilinear_vjp2(P(pts), n, pyr, _U, P(g_f), P(dg_f), P(grad), P(smooth) if need_smooth else None, st),
       "gens_trilinear_vjp2")
    return sdf, grad, (smooth if need_smooth else None)


@torch.no_grad()
def value_only(net, pts: torch.Tensor, volumes, folded: Optional[FoldedSDF] = None) -> torch.Tensor:
    """SDF values (n,1) with the same fused stages, primal rows only: one K3 launch, one encode launch,
    one feature-part GEMM, then GEMM + fused bias/softplus per layer -- the evaluation the up-sampling loop
    (112 of the 240 SDF evaluations per ray) and the mesh lattice use."""
    _lib.require_cuda(pts)
    L = _lib.lib()
    fw = FoldedSDF(net) if folded is None else folded
    pts = _lib.f32c(pts.reshape(-1, 3))
    n = pts.shape[0]
    dev = pts.device
    st = _lib.stream_ptr(dev)
    vols = [volumes] if isinstance(volumes, torch.Tensor) else list(volumes)
    packed = [packed_volume(v) for v in vols]
    pyr = _lib.make_pyramid(packed, [v.shape[2] for v in vols])
    nf = 4 * len(vols)
    new = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)
    P = _lib.ptr
    feats = new(n, nf)
    _c(L.gens_trilinear_fwd(P(pts), n, pyr, P(feats), st), "gens_trilinear_fwd")
    pos, fe = new(n, fw.pe_in), new(n, fw.pe_feat)
    _c(L.gens_sdf_encode(P(pts), P(feats), None, n, fw.scale, _U, fw.multires, fw.feat_multires, nf, P(pos), P(fe), st),
       "gens_sdf_encode")
    if USE_TC and fw.tc_supported():
        from . import mlp_tc
        return mlp_tc.sdf_values(fw.packed(), pos, fe)
    featpart = fe @ fw.wf_t
    ldfp = featpart.shape[1]
    last = fw.n_layers - 1
    inv_sqrt2 = 1.0 / math.sqrt(2.0)
    x = pos
    for l in range(last):
        fo = fw.fo[l]
        y = x @ fw.wx_t[l]
        nxt_skip = (l + 1) in fw.skip_in
        width = fo + (fw.pe_in if nxt_skip else 0)
        x_next = new(n, width)
        fp_ptr = None if l == 0 else ctypes.c_void_p(featpart.data_ptr() + 4 * fw.off[l - 1])
        _c(L.gens_sdf_act_fwd(P(y), fp_ptr, ldfp, P(fw.bias[l]), n, fo, 100.0, inv_sqrt2 if nxt_skip else 1.0,
                              P(x_next), width, None, None, st), "gens_sdf_act_fwd")
        if nxt_skip:
            _c(L.gens_copy_scaled(P(pos), fw.pe_in, n, inv_sqrt2, P(x_next), width, fo, st), "gens_copy_scaled")
        x = x_next
    y_last = x @ fw.wx_t[last]
    return (y_last + featpart[:, fw.off[last - 1]: fw.off[last - 1] + 1] + fw.bias[last]) / fw.scale
