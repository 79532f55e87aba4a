"""Tensor-core (tcgen05, 3xTF32) value pass of the SDF MLP -- host side of csrc/sdf_mlp_tc.cu.

The network of reference models/modules/sdf_network.py:98-123 is re-expressed as a stream of "k-steps":
one k-step = 16 input channels of one layer, i.e. an (N_l x 16) weight block that the kernel multiplies
with a (128 x 16) slice of one of three resident A operands (two tcgen05 K = 8 instructions per term),

    F  the 100-channel volume-feature encoding (shared memory, same for every layer >= 1),
    P  the 27-channel position encoding        (shared memory; layer 0 and the skip layer),
    H  the previous layer's activations        (tensor memory),

and accumulates into that layer's fp32 accumulator.  Per layer the F and P k-steps come first (they do not
depend on the previous layer, so they overlap its epilogue), the H k-steps last.

Packed format (what gens_sdf_mlp_value_tc expects):
  wstream  float32: per k-step [hi block | lo block]; a block is the (N_l x 16) weights as [4][N_l][4]
           (four 16-byte K chunks, rows contiguous inside a chunk: the canonical K-major no-swizzle UMMA
           layout with LBO = 16 N_l bytes, SBO = 128 bytes); hi = weights rounded to TF32, lo = w - hi.
  ksteps   uint32 (n,4): byte offset, byte count (128 N_l), A source | index << 8 | accumulator column << 16,
           flags (1 overwrite the accumulator, 2 commit after this k-step [16: to barrier 1], 4 wait for the
           epilogue's A operand) | N_l << 16.
  bias     float32 (n_layers, 128), zero padded.
Fan-outs are padded to multiples of 16 with zero rows (101 -> 112, the single SDF row -> 16), fan-ins to
multiples of 16 with zero columns; the 1/sqrt(2) of the skip connection is folded into that layer's weights.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import torch

from . import _lib

A_F, A_P, A_H = 0, 1, 2
ACC0, ACC1 = 256, 384  # TMEM columns of the two accumulators (csrc/sdf_mlp_tc.cu)
F_K, P_K = 112, 32  # padded widths of the resident encodings (csrc/sdf_mlp_tc.cu kFChunks / kPChunks)


def _split_tf32(w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    hi = ((w.contiguous().view(torch.int32) + 0x1000) & -8192).view(torch.float32)
    return hi, w - hi


def _blocks(w: torch.Tensor, n_pad: int) -> torch.Tensor:
    """(fo, k) weights -> (k_pad/16, 2 [hi, lo], 4 chunks, n_pad rows, 4) blocks."""
    fo, k = w.shape
    k_pad = (k + 15) // 16 * 16
    full = w.new_zeros((n_pad, k_pad))
    full[:fo, :k] = w
    hi, lo = _split_tf32(full)
    tiles = [t.reshape(n_pad, k_pad // 16, 4, 4).permute(1, 2, 0, 3) for t in (hi, lo)]
    return torch.stack(tiles, dim=1).contiguous()


class PackedSDF:
    """The folded SDF network in the kernel's streaming format (built once per weight version)."""

    def __init__(self, fw):
        """`fw` is a sdf_analytic.FoldedSDF."""
        if fw.pe_in > P_K or fw.pe_feat > F_K:
            raise RuntimeError("gens_b200 tensor-core SDF kernel: encodings wider than the resident operands")
        dev = fw.wx[0].device
        last = fw.n_layers - 1
        inv_sqrt2 = 1.0 / math.sqrt(2.0)
        chunks: List[torch.Tensor] = []
        steps: List[List[int]] = []
        off = 0
        bias = torch.zeros((fw.n_layers, 128), device=dev, dtype=torch.float32)
        for l in range(fw.n_layers):
            fo = fw.fo[l]
            n_pad = (fo + 15) // 16 * 16
            if n_pad > 128:
                raise RuntimeError("gens_b200 tensor-core SDF kernel: layers wider than 128 are not supported")
            bias[l, :fo] = fw.bias[l]
            segs = []
            wx = fw.wx[l]
            if l >= 1:
                segs.append((A_F, fw.wf[fw.off[l - 1]: fw.off[l - 1] + fo]))
            if l == 0:
                segs.append((A_P, wx))
            elif l in fw.skip_in:
                h_in = fw.fo[l - 1]
                segs.append((A_P, wx[:, h_in:] * inv_sqrt2))
                segs.append((A_H, wx[:, :h_in] * inv_sqrt2))
            else:
                segs.append((A_H, wx))
            first = True
            for si, (kind, w) in enumerate(segs):
                blk = _blocks(w.float(), n_pad)
                nb = blk.shape[0]
                nbytes = n_pad * 128
                for j in range(nb):
                    flags = n_pad << 16
                    if first:
                        flags |= 1
                        first = False
                    if kind == A_H and j == 0:
                        flags |= 4
                    if si == len(segs) - 1 and j == nb - 1:
                        flags |= 2 | (16 if l & 1 else 0)
                    steps.append([off, nbytes, kind | (j << 8) | ((ACC1 if l & 1 else ACC0) << 16), flags])
                    off += nbytes
                chunks.append(blk.reshape(-1))
        self.wstream = torch.cat(chunks).contiguous()
        assert self.wstream.numel() * 4 == off
        self.ksteps = torch.tensor(steps, dtype=torch.int64, device="cpu").to(torch.int32).to(dev).contiguous()
        self.n_ksteps = len(steps)
        self.bias = bias.contiguous()
        self.n_layers = fw.n_layers
        self.scale = float(fw.scale)
        self.n_sm = torch.cuda.get_device_properties(dev).multi_processor_count if dev.type == "cuda" else 0


class PackedSDFReverse:
    """The transposed network for the reverse sweep (gens_sdf_mlp_rev_tc): per hidden layer l = L-1..0 the
    x-part W_x^T (cotangent of the layer input, accumulator 0, overwritten per layer) and, for l >= 1, the
    feature part W_f^T (cotangent of the feature encoding, accumulator 1, summed over the layers)."""

    def __init__(self, fw):
        dev = fw.wx[0].device
        last = fw.n_layers - 1
        inv_sqrt2 = 1.0 / math.sqrt(2.0)
        chunks: List[torch.Tensor] = []
        steps: List[List[int]] = []
        off = 0
        self.skip_layer, self.skip_col = 0, 0
        for l in range(last - 1, -1, -1):
            wx = fw.wx[l].float()
            if l in fw.skip_in:
                wx = wx * inv_sqrt2
                self.skip_layer, self.skip_col = l, fw.fo[l - 1]
            segs = [(ACC0, wx.t().contiguous())]
            if l >= 1:
                segs.append((ACC1, fw.wf[fw.off[l - 1]: fw.off[l - 1] + fw.fo[l]].float().t().contiguous()))
            for si, (acc, w) in enumerate(segs):
                n_pad = (w.shape[0] + 15) // 16 * 16
                if n_pad > 128:
                    raise RuntimeError("gens_b200 tensor-core SDF kernel: layer inputs wider than 128 are not supported")
                blk = _blocks(w, n_pad)
                nb = blk.shape[0]
                nbytes = n_pad * 128
                for j in range(nb):
                    flags = n_pad << 16
                    if j == 0 and (acc == ACC0 or l == last - 1):
                        flags |= 1
                    if j == 0 and si == 0:
                        flags |= 4
                    if j == nb - 1:
                        # x-part done -> barrier 0: the epilogue of the next layer may read accumulator 0 while the
                        # feature-part MMAs below still run; feature part done -> barrier 1: the A operand in tensor
                        # memory is no longer being read and may be overwritten
                        flags |= 2 | (16 if si == 1 else 0)
                    steps.append([off, nbytes, A_H | (j << 8) | (acc << 16), flags])
                    off += nbytes
                chunks.append(blk.reshape(-1))
        self.wstream = torch.cat(chunks).contiguous()
        self.ksteps = torch.tensor(steps, dtype=torch.int64, device="cpu").to(torch.int32).to(dev).contiguous()
        self.n_ksteps = len(steps)
        consts = torch.zeros((2, 128), device=dev, dtype=torch.float32)
        consts[0, : fw.wx[last].shape[1]] = fw.wx[last][0] / fw.scale
        consts[1, : fw.pe_feat] = fw.wf[fw.off[last - 1]] / fw.scale
        self.consts = consts.contiguous()
        self.n_hidden = last
        self.n_sm = torch.cuda.get_device_properties(dev).multi_processor_count if dev.type == "cuda" else 0


def sdf_values(packed: PackedSDF, pos: torch.Tensor, fe: torch.Tensor) -> torch.Tensor:
    """(n,27) position encoding, (n,100) feature encoding -> (n,1) SDF values."""
    _lib.require_cuda(pos, fe)
    n = pos.shape[0]
    out = torch.empty((n, 1), device=pos.device, dtype=torch.float32)
    if n == 0:
        return out
    _lib.check(_lib.lib().gens_sdf_mlp_value_tc(
        _lib.ptr(pos), _lib.ptr(fe), n, _lib.ptr(packed.wstream), _lib.ptr(packed.ksteps), packed.n_ksteps,
        _lib.ptr(packed.bias), packed.n_layers, packed.scale, packed.n_sm, _lib.ptr(out),
        _lib.stream_ptr(pos.device)), "gens_sdf_mlp_value_tc")
    return out


TAPE_TILE_POINTS = 64        # points per tile of the JVP forward / reverse kernels
TAPE_BLOCK_FLOATS = 2 * 32 * 64 * 4   # one (tile, layer) block: [sp' | sp'' da][chunk 32][slot 64] float4 = 64 KB


def sdf_jvp(packed: PackedSDF, pos: torch.Tensor, fe: torch.Tensor, n: int):
    """pos (2n,27) / fe (2n,100) with tangent rows in [n,2n) -> (sdf (n,1), tape).  The tape holds sp'(a) and
    sp''(a) da of every hidden channel for the reverse sweep, as ceil(n/64) x (n_layers-1) blocks of 64 KB laid out
    exactly as the reverse kernel stages them in shared memory (csrc/sdf_mlp_tc.cu: tape_slot)."""
    _lib.require_cuda(pos, fe)
    dev = pos.device
    sdf = torch.empty((n, 1), device=dev, dtype=torch.float32)
    tiles = (n + TAPE_TILE_POINTS - 1) // TAPE_TILE_POINTS
    tape = torch.empty((tiles, packed.n_layers - 1, TAPE_BLOCK_FLOATS), device=dev, dtype=torch.float32)
    if n == 0:
        return sdf, tape
    _lib.check(_lib.lib().gens_sdf_mlp_jvp_tc(
        _lib.ptr(pos), _lib.ptr(fe), n, _lib.ptr(packed.wstream), _lib.ptr(packed.ksteps), packed.n_ksteps,
        _lib.ptr(packed.bias), packed.n_layers, packed.scale, packed.n_sm, _lib.ptr(sdf), _lib.ptr(tape),
        _lib.stream_ptr(dev)), "gens_sdf_mlp_jvp_tc")
    return sdf, tape


def sdf_reverse(rev: PackedSDFReverse, tape: torch.Tensor, n: int):
    """Reverse sweep -> (g_pos (2n,27), g_fe (2n,100)), primal rows first."""
    dev = tape.device
    g_pos = torch.empty((2 * n, 27), device=dev, dtype=torch.float32)
    g_fe = torch.empty((2 * n, 100), device=dev, dtype=torch.float32)
    if n == 0:
        return g_pos, g_fe
    if tape.shape[1] != rev.n_hidden:
        raise RuntimeError("tape / reverse network mismatch")
    _lib.check(_lib.lib().gens_sdf_mlp_rev_tc(
        _lib.ptr(tape), n, _lib.ptr(rev.wstream), _lib.ptr(rev.ksteps), rev.n_ksteps, _lib.ptr(rev.consts),
        rev.n_hidden, rev.skip_layer, rev.skip_col, rev.n_sm, _lib.ptr(g_pos), _lib.ptr(g_fe), _lib.stream_ptr(dev)),
        "gens_sdf_mlp_rev_tc")
    return g_pos, g_fe


def _replay(ksteps, stream, srcs, n_rows, on_commit):
    """Walks a k-step stream in float64; `on_commit(acc)` is called at every commit with the accumulators
    ({column: tensor}) and may replace srcs[A_H]."""
    acc = {}
    for off, nbytes, a, flags in ksteps:
        n_pad = (flags >> 16) & 0x1ff
        assert nbytes == n_pad * 128 and off % 16 == 0
        blk = stream[off // 4: off // 4 + nbytes // 4].reshape(2, 4, n_pad, 4)
        w = (blk[0] + blk[1]).permute(1, 0, 2).reshape(n_pad, 16)         # hi + lo, (N, 16)
        kind, j, col = a & 0xff, (a >> 8) & 0xff, (a >> 16) & 0xfff
        if flags & 1:
            acc[col] = torch.zeros((n_rows, 128), dtype=torch.float64)
        acc[col][:, :n_pad] += srcs[kind][:, 16 * j: 16 * j + 16] @ w.t()
        if flags & 2:
            on_commit(acc, col, 1 if flags & 16 else 0)


def _sp(y):
    t = y * 100.0
    return torch.where(t > 20.0, y, torch.log1p(torch.exp(t.clamp(max=20.0))) / 100.0)


def emulate_grad(packed: PackedSDF, rev: PackedSDFReverse, pos: torch.Tensor, fe: torch.Tensor, n: int):
    """Float64 replay of the JVP forward and the reverse sweep (tests only).  pos (2n,27), fe (2n,100) ->
    (sdf (n,1), g_pos (2n,27), g_fe (2n,100))."""
    rows = 2 * n
    srcs = {A_F: torch.zeros((rows, F_K), dtype=torch.float64), A_P: torch.zeros((rows, P_K), dtype=torch.float64),
            A_H: torch.zeros((rows, 128), dtype=torch.float64)}
    srcs[A_F][:, :fe.shape[1]] = fe.double().cpu()
    srcs[A_P][:, :pos.shape[1]] = pos.double().cpu()
    bias = packed.bias.double().cpu()
    state = {"layer": 0, "s1": [], "t2": [], "sdf": None}

    def fwd_commit(acc, col, _barrier):
        l = state["layer"]
        y = acc[col]
        if l + 1 < packed.n_layers:
            a = y[:n] + bias[l]
            da = y[n:]
            sig = torch.where(a * 100.0 > 20.0, torch.ones_like(a), torch.sigmoid(a * 100.0))
            s2 = torch.where(a * 100.0 > 20.0, torch.zeros_like(a), 100.0 * sig * (1.0 - sig))
            state["s1"].append(sig)
            state["t2"].append(s2 * da)
            srcs[A_H] = torch.cat([_sp(a), sig * da], 0)
        else:
            state["sdf"] = ((y[:n, :1] + bias[l, :1]) / packed.scale).float()
        state["layer"] = l + 1

    _replay(packed.ksteps.cpu().tolist(), packed.wstream.double().cpu(), srcs, rows, fwd_commit)

    consts = rev.consts.double().cpu()
    rstate = {"layer": rev.n_hidden - 1, "g_pos": torch.zeros((rows, 27), dtype=torch.float64), "g_fe": None}

    def load_a(l, g):
        s1, t2 = state["s1"][l], state["t2"][l]
        return torch.cat([s1 * g[:n], t2 * g[:n] + s1 * g[n:]], 0)

    g_top = torch.zeros((rows, 128), dtype=torch.float64)
    g_top[:n] = consts[0]
    rsrcs = {A_H: load_a(rev.n_hidden - 1, g_top)}

    def rev_commit(acc, col, barrier):
        l = rstate["layer"]          # the layer whose MMAs are being committed
        if barrier == 0:             # x-part done: the next A operand can be COMPUTED from accumulator 0 ...
            g = acc[ACC0]
            if l == rev.skip_layer and l > 0:
                rstate["g_pos"] += g[:, rev.skip_col: rev.skip_col + 27]
            if l > 0:
                rstate["pending_a"] = load_a(l - 1, g)
            else:                    # layer 0 has no feature part: the tile is finished
                rstate["g_pos"] += g[:, :27]
                gfe = acc[ACC1][:, :100].clone()
                gfe[:n] += consts[1, :100]
                rstate["g_fe"] = gfe
                rstate["layer"] = l - 1
        else:                        # ... but only WRITTEN once the feature-part MMAs have read the current one
            rsrcs[A_H] = rstate.pop("pending_a")
            rstate["layer"] = l - 1

    _replay(rev.ksteps.cpu().tolist(), rev.wstream.double().cpu(), rsrcs, rows, rev_commit)
    return state["sdf"], rstate["g_pos"].float(), rstate["g_fe"].float()


def emulate(packed: PackedSDF, pos: torch.Tensor, fe: torch.Tensor) -> torch.Tensor:
    """Host restatement of the value kernel's k-step machine in float64 (tests only: checks the packing, not
    the tensor-core arithmetic).  Same inputs and result as sdf_values."""
    n = pos.shape[0]
    srcs = {A_F: torch.zeros((n, F_K), dtype=torch.float64), A_P: torch.zeros((n, P_K), dtype=torch.float64),
            A_H: torch.zeros((n, 128), dtype=torch.float64)}
    srcs[A_F][:, :fe.shape[1]] = fe.double().cpu()
    srcs[A_P][:, :pos.shape[1]] = pos.double().cpu()
    bias = packed.bias.double().cpu()
    state = {"layer": 0, "sdf": None}

    def commit(acc, col, _barrier):
        l = state["layer"]
        y = acc[col] + bias[l]
        if l + 1 < packed.n_layers:
            srcs[A_H] = _sp(y)
        else:
            state["sdf"] = (y[:, :1] / packed.scale).float()
        state["layer"] = l + 1

    _replay(packed.ksteps.cpu().tolist(), packed.wstream.double().cpu(), srcs, n, commit)
    return state["sdf"]
