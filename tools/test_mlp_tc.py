"""GPU check of the tcgen05 SDF value kernel against the fp32 cuBLAS path (GPU box only)."""
import sys, time
sys.path.insert(0, '.')
import torch
from gens_b200 import _lib, sdf_analytic, mlp_tc
from gens_b200.config import gens_model_conf
from gens_b200.implicit_surface import ImplicitSurface
from gens_b200.sdf_analytic import FoldedSDF
from gens_b200.projector import packed_volume

dev = torch.device('cuda:0')
torch.manual_seed(0)
surf = ImplicitSurface(gens_model_conf(perturb=0.0)["implicit_surface"]).to(dev)
net = surf.sdf_network
g = torch.Generator(device=dev).manual_seed(1)
dims = [64, 32, 16, 8, 4]
vols = [torch.randn(1, 4, d, d, d, device=dev, generator=g) * 0.5 for d in dims]
fw = FoldedSDF(net)
packed = mlp_tc.PackedSDF(fw)
print('ksteps', packed.n_ksteps, 'stream MB', packed.wstream.numel() * 4 / 1e6, flush=True)
L = _lib.lib()
for n in [128, 1000, 128 * 148 * 3 + 77, 1 << 21]:
    pts = torch.rand(n, 3, device=dev, generator=g) * 2 - 1
    sdf_analytic.USE_TC = False
    ref = sdf_analytic.value_only(net, pts, vols, fw)
    # encodings exactly as value_only computes them
    pv = [packed_volume(v) for v in vols]
    pyr = _lib.make_pyramid(pv, dims)
    feats = torch.empty(n, 20, device=dev)
    _lib.check(L.gens_trilinear_fwd(_lib.ptr(pts), n, pyr, _lib.ptr(feats), _lib.stream_ptr(dev)), 'tri')
    pos, fe = torch.empty(n, fw.pe_in, device=dev), torch.empty(n, fw.pe_feat, device=dev)
    _lib.check(L.gens_sdf_encode(_lib.ptr(pts), _lib.ptr(feats), None, n, fw.scale, sdf_analytic._U, fw.multires,
                                 fw.feat_multires, 20, _lib.ptr(pos), _lib.ptr(fe), _lib.stream_ptr(dev)), 'enc')
    torch.cuda.synchronize()
    out = mlp_tc.sdf_values(packed, pos, fe)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    print(f'n={n}: max abs err {err:.3e}  (ref range {ref.min().item():.3f}..{ref.max().item():.3f}) nan={bool(out.isnan().any())}', flush=True)
    if n == 1000:
        ex = mlp_tc.emulate(packed, pos, fe).to(dev)   # float64 replay of the same network
        print(f'   vs float64: tcgen05 {float((out - ex).abs().max()):.3e}   fp32 cuBLAS path {float((ref - ex).abs().max()):.3e}', flush=True)
    if n >= 1 << 20:
        for fn, name in ((lambda: mlp_tc.sdf_values(packed, pos, fe), 'tcgen05 MLP only'),
                         (lambda: sdf_analytic.value_only(net, pts, vols, fw), 'fp32 value_only (incl. lookup+encode)')):
            for _ in range(2): fn()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(5): fn()
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
            print(f'  {name}: {dt*1e3:.2f} ms  ({n/dt/1e6:.1f} Mpts/s)', flush=True)

# ---- value + gradient + second-order term: tensor-core sweep vs the fp32 cuBLAS sweep ----------------------
def cmp(name, a, b):
    err = (a - b).abs()
    tol = 1e-5 * max(1.0, float(b.abs().max())) + 1e-4 * b.abs()
    bad = int((err > tol).sum())
    print(f'   {name}: max abs err {float(err.max()):.3e} (scale {float(b.abs().max()):.3f}), beyond 1e-4 tolerance: {bad}/{b.numel()}', flush=True)

for n in [64, 1000, 64 * 148 * 2 + 13, 1 << 20]:
    pts = torch.rand(n, 3, device=dev, generator=g) * 2 - 1
    sdf_analytic.USE_TC = False
    r = sdf_analytic.value_grad_smooth(net, pts, vols, fw)
    sdf_analytic.USE_TC = True
    o = sdf_analytic.value_grad_smooth(net, pts, vols, fw)
    torch.cuda.synchronize()
    print(f'n={n}: nan={any(bool(t.isnan().any()) for t in o)}', flush=True)
    for nm, a, b in zip(('sdf', 'grad', 'smooth'), o, r):
        cmp(nm, a, b)
    if n >= 1 << 20:
        for flag, name in ((True, 'tcgen05 sweep'), (False, 'fp32 cuBLAS sweep')):
            sdf_analytic.USE_TC = flag
            fn = lambda: sdf_analytic.value_grad_smooth(net, pts, vols, fw)
            for _ in range(2): fn()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(3): fn()
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
            print(f'  {name}: {dt*1e3:.2f} ms  ({n/dt/1e6:.1f} Mpts/s)', flush=True)
sdf_analytic.USE_TC = True
