"""Per-phase cycle counts of the reverse tensor-core SDF kernel (block 0), via gens_debug_tc_profile (GPU box only)."""
import sys
sys.path.insert(0, '.')
import torch
from gens_b200 import _lib
from gens_b200.config import gens_model_conf
from gens_b200.implicit_surface import ImplicitSurface
from gens_b200.synthetic import make_reg_volumes
dev = torch.device('cuda:0')
torch.manual_seed(0)
surf = ImplicitSurface(gens_model_conf(perturb=0.0)["implicit_surface"]).to(dev)
vols = [v.to(dev) for v in make_reg_volumes([64, 32, 16, 8, 4], seed=3)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
pts = torch.rand(n, 3, device=dev) * 2 - 1
buf = torch.zeros(16, dtype=torch.int64, device=dev)
L = _lib.lib()
with torch.no_grad():
    for _ in range(2):
        surf.sdf_network.value_grad_smooth_nograd(pts, vols)
    torch.cuda.synchronize()
    # forward kernels first: [0] staging, [1] MMA wait, [2] epilogue work, [3] tiles, [7] epilogues, [8-11] issuer
    for target, name, fn in ((0, "value kernel", lambda: surf.sdf_network.sdf_nograd(pts, vols)),
                             (1, "JVP forward kernel", lambda: surf.sdf_network.value_grad_smooth_nograd(pts, vols))):
        buf.zero_()
        L.gens_debug_tc_profile(_lib.ptr(buf), target)
        fn()
        torch.cuda.synchronize()
        q = buf.cpu().tolist()
        L.gens_debug_tc_profile(None, 2)
        per_tile = q[12] / max(q[3], 1)
        print(f"{name}: block 0 {q[12]} cycles, {q[3]} tiles ({per_tile:.0f} per tile), {q[7]} layer epilogues")
        print(f"  staging the encodings {q[0] / max(q[3], 1):8.0f} per tile ({100 * q[0] / max(q[12], 1):.1f} %), of which "
              f"proxy fence + arrive {q[4] / max(q[3], 1):.0f}")
        print(f"  waiting for MMAs      {q[1] / max(q[7], 1):8.0f} per layer ({100 * q[1] / max(q[12], 1):.1f} %)")
        print(f"  epilogue work         {q[2] / max(q[7], 1):8.0f} per layer ({100 * q[2] / max(q[12], 1):.1f} %)")
        mt = q[8] + q[9] + q[10]
        print(f"  MMA issuer: wait A / input {100 * q[8] / max(mt, 1):.1f} %, wait weights {100 * q[9] / max(mt, 1):.1f} %, "
              f"issue {100 * q[10] / max(mt, 1):.1f} % ({q[10] / max(q[11], 1):.0f} cycles per k-step, {q[11]} k-steps)")
    buf.zero_()
    L.gens_debug_tc_profile(_lib.ptr(buf), 2)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); surf.sdf_network.value_grad_smooth_nograd(pts, vols); b.record()
    torch.cuda.synchronize()
    L.gens_debug_tc_profile(None, 2)
p = buf.cpu().tolist()
names = ["wait x-part MMAs", "wait s1/t2 (cp.async)", "acc load + arithmetic (first half)", "wait feature-part MMAs",
         "stores + second half + arrive", "prefetch issue", "-", "layers"]
tot = sum(p[:6])
print(f"{n} points, whole sweep {a.elapsed_time(b):.2f} ms; reverse kernel block 0: {p[12]} cycles, {p[7]} layer epilogues")
names[6] = 'tile tails (last MMA wait + result stores)'
tot = sum(p[:7])
for i in (0, 1, 2, 3, 4, 5, 6):
    print(f"  epilogue thread 0: {names[i]:36s} {p[i]:12d} cycles  {100 * p[i] / max(tot, 1):5.1f} %   {p[i] / max(p[7], 1):8.0f} per layer")
mt = p[8] + p[9] + p[10]
for i, nm in ((8, "wait A operand"), (9, "wait weights"), (10, "issue 6 MMAs + commits")):
    print(f"  MMA issuer:        {nm:36s} {p[i]:12d} cycles  {100 * p[i] / max(mt, 1):5.1f} %   {p[i] / max(p[11], 1):8.0f} per k-step")
