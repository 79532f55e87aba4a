"""Dumps torch.linalg.inv_ex results on the GPU for random camera-to-world matrices (GPU box only), so that
the LU variant ATen/cuBLAS uses can be identified offline (tools/fit_inverse.py)."""
import sys
sys.path.insert(0, '.')
import numpy as np
import torch

dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(5)
n = 20000
# random rotations (QR of gaussian) with translations of a few units, plus the synthetic scene's own cameras
q, r = torch.linalg.qr(torch.randn(n, 3, 3, generator=g, dtype=torch.float64))
q = q * torch.sign(torch.diagonal(r, dim1=1, dim2=2))[:, None, :]
q[torch.det(q) < 0, :, 0] *= -1
c2w = torch.zeros(n, 4, 4, dtype=torch.float64)
c2w[:, :3, :3] = q
c2w[:, :3, 3] = torch.randn(n, 3, generator=g, dtype=torch.float64) * 2
c2w[:, 3, 3] = 1
# half of them: small rotations (DTU-like), so pivoting rarely permutes
small = torch.linalg.matrix_exp(torch.cross(torch.eye(3, dtype=torch.float64)[None].expand(n // 2, 3, 3),
                                            (torch.randn(n // 2, 3, generator=g, dtype=torch.float64) * 0.2)[:, None, :].expand(-1, 3, -1), dim=-1))
c2w[: n // 2, :3, :3] = small
c2w = c2w.float()
from gens_b200.synthetic import make_scene
extra = torch.cat([make_scene(480, 640, nv, seed=s, with_images=False).c2ws for nv in (3, 5) for s in (0, 1)], 0)
c2w = torch.cat([extra, c2w], 0).contiguous()
outs = {}
for bs in (3, 5, c2w.shape[0]):   # the batch size may select a different algorithm
    res = []
    for a in range(0, c2w.shape[0] if bs > 5 else 3000, bs):
        res.append(torch.linalg.inv_ex(c2w[a:a + bs].to(dev), check_errors=False)[0].cpu())
    outs[f"inv_bs{bs}"] = torch.cat(res, 0).numpy()
outs["inv_torch_inverse"] = torch.inverse(c2w[:3000].to(dev)).cpu().numpy()
np.savez_compressed('gpurun_out/inverse_dump.npz', c2w=c2w.numpy(), **outs)
print({k: v.shape for k, v in outs.items()})
print("bs3 == bs5 on common:", np.array_equal(outs["inv_bs3"][:2995], outs["inv_bs5"][:2995]),
      "bs3 == full:", np.array_equal(outs["inv_bs3"], outs[f"inv_bs{c2w.shape[0]}"][:outs["inv_bs3"].shape[0]]),
      "inverse == inv_ex:", np.array_equal(outs["inv_torch_inverse"], outs[f"inv_bs{c2w.shape[0]}"][:3000]))
