import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
from gens_b200 import projector, _lib, build
from gens_b200.synthetic import make_scene
build.build()
g=np.load('tests/golden/render.npz')
DEV='cuda:0'
scene = make_scene(96, 128, 3, seed=11).to(DEV)
pts = torch.from_numpy(g["sdf_pts"]).to(DEV)
projector.ATEN_CUDA_FLAVOUR = 0
fv, rd, mk = projector.lookup_feature(pts, scene.imgs, scene.intrs, scene.c2ws, scene.features)
ref=torch.from_numpy(g["lf_feat"]).to(DEV)
d=(fv-ref).abs()
print('per-channel max diff', d.amax(dim=(0,1)).cpu().numpy().round(5))
print('per-view max diff', d.amax(dim=(0,2)).cpu().numpy())
bad=(d>1e-4).any(-1)
print('bad rows', bad.sum().item(), 'of', bad.numel(), 'mask true', mk.sum().item())
i=torch.nonzero(bad)[0]; print(i, fv[i[0],i[1]].cpu().numpy().round(4), ref[i[0],i[1]].cpu().numpy().round(4), mk[i[0],i[1]].item())
