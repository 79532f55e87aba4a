import sys; sys.path.insert(0,'.')
import numpy as np, torch
sys.path.insert(0,'tests'); from test_volume_gpu import _run_k1
from gens_b200 import _lib, build
from gens_b200.synthetic import make_scene
from gens_b200.volume import stage_cameras
from oracle import c_oracle
build.build(); _lib.lib()
sc = make_scene(240, 320, 3, seed=3, n_scales=6)
for div_mode in (0,1):
  for i,d in enumerate([128,64]):
    w2c,k = stage_cameras(sc.intrs, sc.c2ws, i)
    ovol, omsk, oix, oiy, ovm = c_oracle.volume_agg(sc.features[i].numpy(), w2c.numpy(), k.numpy(), torch.linspace(-1,1,d).numpy(), div_mode=div_mode, debug=True)
    vol, msk, ix0, iy0, valid = _run_k1(sc.features[i], w2c, k, d, div_mode)
    neq = vol != ovol
    print('div',div_mode,'D',d,'mismatch per channel',neq.reshape(8,-1).sum(1), 'of', d**3, 'maxabs', np.abs(vol-ovol).max())
    cnt = ovm.sum(0)
    for c in range(1,4):
        sel = cnt==c
        print('   cnt',c,'voxels',sel.sum(),'mean-mismatch',neq[0][sel].sum(),'var-mismatch',neq[4][sel].sum())
    idx = np.argwhere(neq[0])
    if len(idx):
        a,b,c = idx[0]; print('   first', idx[0], vol[0,a,b,c], ovol[0,a,b,c], 'cnt', cnt[a,b,c], 'views', ovm[:,a,b,c])
