"""Times the compiled-in K1 launch variants on the config-2 scales (GPU box only)."""
import sys
sys.path.insert(0, '.')
import torch
from gens_b200 import _lib, build
from gens_b200.synthetic import make_scene
from gens_b200.volume import agg_scale_into, pack_feature_maps, stage_camera_slots, stage_cameras

build.build(); L = _lib.lib()
dev = torch.device('cuda:0')
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 3
# '0c' = variant 0 with the cameras staged in the constant bank (gens_stage_cameras), '0' = shared-memory cameras
variants = sys.argv[2].split(',') if len(sys.argv) > 2 else ['10', '11', '12c', '13', '0', '0c']
sc = make_scene(480, 640, nv, seed=0, with_images=False).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ref = {}
for vname in variants:
    const_cams = vname.endswith('c')
    variant = int(vname.rstrip('c'))
    L.gens_debug_set_variant(variant)
    line = [f'variant {vname}:']
    tot = 0.0
    for i, d in enumerate([256, 128, 64]):
        feat = pack_feature_maps(sc.features[i])
        h, w = sc.features[i].shape[-2:]
        w2c, k = stage_cameras(sc.intrs, sc.c2ws, i)
        grid = torch.linspace(-1, 1, d, device=dev)
        vol = torch.empty((8, d, d, d), device=dev); msk = torch.empty((d, d, d), device=dev)
        slot = stage_camera_slots(w2c, k, [1.0])[0] if const_cams else 0
        def run():
            agg_scale_into(feat, (h, w), w2c, k, 1.0, grid, d, vol, msk, None, 1, 1, slot)
        for _ in range(3): run()
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort(); t = ts[len(ts) // 2]
        tot += t
        key = (i,)
        if key in ref:
            same = torch.equal(ref[key][0], vol) and torch.equal(ref[key][1], msk)
        else:
            ref[key] = (vol.clone(), msk.clone()); same = True
        gbs = (d ** 3 * 36 + nv * 16 * h * w) / t / 1e6
        line.append(f'D{d}: {t*1e3:7.1f} us ({gbs:6.0f} GB/s){"" if same else " MISMATCH"}')
    line.append(f'sum {tot*1e3:.1f} us')
    print(' '.join(line), flush=True)
L.gens_debug_set_variant(0)
