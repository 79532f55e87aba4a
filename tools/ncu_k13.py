"""The K13 launches of one whole-volume RegNetwork forward at the config-2 sizes (for ncu -k regex:conv3d|deconv3d|norm_relu)."""
import sys
sys.path.insert(0, '.')
import torch
from gens_b200.reg_network import RegNetwork
dev = torch.device('cuda:0')
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
net = RegNetwork().to(dev).eval()
vols = [torch.randn(1, 8, d, d, d, device=dev) for d in (256, 128, 64, 32, 16)]
with torch.no_grad():
    for _ in range(2):
        net(vols)
torch.cuda.synchronize()
