"""Dev probe: graph-replayed sampling with one switch of generativemodels_b200.ops off / on in ONE process on ONE box
(box-to-box variance is a few percent, the same order as the effects): brain-LDM latent UNet (3-D, 20x28x20) and the C2
latent UNet (2-D 64x64), batch 1, DDIM-50 loops without the decoder.

    python tools/splitk_ab.py [_SPLIT_K | _GN_SMALL]      (default _SPLIT_K)
"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from generativemodels_b200 import ops
from generativemodels_b200.cuda_graph import graphed
from generativemodels_b200.networks.nets import DiffusionModelUNet
from generativemodels_b200.networks.schedulers import DDIMScheduler

FLAG = sys.argv[1] if len(sys.argv) > 1 else "_SPLIT_K"
assert hasattr(ops, FLAG), FLAG
torch.manual_seed(0)


def redraw(m):
    with torch.no_grad():
        for p in m.parameters():
            if float(p.detach().abs().max()) == 0:
                p.normal_(0, 0.02)
    return m


brain = redraw(DiffusionModelUNet(spatial_dims=3, in_channels=7, out_channels=3, num_channels=(256, 512, 768),
                                  num_res_blocks=2, attention_levels=(False, True, True), norm_num_groups=32,
                                  norm_eps=1e-6, resblock_updown=True, num_head_channels=(0, 512, 768),
                                  with_conditioning=True, transformer_num_layers=1,
                                  cross_attention_dim=4)).cuda().eval()
c2 = redraw(DiffusionModelUNet(2, 3, 3, num_res_blocks=2, num_channels=(128, 256, 512),
                               attention_levels=(False, True, True), num_head_channels=(0, 256, 512))).cuda().eval()
c5 = redraw(DiffusionModelUNet(2, 3, 3, num_res_blocks=1, num_channels=(128, 256, 256),
                               attention_levels=(False, True, True), num_head_channels=256, with_conditioning=True,
                               cross_attention_dim=1)).cuda().eval()
sched = DDIMScheduler(1000, "linear_beta", beta_start=0.0015, beta_end=0.0195)
sched.set_timesteps(50)
cases = {"brain-LDM UNet 7x20x28x20": (brain, torch.randn(1, 7, 20, 28, 20).cuda(), torch.randn(1, 1, 4).cuda(), 3),
         "C2 UNet 3x64x64": (c2, torch.randn(1, 3, 64, 64).cuda(), None, 3),
         "C2 UNet 32x3x64x64": (c2, torch.randn(32, 3, 64, 64).cuda(), None, 3),
         "C5 UNet 2x3x256x256 (CFG batch, no ControlNet residuals)": (c5, torch.randn(2, 3, 256, 256).cuda(),
                                                                     torch.tensor([[[-1.0]], [[1.0]]]).cuda(), 3)}


def loop(net, x, ctx, keep):
    img = x[:, :keep]
    for t in sched.timesteps:
        out = net(x, timesteps=torch.tensor([int(t)], device=x.device), context=ctx)
        img, _ = sched.step(out, t, img)
        x[:, :keep] = img
    return img


for name, (net, x, ctx, keep) in cases.items():
    res = {}
    for rep in range(2):
        for split in (False, True):
            setattr(ops, FLAG, split)
            g = graphed(net)
            loop(g, x.clone(), ctx, keep)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                loop(g, x.clone(), ctx, keep)
            torch.cuda.synchronize()
            res.setdefault(split, []).append((time.perf_counter() - t0) / 3 / 50 * 1e3)
    print(f"{name}: ms/step {FLAG}=False {min(res[False]):.3f} (runs {res[False]}), {FLAG}=True {min(res[True]):.3f} "
          f"(runs {res[True]})")
