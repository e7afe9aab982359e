"""Dev probe for ncu: warm the network, then ONE eager forward between cudaProfilerStart/Stop.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches_brain.csv python tools/one_forward.py brain
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from generativemodels_b200.networks.nets import DiffusionModelUNet

which = sys.argv[1] if len(sys.argv) > 1 else "brain"
torch.manual_seed(0)
if which == "brain":
    net = DiffusionModelUNet(spatial_dims=3, in_channels=7, out_channels=3, num_channels=(256, 512, 768),
                             num_res_blocks=2, attention_levels=(False, True, True), norm_num_groups=32, norm_eps=1e-6,
                             resblock_updown=True, num_head_channels=(0, 512, 768), with_conditioning=True,
                             transformer_num_layers=1, cross_attention_dim=4).cuda().eval()
    x, ctx = torch.randn(1, 7, 20, 28, 20).cuda(), torch.randn(1, 1, 4).cuda()
elif which == "c3":          # the bench's 3-D UNet at the full 160 x 224 x 160 volume
    net = DiffusionModelUNet(spatial_dims=3, in_channels=1, out_channels=1, num_channels=(256, 256, 512),
                             attention_levels=(False, False, True), num_head_channels=(0, 0, 512),
                             num_res_blocks=2).cuda().eval()
    x, ctx = torch.randn(1, 1, 160, 224, 160).cuda(), None
elif which == "c5":          # one guided step's UNet forward on the doubled batch (ControlNet residuals omitted)
    net = DiffusionModelUNet(2, 3, 3, num_res_blocks=1, num_channels=(128, 256, 256),
                             attention_levels=(False, True, True), num_head_channels=256, with_conditioning=True,
                             cross_attention_dim=1).cuda().eval()
    x, ctx = torch.randn(2, 3, 256, 256).cuda(), torch.tensor([[[-1.0]], [[1.0]]]).cuda()
else:                        # c2 (batch 1) / c2n32 (batch 32)
    net = DiffusionModelUNet(2, 3, 3, num_res_blocks=2, num_channels=(128, 256, 512),
                             attention_levels=(False, True, True), num_head_channels=(0, 256, 512)).cuda().eval()
    x, ctx = torch.randn(32 if which == "c2n32" else 1, 3, 64, 64).cuda(), None
with torch.no_grad():
    for p in net.parameters():
        if float(p.abs().max()) == 0:
            p.normal_(0, 0.02)
t = torch.tensor([500]).cuda()
for _ in range(3):
    net(x, timesteps=t, context=ctx)
torch.cuda.synchronize()
# the implicit-GEMM problems of the profiled forward, in launch order (joined with the ncu list by tools/join_shapes.py)
import json
from generativemodels_b200 import _lib
lib = _lib.require_device()
shapes, raw = [], lib.b200_igemm


def logged(pp, stream):
    p = pp._obj
    shapes.append(dict(rows=p.out_N * p.out_D * p.out_H * p.out_W, cout=p.cout,
                       chunks=sum(p.seg[i].nchunks for i in range(p.n_seg)), n_seg=p.n_seg,
                       split=bool(p.split_ws), stride=p.stride_w, out_f32=p.out_dtype != 0))
    return raw(pp, stream)


lib.b200_igemm = logged
torch.cuda.profiler.start()
net(x, timesteps=t, context=ctx)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
lib.b200_igemm = raw
Path("gpurun_out").mkdir(exist_ok=True)
Path(f"gpurun_out/shapes_{which}.json").write_text(json.dumps(shapes))
print("done", len(shapes), "igemm calls")
