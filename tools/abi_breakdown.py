"""Dev probe: where one network forward spends its GPU time, per C-ABI entry point (CUDA events around every call of
the library, eager mode, warm) — and for b200_igemm per problem class.  Used for the latency-bound configurations
(brain-LDM latent UNet, C2 latent UNet at batch 1) where no single kernel dominates.

    python tools/abi_breakdown.py brain|c2
"""
import sys
from collections import defaultdict
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from generativemodels_b200 import _lib, ops
from generativemodels_b200.networks.nets import DiffusionModelUNet

which = sys.argv[1] if len(sys.argv) > 1 else "brain"
torch.manual_seed(0)
if which == "brain":
    net = DiffusionModelUNet(spatial_dims=3, in_channels=7, out_channels=3, num_channels=(256, 512, 768),
                             num_res_blocks=2, attention_levels=(False, True, True), norm_num_groups=32, norm_eps=1e-6,
                             resblock_updown=True, num_head_channels=(0, 512, 768), with_conditioning=True,
                             transformer_num_layers=1, cross_attention_dim=4).cuda().eval()
    x, ctx = torch.randn(1, 7, 20, 28, 20).cuda(), torch.randn(1, 1, 4).cuda()
else:
    net = DiffusionModelUNet(2, 3, 3, num_res_blocks=2, num_channels=(128, 256, 512),
                             attention_levels=(False, True, True), num_head_channels=(0, 256, 512)).cuda().eval()
    x, ctx = torch.randn(1, 3, 64, 64).cuda(), None
with torch.no_grad():
    for p in net.parameters():
        if float(p.abs().max()) == 0:
            p.normal_(0, 0.02)
t = torch.tensor([500]).cuda()
lib = _lib.require_device()
for _ in range(3):
    net(x, timesteps=t, context=ctx)
torch.cuda.synchronize()

events = []
skip = ("b200_last_error_string", "b200_version", "b200_device_check", "b200_sm_count", "b200_abi_sizeof",
        "b200_groupnorm_workspace_bytes", "b200_attention_flash_workspace_bytes", "b200_igemm_split_workspace_bytes")
for name in _lib.SIGNATURES:
    if name in skip:
        continue
    fn = getattr(lib, name)

    def timed(*a, _fn=fn, _name=name):
        label = _name
        if _name == "b200_igemm":
            p = a[0]._obj
            rows = p.out_N * p.out_D * p.out_H * p.out_W
            k = sum(p.seg[i].nchunks for i in range(p.n_seg)) * 64
            label = f"b200_igemm rows={rows} cout={p.cout} K={k}{' split' if p.split_ws else ''}"
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = _fn(*a)
        e1.record()
        events.append((label, e0, e1))
        return rc
    setattr(lib, name, timed)

w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
w0.record()
net(x, timesteps=t, context=ctx)
w1.record()
torch.cuda.synchronize()
tot, cnt = defaultdict(float), defaultdict(int)
for label, e0, e1 in events:
    tot[label] += e0.elapsed_time(e1)
    cnt[label] += 1
wall = w0.elapsed_time(w1)
inside = sum(tot.values())
print(f"{which}: one eager forward {wall:.2f} ms on the stream, {inside:.2f} ms inside {len(events)} library calls "
      f"(split-K launches so far: {ops._SPLIT_LAUNCHES})")
by_fn = defaultdict(float)
for label, v in tot.items():
    by_fn[label.split()[0]] += v
for name, v in sorted(by_fn.items(), key=lambda kv: -kv[1]):
    print(f"  {v:8.3f} ms  {name}")
print("  -- b200_igemm by problem --")
for label, v in sorted(((k, v) for k, v in tot.items() if k.startswith("b200_igemm ")), key=lambda kv: -kv[1])[:25]:
    print(f"  {v:8.3f} ms  x{cnt[label]:<3d} {label[11:]}")
