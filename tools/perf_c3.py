"""Dev probe (not the bench): one C3 UNet forward at a chosen volume with a per-operator time breakdown."""
import argparse
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import collections
import time

import torch

from generativemodels_b200 import ops
from generativemodels_b200.networks.nets import DiffusionModelUNet

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="160,224,160")
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--breakdown", type=int, default=1)
ap.add_argument("--gnfuse", type=int, default=1)
a = ap.parse_args()
shape = tuple(int(v) for v in a.shape.split(","))
ops._GN_FUSE = bool(a.gnfuse)

torch.manual_seed(0)
m = DiffusionModelUNet(spatial_dims=3, in_channels=1, out_channels=1, num_channels=(256, 256, 512),
                       attention_levels=(False, False, True), num_head_channels=(0, 0, 512), num_res_blocks=2).cuda().eval()
for p in m.parameters():
    if float(p.abs().max()) == 0:
        p.data.normal_(0, 0.02)
x = torch.randn(1, 1, *shape).cuda()
t = torch.Tensor((500,)).cuda()

acc = collections.defaultdict(float)
cnt = collections.defaultdict(int)


def wrap(name):
    fn = getattr(ops, name)

    def timed(*args, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*args, **kw)
        e1.record()
        e1.synchronize()
        acc[name] += e0.elapsed_time(e1)
        cnt[name] += 1
        return r
    return timed


torch.cuda.synchronize()
t0 = time.time()
y = m(x, t)
torch.cuda.synchronize()
print(f"first forward (incl. weight packing): {time.time() - t0:.2f} s, out {tuple(y.shape)} finite={bool(torch.isfinite(y).all())}")
print(f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
for i in range(a.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = m(x, t)
    e1.record()
    torch.cuda.synchronize()
    print(f"forward {i}: {e0.elapsed_time(e1):.1f} ms")
if a.breakdown:
    names = ["conv", "groupnorm", "attention", "linear", "linear_transposed", "upsample_nearest2x", "to_cl", "from_cl_f32",
             "small_linear", "axpy"]
    import generativemodels_b200.networks.nets.diffusion_model_unet as U
    import generativemodels_b200.networks._holders as H
    saved = {n: getattr(ops, n) for n in names}
    for n in names:
        setattr(ops, n, wrap(n))
    y = m(x, t)
    torch.cuda.synchronize()
    for n in names:
        setattr(ops, n, saved[n])
    tot = sum(acc.values())
    for n in sorted(acc, key=lambda k: -acc[k]):
        print(f"  {n:20s} {acc[n]:9.1f} ms  x{cnt[n]:4d}  {100 * acc[n] / tot:5.1f}%")
    print(f"  total {tot:.1f} ms")

if a.breakdown:
    # per-launch view of the implicit-GEMM kernel: executed FLOPs (64-channel chunks as issued) and rate
    groups = collections.OrderedDict()
    raw = ops.igemm_raw

    def timed_raw(p):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        raw(p)
        e1.record()
        e1.synchronize()
        M = p.out_N * p.out_D * p.out_H * p.out_W
        K = 64 * sum(p.seg[i].nchunks for i in range(p.n_seg))
        key = (p.out_N, p.out_D, p.out_H, p.out_W, p.cout, p.n_seg, K, p.w_batched, p.out_dtype, bool(p.res_ptr), bool(p.stat_ptr))
        g = groups.setdefault(key, [0, 0.0, 2.0 * M * p.cout * K])
        g[0] += 1
        g[1] += e0.elapsed_time(e1)

    ops.igemm_raw = timed_raw
    y = m(x, t)
    torch.cuda.synchronize()
    ops.igemm_raw = raw
    print("  igemm launches grouped by shape: N,OD,OH,OW,cout,n_seg,K,batched,odt,res,stat | count | ms total | TFLOP/s")
    for k, (c, ms, fl) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
        print(f"  {str(k):64s} x{c:3d} {ms:8.2f} ms  {fl * c / ms / 1e9:8.1f} TF/s")
    print(f"  igemm total {sum(g[1] for g in groups.values()):.1f} ms")
