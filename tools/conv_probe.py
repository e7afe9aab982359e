"""Dev probe: the two dominant C3 convolutions (3x3x3, 160x224x160, 256 -> 256 and 512 (virtual concat) -> 256) timed with
CUDA events, plus a checksum so that variants (B200_IGEMM_PAIR=0/1) can be compared for equality."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from generativemodels_b200 import ops

torch.manual_seed(0)
D, H, W = 160, 224, 160
a = ops.CL((torch.randn(1, D, H, W, 256, device="cuda") * 0.5).to(ops.H16), 256, 3)
b = ops.CL((torch.randn(1, D, H, W, 256, device="cuda") * 0.5).to(ops.H16), 256, 3)
w1 = torch.randn(256, 256, 3, 3, 3, device="cuda") * 0.02
w2 = torch.randn(256, 512, 3, 3, 3, device="cuda") * 0.02
bias = torch.randn(256, device="cuda") * 0.1
pc1 = ops.PackedConv(w1, bias, 1, 1)
pc2 = ops.PackedConv(w2, bias, 1, 1, splits=[256, 256])
for name, fn, flop in (("256->256", lambda: ops.conv(a, pc1), 2.0 * D * H * W * 256 * 256 * 27),
                       ("512->256 (two sources)", lambda: ops.conv([a, b], pc2, residual=a), 2.0 * D * H * W * 512 * 256 * 27)):
    out = fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    chk = out.t.float().double()
    print(f"conv {name}: ms {' '.join(f'{t:.2f}' for t in ts)} -> best {flop/min(ts)/1e9:.0f} TFLOP/s; checksum {chk.sum().item():.6e} {chk.abs().sum().item():.6e}"
          f" gn {None if out.gn is None else out.gn.double().sum().item():.6e}")
