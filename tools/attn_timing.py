"""Dev probe: softmax-warp cycle breakdown of the flash kernel (needs the -DFA_TIMING build, B200_DEV_LIB=...)."""
import math, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from generativemodels_b200 import _lib
if os.environ.get("B200_DEV_LIB"):
    _lib.LIB_PATH = Path(os.environ["B200_DEV_LIB"]).resolve()
from generativemodels_b200 import ops
ops._KEEP_FLASH_WS = True
T = S = 89600
dh = 512
torch.manual_seed(0)
q = (torch.randn(1, T, dh, device="cuda") * 0.5).to(ops.H16)
k = (torch.randn(1, S, dh, device="cuda") * 0.5).to(ops.H16)
vt = torch.randn(1, dh, S, device="cuda").to(ops.H16)
for i in range(2):
    o = ops.attention(q, k, None, 1, dh, 1 / math.sqrt(dh), vt=vt)
    torch.cuda.synchronize()
ws = ops._LAST_FLASH_WS
n_kv = (S + 63) // 64
grid = 148
slab = grid * 128 * n_kv * 64 * 2
fac = ws[slab: slab + grid * 4 * n_kv * 32 * 4].view(torch.float32)
c = fac[-8:].cpu().tolist()
names = ["loop/other", "wait s_full", "tmem ld S + arrive", "max + decision", "wait p_empty", "exp2 + pack", "st P + slab + arrive", "-"]
tot = sum(c)
print(f"softmax warp (CTA 0, warp 2, first item, {n_kv} blocks): total {tot:.0f} cycles = {tot / n_kv:.0f} per block")
for n, v in zip(names, c):
    print(f"  {n:24s} {v / n_kv:8.1f} cyc/block  {100 * v / tot:5.1f}%")
