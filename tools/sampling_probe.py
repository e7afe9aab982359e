"""Dev probe: cost of the per-token sampling ops of VQVAETransformerInferer.sample on the GPU."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

for N in (1, 8):
    lg = torch.randn(N, 1, 257, device="cuda")
    def t(fn, n=300):
        fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
    logits = lg[:, -1, :] / 1.0
    probs = torch.softmax(logits, -1)
    print(f"N={N}: div {t(lambda: lg[:, -1, :] / 1.0):.0f} us, softmax {t(lambda: torch.softmax(logits, -1)):.0f} us, "
          f"zero-bos {t(lambda: probs.__setitem__((slice(None), 256), 0)):.0f} us, "
          f"multinomial {t(lambda: torch.multinomial(probs, num_samples=1)):.0f} us, "
          f"cat {t(lambda: torch.cat((torch.zeros(N, 500, dtype=torch.long, device='cuda'), torch.zeros(N, 1, dtype=torch.long, device='cuda')), 1)):.0f} us")
