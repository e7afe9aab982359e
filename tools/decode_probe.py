"""Dev probe: where a transformer decode step spends its time (graph replay vs sampling ops vs eager step)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from generativemodels_b200.networks.nets import DecoderOnlyTransformer

tr = DecoderOnlyTransformer(num_tokens=257, max_seq_len=1025, attn_layers_dim=512, attn_layers_depth=12,
                            attn_layers_heads=8).cuda().eval()
for N in (1, 8):
    cache = tr.new_cache(N, torch.device("cuda"), None, graph=True)
    tok = torch.full((N, 1), 256).cuda()
    for _ in range(4):
        lg = tr.step(tok, cache)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 400
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        cache.length += 1
        cache.graph.replay()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n
    print(f"N={N}: graph replay {e0.elapsed_time(e1) / n * 1e3:.0f} us GPU, {wall * 1e6:.0f} us wall per token (prefix ~{cache.length})")
    t0 = time.perf_counter()
    for _ in range(200):
        logits = lg[:, -1, :] / 1.0
        probs = torch.nn.functional.softmax(logits, dim=-1)
        probs[:, 256] = 0
        idx = torch.multinomial(probs, num_samples=1)
        tok = torch.cat((tok, idx), dim=1)
    torch.cuda.synchronize()
    print(f"N={N}: sampling ops {(time.perf_counter() - t0) / 200 * 1e6:.0f} us wall per token")
    ec = tr.new_cache(N, torch.device("cuda"), None, graph=False)
    tok = torch.full((N, 1), 256).cuda()
    tr.step(tok, ec)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(100):
        tr.step(tok, ec)
    torch.cuda.synchronize()
    print(f"N={N}: eager step {(time.perf_counter() - t0) / 100 * 1e6:.0f} us wall per token")
