"""Dev probe: short-reduction GEMM-shaped calls of b200_igemm (transformer linears of the latent / conditioned UNets:
K = 256..1024, one to fourteen output tiles per SM), graph-replayed timing per shape, optionally one warm call between
cudaProfilerStart/Stop for ncu.

    python tools/gemm_probe.py                     # timing table
    ncu --profile-from-start off --set full --import-source on -o gpurun_out/prof_gemm python tools/gemm_probe.py --profile 32768,2048,256
"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from generativemodels_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--split-ab", action="store_true")
ap.add_argument("--profile", default=None, help="M,N,K[,res]: one profiled call instead of the timing table")
args = ap.parse_args()
torch.manual_seed(0)
dev = torch.device("cuda", 0)


def make(M, N, K, res):
    x = ops.CL(torch.randn(1, 1, 1, M, K, device=dev).to(ops.H16), K, 2)
    pl = ops.PackedLinear(torch.randn(N, K, device=dev) / K ** 0.5, torch.randn(N, device=dev))
    r = ops.CL(torch.randn(1, 1, 1, M, N, device=dev).to(ops.H16), N, 2) if res else None
    return x, pl, r


if args.profile:
    f = [int(v) for v in args.profile.split(",")]
    x, pl, r = make(f[0], f[1], f[2], len(f) > 3 and f[3])
    for _ in range(3):
        ops.linear(x, pl, residual=r)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ops.linear(x, pl, residual=r)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)



def timed(fn, reps=20):
    for _ in range(3):
        fn()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay()
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * reps)


if args.split_ab:
    # the split-K decision per shape: same call with ops._SPLIT_K off / on (GEMM + reduction kernel), graph-replayed
    for (M, N, K) in [(8192, 256, 2304), (8192, 256, 1024), (8192, 256, 4608), (8192, 512, 4608), (4096, 128, 1152),
                      (4096, 128, 2304), (1024, 256, 2304), (1024, 256, 1024), (256, 512, 4608), (256, 512, 2048),
                      (1400, 512, 13824), (175, 768, 20736), (2048, 256, 2304), (2048, 512, 2304), (512, 512, 4608)]:
        x, pl, r = make(M, N, K, 1)
        res = {}
        for flag in (False, True):
            ops._SPLIT_K = flag
            n0 = ops._SPLIT_LAUNCHES
            res[flag] = timed(lambda: ops.linear(x, pl, residual=r))
            took = ops._SPLIT_LAUNCHES > n0
        print(f"M={M:<6d} N={N:<5d} K={K:<6d} one pass {res[False]:7.2f} us   split {res[True]:7.2f} us "
              f"({'split taken' if took else 'not split'})  {2.0 * M * N * K / min(res.values()) / 1e6:7.1f} TFLOP/s best")
    sys.exit(0)

shapes = [(32768, 2048, 256, 0), (32768, 256, 256, 0), (32768, 256, 256, 1), (32768, 256, 1024, 1), (32768, 512, 256, 0),
          (8192, 256, 256, 0), (8192, 256, 256, 1), (8192, 2048, 256, 0), (8192, 256, 1024, 1), (8192, 512, 256, 0),
          (131072, 128, 256, 0), (131072, 128, 1152, 0), (32768, 256, 2304, 0), (1024, 256, 256, 0), (256, 512, 512, 0)]
for (M, N, K, res) in shapes:
    x, pl, r = make(M, N, K, res)
    for _ in range(3):
        ops.linear(x, pl, residual=r)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ops.linear(x, pl, residual=r)
        with torch.cuda.graph(g, stream=s):
            for _ in range(20):
                ops.linear(x, pl, residual=r)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.replay()
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 100
    fl = 2.0 * M * N * K
    by = 2.0 * (M * K + N * K + M * N * (2 if res else 1))
    print(f"M={M:<7d} N={N:<5d} K={K:<5d} res={res}  {us:8.2f} us  {fl / us / 1e6:8.1f} TFLOP/s  {by / us / 1e3:8.1f} GB/s")
