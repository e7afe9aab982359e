#!/usr/bin/env python
"""Benchmark of the diffusion-sampling hot path (BASELINE.json metric: DDIM-50 voxels/s on the 3-D 160x224x160 UNet).

    python bench.py --gpus N --steps K --warmup W            # this repo's arm (one process per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU path (oracle port) on host cores

Definitions:
  * workload  = config C3 of BASELINE.json: DiffusionModelUNet 3-D (256, 256, 512), attention (F, F, T), heads
                (0, 0, 512), 2 res blocks, x = 1 x 1 x 160 x 224 x 160 per GPU, DDIMScheduler(1000,
                "scaled_linear_beta", 0.0005, 0.0195, clip_sample=False), random-init weights (zero-init convs
                redrawn N(0, 0.02^2)), synthetic noise — SURVEY.md §8(d).
  * one STEP  = one DDIM iteration = UNet forward + scheduler.step on the per-GPU batch; all 50 iterations of a
                sample are identical work, so  value = n_gpus * voxels / (50 * seconds_per_step)  [voxels/s].
  * value     = inputs resident in HBM; e2e = the same metric through DiffusionInferer.sample() starting from pinned
                HOST noise (H2D), the per-step timestep H2D the reference API does, and the final D2H of the sample.
  * roofline  = tensor-pipe: algorithmic FLOPs of the 3x3x3-conv launches of igemm_tc_kernel<256,4> in the timed
                region / their CUDA-event time, against MEASURED_PEAKS.json's sustained bf16 GEMM throughput.
  * weak scaling: every GPU samples its own volume; the only collective is one all_gather of the finished samples.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

C3 = dict(spatial_dims=3, in_channels=1, out_channels=1, num_channels=(256, 256, 512),
          attention_levels=(False, False, True), num_head_channels=(0, 0, 512), num_res_blocks=2)
C3_SCHED = dict(num_train_timesteps=1000, schedule="scaled_linear_beta", beta_start=0.0005, beta_end=0.0195,
                clip_sample=False)
C3_VOLUME = (160, 224, 160)
CPU_VOLUME = (32, 40, 32)       # the reference tutorial's volume: the full one cannot run on CPU
DDIM_STEPS = 50
METRIC = "ddim50_voxels_per_s"


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        p = json.loads(f.read_text())
        return p.get("bf16_tflops_sustained", 1449.3), p.get("hbm_gbs", 6582.5), "measured"
    return 1400.0, 6650.0, "fallback"


def build_model_state(seed=0):
    """C3 UNet with the reference's parameter tree, zero-init convs redrawn (SURVEY.md §8d 'Synthetic inputs')."""
    import torch
    from generativemodels_b200.networks.nets import DiffusionModelUNet
    torch.manual_seed(seed)
    m = DiffusionModelUNet(**C3).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for p in m.parameters():
            if float(p.abs().max()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    return m


# ---------------------------------------------------------------------------------------------------------------
# reference / CPU arm: the oracle port of the reference's CPU path
# ---------------------------------------------------------------------------------------------------------------
def pick_cpu_threads():
    """All host cores are available to the reference arm, but oneDNN convolutions of this size often run slower when
    heavily oversubscribed across sockets; time a representative conv at a few thread counts and keep the fastest."""
    import torch
    import torch.nn.functional as F
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (ncpu, ncpu // 2, 64, 32, 16, 8) if 1 <= c <= ncpu}, reverse=True)
    x, w = torch.randn(1, 256, 16, 20, 16), torch.randn(256, 256, 3, 3, 3)
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        F.conv3d(x, w, padding=1)
        t0 = time.perf_counter()
        for _ in range(2):
            F.conv3d(x, w, padding=1)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def cpu_steps(model, steps: int, warmup: int):
    import torch
    from oracle import torch_oracle as O
    pick_cpu_threads()
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    cfg = dict(num_head_channels=C3["num_head_channels"], norm_num_groups=32, norm_eps=1e-6, with_conditioning=False)
    sched = O.DDIMOracle(**C3_SCHED)
    sched.set_timesteps(DDIM_STEPS)
    torch.manual_seed(1234)
    x = torch.randn(1, 1, *CPU_VOLUME)
    times = []
    with torch.no_grad():
        for i, t in enumerate(sched.timesteps[: warmup + steps]):
            t0 = time.perf_counter()
            eps = O.unet_forward(sd, cfg, x, torch.Tensor((t,)))
            x, _ = sched.step(eps, int(t), x)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    vox = CPU_VOLUME[0] * CPU_VOLUME[1] * CPU_VOLUME[2]
    return vox / (DDIM_STEPS * sec), sec, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model = build_model_state()
    value, sec, cores = cpu_steps(model, args.steps, args.warmup)
    sample = (f"oracle port of the reference CPU path: UNet forward + DDIM step on 1x1x{'x'.join(map(str, CPU_VOLUME))} "
              f"(the full 160x224x160 volume needs ~60 GB of fp32 activations and 2x29.9 GiB attention scores on CPU), "
              f"{args.steps} steps")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C3: 3D DiffusionModelUNet (256,256,512) DDIM-50", "volume": list(CPU_VOLUME),
                   "per_gpu_batch": 1},
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) > 3 + j and r[3 + j].lower() == "active" for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


KERNELS_PER_CALL = {"b200_groupnorm_stats": 2}


def run_b200(args):
    import torch
    import torch.distributed as dist

    from generativemodels_b200 import _lib, ops
    from generativemodels_b200.inferers import DiffusionInferer
    from generativemodels_b200.networks.schedulers import DDIMScheduler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.require_device()

    model = build_model_state()
    cpu_line = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sec, cores = cpu_steps(model, 3, 1)
        cpu_line = {"value": v, "unit": "voxels/s", "cores": cores, "kind": "port",
                    "sample": f"oracle UNet forward + DDIM step on 1x1x{'x'.join(map(str, CPU_VOLUME))}, 3 steps "
                              f"({sec:.2f} s/step)"}
    model = model.cuda()
    vol = tuple(args.volume) if args.volume else C3_VOLUME
    voxels = vol[0] * vol[1] * vol[2] * args.batch
    sched = DDIMScheduler(**C3_SCHED)
    sched.set_timesteps(DDIM_STEPS)
    torch.manual_seed(1234)
    noise_all = torch.randn(world * args.batch, 1, *vol)            # one global draw, sliced per rank
    noise_host = noise_all[rank * args.batch:(rank + 1) * args.batch].contiguous().pin_memory()
    x = noise_host.cuda(non_blocking=True)

    # ---- instrumentation: launch counter + CUDA events around the dominant kernel's launches ----
    launches = {"n": 0, "on": False}
    for name in _lib.SIGNATURES:
        if name in ("b200_last_error_string", "b200_version", "b200_device_check", "b200_sm_count", "b200_abi_sizeof",
                    "b200_groupnorm_workspace_bytes", "b200_attention_flash_workspace_bytes",
                    "b200_igemm_split_workspace_bytes"):
            continue
        fn = getattr(lib, name)

        def counted(*a, _fn=fn, _k=KERNELS_PER_CALL.get(name, 1)):
            if launches["on"]:
                launches["n"] += _k
            return _fn(*a)
        setattr(lib, name, counted)

    conv_events, conv_flops = [], [0.0]
    raw = ops.igemm_raw

    def timed_igemm(p):
        dominant = launches["on"] and p.n_seg >= 27 and p.out_cols > 128          # 3x3x3 convs -> <256,4>
        if not dominant:
            return raw(p)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        raw(p)
        e1.record()
        rows = p.out_N * p.out_D * p.out_H * p.out_W
        kval = sum(min(p.a_C[p.seg[i].src] - p.seg[i].c0 * 64, p.seg[i].nchunks * 64) for i in range(p.n_seg))
        conv_flops[0] += 2.0 * rows * p.cout * kval
        conv_events.append((e0, e1))
    ops.igemm_raw = timed_igemm

    def one_step(x, t):
        eps = model(x, timesteps=torch.Tensor((t,)).to(x.device))
        x, _ = sched.step(eps, t, x)
        return x

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    ts = [int(t) for t in sched.timesteps]
    k = 0
    for _ in range(args.warmup):
        x = one_step(x, ts[k % DDIM_STEPS]); k += 1
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches["on"] = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        x = one_step(x, ts[k % DDIM_STEPS]); k += 1
    if world > 1:                                   # the path's only collective: gather the finished samples
        gathered = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(gathered, x)
    e1.record()
    barrier()
    launches["on"] = False
    ms_total = e0.elapsed_time(e1)
    clock_line = clocks.stop() if rank == 0 else None
    if world > 1:
        tmax = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms_total = float(tmax.item())
    ms_step = ms_total / args.steps
    value = world * voxels / (DDIM_STEPS * ms_step * 1e-3)
    conv_ms = sum(a.elapsed_time(b) for a, b in conv_events)
    n_launch = launches["n"]
    finite = bool(torch.isfinite(x).all())

    # ---- e2e through the public API, host buffers in, host buffer out ----
    ops.igemm_raw = raw
    ke = args.steps
    sched_e = DDIMScheduler(**C3_SCHED)
    sched_e.set_timesteps(ke)
    inferer = DiffusionInferer(sched_e)
    out_host = torch.empty_like(noise_host).pin_memory()
    barrier()
    t0 = time.perf_counter()
    xin = noise_host.cuda(non_blocking=True)
    sample = inferer.sample(input_noise=xin, diffusion_model=model, scheduler=sched_e, verbose=False)
    out_host.copy_(sample, non_blocking=True)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tmax = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s = float(tmax.item())
    e2e_value = world * voxels / (e2e_s * DDIM_STEPS / ke)
    nbytes = noise_host.numel() * 4

    if rank == 0:
        peak_tf, _, which = peaks()
        achieved = conv_flops[0] / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "C3: 3D DiffusionModelUNet (256,256,512) attn (F,F,T) heads (0,0,512), DDIM-50; "
                                   "one step = UNet forward + DDIMScheduler.step",
                       "volume": list(vol), "per_gpu_batch": args.batch, "samples_per_s": value / (voxels / args.batch),
                       "l2": "per-step working set (tens of GB of activations) >> 126 MB L2; no explicit flush",
                       "finite_output": finite},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf if peak_tf else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel (the 256->256
                         # 3x3x3 conv at 160x224x160; algorithmic 5.88e9 B) from the ncu --set full capture in
                         # profiles/r1_ncu_igemm_conv256_fullres_v6_details.txt (4.24 GB read + 2.91 GB written,
                         # L2 hit rate 96.5 %, tensor pipe 97.6 % active in that capture)
                         "traffic": 7.14e9,
                         "kernel": "igemm_tc_kernel<256,4> (3x3x3 convolutions)", "peak_source": which + " sustained bf16",
                         "share_of_step": conv_ms / ms_total if ms_total else None,
                         "launches_timed": len(conv_events)},
            "cpu_baseline": cpu_line,
            "e2e": {"value": e2e_value, "unit": "voxels/s", "h2d_bytes_per_step": nbytes / ke + 4,
                    "d2h_bytes_per_step": nbytes / ke, "steps_run": ke,
                    "api": "DiffusionInferer.sample(pinned host noise -> cuda, DDIMScheduler) -> pinned host"},
            "gpu_launches": n_launch,
            "clocks": clock_line,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1, help="samples per GPU")
    ap.add_argument("--volume", type=int, nargs=3, default=None, help="override the C3 volume (debug only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
