"""Dev probe: the brain-LDM bundle at its published size — latent 3x20x28x20 -> 160x224x160,
UNet (256,512,768) with cross-attention, AutoencoderKL (64,128,128,128), DDIM-50 — through
generativemodels_b200.bundle.Sampler on one GPU, random-init weights.  Prints the 50-step loop and the decode
separately, eager vs CUDA graph, the NiftiSaver time, and cold-start with / without the packed-weight cache file."""
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from generativemodels_b200.bundle import NiftiSaver, Sampler, load_packed, save_packed
from generativemodels_b200.networks.nets import AutoencoderKL, DiffusionModelUNet
from generativemodels_b200.networks.schedulers import DDIMScheduler

AE_KW = dict(spatial_dims=3, in_channels=1, out_channels=1, latent_channels=3, num_channels=(64, 128, 128, 128),
             num_res_blocks=2, norm_num_groups=32, norm_eps=1e-06, attention_levels=(False,) * 4,
             with_encoder_nonlocal_attn=False, with_decoder_nonlocal_attn=False)
UNET_KW = dict(spatial_dims=3, in_channels=7, out_channels=3, num_channels=(256, 512, 768), num_res_blocks=2,
               attention_levels=(False, True, True), norm_num_groups=32, norm_eps=1e-06, resblock_updown=True,
               num_head_channels=(0, 512, 768), with_conditioning=True, transformer_num_layers=1,
               cross_attention_dim=4, upcast_attention=True, use_flash_attention=False)


def redraw(m):
    with torch.no_grad():
        for p in m.parameters():
            if float(p.detach().abs().max()) == 0:
                p.normal_(0, 0.02)
    return m


def wall(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, out


torch.manual_seed(0)
ae = AutoencoderKL(**AE_KW).cuda().eval()
unet = redraw(DiffusionModelUNet(**UNET_KW)).cuda().eval()
print(f"params: UNet {sum(p.numel() for p in unet.parameters())/1e6:.1f} M, "
      f"AutoencoderKL {sum(p.numel() for p in ae.parameters())/1e6:.1f} M")
sched = DDIMScheduler(beta_start=0.0015, beta_end=0.0205, num_train_timesteps=1000, schedule="scaled_linear_beta",
                      clip_sample=False)
sched.set_timesteps(num_inference_steps=50)
noise = torch.randn(1, 3, 20, 28, 20).cuda()
cond = torch.tensor([[0.0, 0.1, 0.2, 0.4]]).cuda().unsqueeze(1)

t_cold, _ = wall(lambda: unet(torch.cat([noise, noise[:, :1].expand(-1, 4, -1, -1, -1)], 1),
                              timesteps=torch.tensor([500]).cuda(), context=cond))
print(f"first UNet forward (packs all weights): {t_cold*1e3:.1f} ms")
with tempfile.TemporaryDirectory() as d:
    t_save, n = wall(lambda: save_packed(unet, d + "/unet.packed.pt"))
    fresh = DiffusionModelUNet(**UNET_KW).cuda().eval()
    fresh.load_state_dict(unet.state_dict())
    t_load, ok = wall(lambda: load_packed(fresh, d + "/unet.packed.pt"))
    t_first, _ = wall(lambda: fresh(torch.cat([noise, noise[:, :1].expand(-1, 4, -1, -1, -1)], 1),
                                    timesteps=torch.tensor([500]).cuda(), context=cond))
    print(f"packed cache: {n} entries, {Path(d, 'unet.packed.pt').stat().st_size/1e6:.0f} MB, save {t_save*1e3:.0f} ms, "
          f"load {t_load*1e3:.0f} ms (ok={ok}), first forward after load {t_first*1e3:.1f} ms")
    del fresh

for graph in (False, True):
    smp = Sampler(use_cuda_graph=graph)
    smp.sampling_fn(noise, ae, unet, sched, cond)                       # warm-up (+ graph capture)
    t, out = wall(lambda: smp.sampling_fn(noise, ae, unet, sched, cond))
    print(f"sampling_fn DDIM-50 + decode ({'CUDA graph' if graph else 'eager'}): {t*1e3:.1f} ms -> "
          f"{tuple(out.shape)}, {out.numel()/t/1e6:.2f} Mvoxel/s")
z = torch.randn(1, 3, 20, 28, 20).cuda()
ae.decode_stage_2_outputs(z)
t_dec, out = wall(lambda: ae.decode_stage_2_outputs(z))
print(f"decode_stage_2_outputs alone: {t_dec*1e3:.1f} ms ({15.2/t_dec:.0f} TFLOP/s on SURVEY's 15.2 TFLOP)")
with tempfile.TemporaryDirectory() as d:
    sv = NiftiSaver(d)
    sv.save(out, "warm")
    t_sv, _ = wall(lambda: sv.save(out, "vol"))
    print(f"NiftiSaver.save: {t_sv*1e3:.1f} ms ({Path(d, 'vol.nii.gz').stat().st_size/1e6:.2f} MB)")
print(f"peak memory {torch.cuda.max_memory_allocated()/2**30:.2f} GiB")
