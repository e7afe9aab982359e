"""Dev probe: wall-clock of the other BASELINE.json configs (C2 latent diffusion, C4 VQVAE, C5 ControlNet + CFG) through
the public API on one GPU.  Not the graded bench (bench.py measures C3); numbers go to README/profiles."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from generativemodels_b200.inferers import ControlNetDiffusionInferer, LatentDiffusionInferer
from generativemodels_b200.networks.nets import VQVAE, AutoencoderKL, ControlNet, DiffusionModelUNet
from generativemodels_b200.networks.schedulers import DDIMScheduler


def redraw(m):
    with torch.no_grad():
        for p in m.parameters():
            if float(p.detach().abs().max()) == 0:
                p.normal_(0, 0.02)
    return m


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


torch.manual_seed(0)
# ---- C2: AutoencoderKL (128,128,256) + latent UNet (128,256,512), DDIM-50, 1x256x256 ----
ae = AutoencoderKL(2, 1, 1, num_channels=(128, 128, 256), latent_channels=3, num_res_blocks=2,
                   attention_levels=(False, False, False), with_encoder_nonlocal_attn=False,
                   with_decoder_nonlocal_attn=False).cuda().eval()
un = redraw(DiffusionModelUNet(2, 3, 3, num_res_blocks=2, num_channels=(128, 256, 512),
                               attention_levels=(False, True, True), num_head_channels=(0, 256, 512))).cuda().eval()
s = DDIMScheduler(1000, "linear_beta", beta_start=0.0015, beta_end=0.0195)
s.set_timesteps(50)
inf = LatentDiffusionInferer(s, scale_factor=1.0)
from generativemodels_b200.cuda_graph import graphed
ung = graphed(un)
for N in (1, 8, 32):
    noise = torch.randn(N, 3, 64, 64).cuda()
    t = timed(lambda: inf.sample(noise, ae, un, s, verbose=False), 2)
    print(f"C2 LDM DDIM-50 N={N}: {t*1e3:.1f} ms/call, {N/t:.2f} samples/s, {N*65536/t/1e6:.3f} Mpixel/s")
    t = timed(lambda: inf.sample(noise, ae, ung, s, verbose=False), 2)
    print(f"C2 LDM DDIM-50 N={N} (CUDA graph): {t*1e3:.1f} ms/call, {N/t:.2f} samples/s, {N*65536/t/1e6:.3f} Mpixel/s")

# ---- C4: 3-D VQVAE (256,256), K=256, D=32, 1x128^3 ----
vq = VQVAE(3, 1, 1, num_channels=(256, 256), num_res_channels=256, num_res_layers=2,
           downsample_parameters=((2, 4, 1, 1),) * 2, upsample_parameters=((2, 4, 1, 1, 0),) * 2, num_embeddings=256,
           embedding_dim=32).cuda().eval()
x = torch.rand(1, 1, 128, 128, 128).cuda()
t = timed(lambda: vq(x), 3)
print(f"C4 VQVAE encode->VQ->decode 1x128^3: {t*1e3:.1f} ms, {128**3/t/1e6:.1f} Mvoxel/s ({8.95/t:.0f} TFLOP/s algorithmic)")

# ---- C5: ControlNet-conditioned 2-D UNet 3x256x256, classifier-free guidance, DDIM-50 ----
kw = dict(spatial_dims=2, in_channels=3, num_res_blocks=1, num_channels=(128, 256, 256),
          attention_levels=(False, True, True), num_head_channels=256, with_conditioning=True, cross_attention_dim=1)
un5 = redraw(DiffusionModelUNet(out_channels=3, **kw)).cuda().eval()
cn5 = redraw(ControlNet(conditioning_embedding_in_channels=1, conditioning_embedding_num_channels=(16,), **kw)).cuda().eval()
s5 = DDIMScheduler(1000)
s5.set_timesteps(50)
yy, xx = torch.meshgrid(torch.arange(256), torch.arange(256), indexing="ij")
mask = (((xx - 128) ** 2 + (yy - 128) ** 2) < 100 ** 2).float()[None, None].cuda()


def cfg_sample(N, guidance=7.0):
    img = torch.randn(N, 3, 256, 256).cuda()
    ctx = torch.cat([-torch.ones(N, 1, 1), torch.ones(N, 1, 1)]).cuda()
    cond = mask.expand(2 * N, -1, -1, -1).contiguous()
    for t in s5.timesteps:
        x2 = torch.cat([img] * 2)
        ts = torch.Tensor((t,)).cuda()
        down, mid = cn5(x2, ts, cond, context=ctx, _internal=True)
        eps = un5(x2, ts, context=ctx, down_block_additional_residuals=down, mid_block_additional_residual=mid)
        eu, et = eps.chunk(2)
        img, _ = s5.step(eu + guidance * (et - eu), t, img)
    return img


for N in (1,):
    t = timed(lambda: cfg_sample(N), 1)
    print(f"C5 ControlNet+CFG DDIM-50 N={N}: {t*1e3:.0f} ms/guided sample, {N*3*256*256/t/1e6:.3f} Mvalues/s")

# ---- transformer sampler: 32x32 latent grid = 1024 tokens, 12 layers x 512, 8 heads ----
from generativemodels_b200.inferers import VQVAETransformerInferer
from generativemodels_b200.networks.nets import DecoderOnlyTransformer
from generativemodels_b200.utils.ordering import Ordering

vq2 = VQVAE(2, 1, 1, num_channels=(128, 128), num_res_channels=128, num_res_layers=2,
            downsample_parameters=((2, 4, 1, 1),) * 2, upsample_parameters=((2, 4, 1, 1, 0),) * 2, num_embeddings=256,
            embedding_dim=32).cuda().eval()
tr = DecoderOnlyTransformer(num_tokens=257, max_seq_len=1025, attn_layers_dim=512, attn_layers_depth=12,
                            attn_layers_heads=8).cuda().eval()
order = Ordering("raster_scan", 2, (1, 32, 32))
vinf = VQVAETransformerInferer()
for N in (1, 8):
    start = torch.full((N, 1), 256).cuda()
    t = timed(lambda: vinf.sample((32, 32), start, vq2, tr, order, verbose=False), 1)
    print(f"transformer sampler 1024 tokens N={N} (key/value cache): {t:.2f} s/call, {N*1024/t:.0f} tokens/s")


class _NoCache:          # the reference's loop: full forward over the prefix for every token
    def __init__(self, m): self.m = m; self.max_seq_len = m.max_seq_len
    def __call__(self, x, context=None): return self.m(x, context=context)


start = torch.full((1, 1), 256).cuda()
t = timed(lambda: vinf.sample((16, 16), start, vq2, _NoCache(tr), Ordering("raster_scan", 2, (1, 16, 16)), verbose=False), 1)
t2 = timed(lambda: vinf.sample((16, 16), start, vq2, tr, Ordering("raster_scan", 2, (1, 16, 16)), verbose=False), 1)
print(f"transformer sampler 256 tokens N=1: prefix recompute per token {t:.2f} s, key/value cache {t2:.2f} s")
