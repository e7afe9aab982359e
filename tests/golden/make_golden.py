"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference, CPU fp32, MONAI shim).

Run in the build container:  python -m tests.golden.make_golden
Each fixture holds the reference module's state_dict, the seeded inputs and the reference's outputs, so that both the
CPU oracle (tests/test_oracle_golden.py) and the CUDA path (tests/test_parity_gpu.py) can be checked against the real
reference on a box that does not have it.  Models are tiny on purpose (fixtures total < 1 MB).
"""
from pathlib import Path

import torch

from oracle import ref_import
from tests.golden import configs as G

OUT = Path(__file__).resolve().parent

GOLDEN_UNETS = {
    "g_unet2d": (dict(spatial_dims=2, in_channels=1, out_channels=1, num_res_blocks=1, num_channels=(8, 16),
                      attention_levels=(False, True), norm_num_groups=4, num_head_channels=8), (2, 1, 8, 8), None),
    "g_unet3d_cross": (dict(spatial_dims=3, in_channels=2, out_channels=2, num_res_blocks=1, num_channels=(8, 16),
                            attention_levels=(False, True), norm_num_groups=4, num_head_channels=(0, 8),
                            with_conditioning=True, cross_attention_dim=4), (1, 2, 4, 8, 8), (1, 2, 4)),
}
GOLDEN_CONTROLNET = dict(spatial_dims=2, in_channels=2, num_res_blocks=1, num_channels=(8, 16),
                         attention_levels=(False, True), norm_num_groups=4, num_head_channels=(0, 8),
                         with_conditioning=True, cross_attention_dim=4, conditioning_embedding_in_channels=1,
                         conditioning_embedding_num_channels=(8,))
GOLDEN_AEKL = dict(spatial_dims=2, in_channels=1, out_channels=1, num_channels=(8, 16), latent_channels=3,
                   num_res_blocks=(1, 1), norm_num_groups=4, attention_levels=(False, True),
                   with_encoder_nonlocal_attn=True, with_decoder_nonlocal_attn=True)
GOLDEN_VQVAE = dict(spatial_dims=3, in_channels=1, out_channels=1, num_channels=(8, 8), num_res_channels=8,
                    num_res_layers=1, downsample_parameters=((2, 4, 1, 1),) * 2,
                    upsample_parameters=((2, 4, 1, 1, 0),) * 2, num_embeddings=16, embedding_dim=8)


def main():
    ref_import.import_reference()
    import generative.networks.nets as nets
    import generative.networks.schedulers as sch
    from generative.inferers import DiffusionInferer, LatentDiffusionInferer

    torch.set_grad_enabled(False)
    for name, (kw, shape, ctx_shape) in GOLDEN_UNETS.items():
        torch.manual_seed(0)
        m = G.randomize_zero_params(nets.DiffusionModelUNet(**kw)).eval()
        torch.manual_seed(11)
        x = torch.randn(shape)
        t = torch.randint(0, 1000, (shape[0],)).long()
        ctx = torch.randn(ctx_shape) if ctx_shape else None
        y = m(x, t, context=ctx)
        fx = dict(kwargs=kw, state_dict=m.state_dict(), x=x, t=t, context=ctx, y=y)
        if name == "g_unet2d":       # DDIM-5 and PNDM-6 trajectories through DiffusionInferer.sample
            skw = dict(num_train_timesteps=1000, schedule="linear_beta", beta_start=0.0015, beta_end=0.0195)
            s = sch.DDIMScheduler(**skw)
            s.set_timesteps(5)
            torch.manual_seed(12)
            noise = torch.randn(1, 1, 8, 8)
            fx.update(ddim_kwargs=skw, ddim_steps=5, noise=noise,
                      ddim_sample=DiffusionInferer(s).sample(input_noise=noise, diffusion_model=m, scheduler=s,
                                                             verbose=False))
            pk = dict(num_train_timesteps=1000, skip_prk_steps=True)
            p = sch.PNDMScheduler(**pk)
            p.set_timesteps(6)
            fx.update(pndm_kwargs=pk, pndm_steps=6,
                      pndm_sample=DiffusionInferer(p).sample(input_noise=noise, diffusion_model=m, scheduler=p,
                                                             verbose=False))
            dk = dict(num_train_timesteps=1000)
            d = sch.DDPMScheduler(**dk)
            d.set_timesteps(4)
            torch.manual_seed(13)       # DDPM draws CPU noise from the global generator (ddpm.py:245-247)
            fx.update(ddpm_kwargs=dk, ddpm_steps=4, ddpm_seed=13,
                      ddpm_sample=DiffusionInferer(d).sample(input_noise=noise, diffusion_model=m, scheduler=d,
                                                             verbose=False))
        torch.save(fx, OUT / f"{name}.pt")

    kw = GOLDEN_CONTROLNET
    torch.manual_seed(0)
    cn = G.randomize_zero_params(nets.ControlNet(**kw)).eval()
    ukw = {k: v for k, v in kw.items() if not k.startswith("conditioning_embedding")}
    un = G.randomize_zero_params(nets.DiffusionModelUNet(out_channels=2, **ukw)).eval()
    torch.manual_seed(14)
    x, cond, ctx = torch.randn(2, 2, 8, 8), torch.rand(2, 1, 8, 8), torch.randn(2, 1, 4)
    t = torch.tensor([10, 500]).long()
    down, mid = cn(x, t, cond, conditioning_scale=0.7, context=ctx)
    y = un(x, t, context=ctx, down_block_additional_residuals=down, mid_block_additional_residual=mid)
    torch.save(dict(kwargs=kw, unet_kwargs=dict(out_channels=2, **ukw), cn_state_dict=cn.state_dict(),
                    unet_state_dict=un.state_dict(), x=x, cond=cond, context=ctx, t=t, scale=0.7, down=down, mid=mid,
                    y=y), OUT / "g_controlnet.pt")

    kw = GOLDEN_AEKL
    torch.manual_seed(0)
    ae = nets.AutoencoderKL(**kw).eval()
    torch.manual_seed(15)
    x = torch.randn(2, 1, 16, 16)
    mu, sigma = ae.encode(x)
    rec = ae.decode(mu)
    # LatentDiffusionInferer.sample: latent DDIM-3 then decode
    lkw = dict(spatial_dims=2, in_channels=3, out_channels=3, num_res_blocks=1, num_channels=(8, 16),
               attention_levels=(False, True), norm_num_groups=4, num_head_channels=8)
    lun = G.randomize_zero_params(nets.DiffusionModelUNet(**lkw)).eval()
    skw = dict(num_train_timesteps=1000, schedule="linear_beta", beta_start=0.0015, beta_end=0.0195)
    s = sch.DDIMScheduler(**skw)
    s.set_timesteps(3)
    torch.manual_seed(16)
    lat_noise = torch.randn(1, 3, 8, 8)
    ldm = LatentDiffusionInferer(s, scale_factor=0.8).sample(input_noise=lat_noise, autoencoder_model=ae,
                                                             diffusion_model=lun, scheduler=s, verbose=False)
    torch.save(dict(kwargs=kw, state_dict=ae.state_dict(), x=x, mu=mu, sigma=sigma, rec=rec, latent_unet_kwargs=lkw,
                    latent_unet_state_dict=lun.state_dict(), ddim_kwargs=skw, ddim_steps=3, scale_factor=0.8,
                    latent_noise=lat_noise, ldm_sample=ldm), OUT / "g_aekl2d.pt")

    kw = GOLDEN_VQVAE
    torch.manual_seed(0)
    vq = nets.VQVAE(**kw).eval()
    torch.manual_seed(17)
    x = torch.rand(1, 1, 16, 16, 16)
    rec, loss = vq(x)
    z = vq.encode(x)
    idx = vq.index_quantize(x)
    torch.save(dict(kwargs=kw, state_dict=vq.state_dict(), x=x, z=z, idx=idx, rec=rec, loss=loss,
                    perplexity=vq.quantizer.perplexity.clone(), dec_from_idx=vq.decode_samples(idx)),
               OUT / "g_vqvae3d.pt")

    # reference's own known-answer test for the quantiser (tests/test_vector_quantizer.py:45-62): nearest-code
    # assignment of inputs equal to codebook rows
    from generative.networks.layers.vector_quantizer import EMAQuantizer
    torch.manual_seed(0)
    q = EMAQuantizer(spatial_dims=2, num_embeddings=2, embedding_dim=2, epsilon=0, decay=0).eval()
    original_weight_0 = q.embedding.weight[0].clone()
    original_weight_1 = q.embedding.weight[1].clone()
    x_0 = original_weight_0[None, :, None, None] + 0.001
    x_1 = original_weight_1[None, :, None, None]
    xq = torch.cat([x_0.expand(1, 2, 1, 1), x_1.expand(1, 2, 1, 1)], dim=0)
    _, _, idxq = q(xq)
    torch.save(dict(codebook=q.embedding.weight.detach().clone(), x=xq, idx=idxq), OUT / "g_vq_ema_case.pt")
    for f in sorted(OUT.glob("*.pt")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
