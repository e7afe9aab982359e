"""The reference's own unit-test configurations for the hot-path networks (tests/test_diffusion_model_unet.py:23-232,
tests/test_autoencoderkl.py, tests/test_vqvae.py, tests/test_controlnet.py, tests/test_diffusion_inferer.py,
tests/test_latent_diffusion_inferer.py — channel widths 4-8, spatial 8-32, head dims 2-8) run through the CUDA path.
The reference only asserts output shapes; here every case is additionally compared with the CPU oracle."""
import pytest
import torch

from oracle import torch_oracle as O
from tests.golden import configs as G

pytestmark = pytest.mark.gpu
TOL = 3e-2


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def nets():
    import generativemodels_b200.networks.nets as N
    return N


B2 = dict(spatial_dims=2, in_channels=1, out_channels=1, num_res_blocks=1, num_channels=(8, 8, 8), norm_num_groups=8)
B3 = dict(B2, spatial_dims=3)
UNCOND = [
    dict(B2, attention_levels=(False, False, False)),
    dict(B2, attention_levels=(False, False, False), num_res_blocks=(1, 1, 2)),
    dict(B2, attention_levels=(False, False, False), resblock_updown=True),
    dict(B2, attention_levels=(False, False, True), num_head_channels=8),
    dict(B2, attention_levels=(False, False, True), num_head_channels=8, resblock_updown=True),
    dict(B2, attention_levels=(False, False, True), num_head_channels=4),
    dict(B2, attention_levels=(False, True, True), num_head_channels=(0, 2, 4)),
    dict(B3, attention_levels=(False, False, False)),
    dict(B3, attention_levels=(False, False, False), resblock_updown=True),
    dict(B3, attention_levels=(False, False, True), num_head_channels=8),
    dict(B3, attention_levels=(False, False, True), num_head_channels=8, resblock_updown=True),
    dict(B3, attention_levels=(False, False, True), num_head_channels=4),
    dict(B3, attention_levels=(False, False, True), num_head_channels=(0, 0, 4)),
]
COND = [
    dict(B2, attention_levels=(False, False, True), num_head_channels=4, with_conditioning=True,
         transformer_num_layers=1, cross_attention_dim=3),
    dict(B2, attention_levels=(False, False, True), num_head_channels=4, with_conditioning=True,
         transformer_num_layers=1, cross_attention_dim=3, resblock_updown=True),
    dict(B2, attention_levels=(False, False, True), num_head_channels=4, with_conditioning=True,
         transformer_num_layers=1, cross_attention_dim=3, upcast_attention=True),
]


@pytest.mark.parametrize("kw", UNCOND, ids=[str(i) for i in range(len(UNCOND))])
def test_unet_unconditioned_cases(cuda_device, kw):
    torch.manual_seed(0)
    m = G.randomize_zero_params(nets().DiffusionModelUNet(**kw)).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    shape = (1, 1, 16, 16) if kw["spatial_dims"] == 2 else (1, 1, 16, 16, 16)
    torch.manual_seed(1)
    x, t = torch.rand(shape), torch.randint(0, 1000, (1,)).long()
    want = O.unet_forward(sd, G.unet_oracle_cfg(kw), x, t)
    got = m.cuda()(x.cuda(), timesteps=t.cuda())
    assert got.shape == want.shape == shape
    assert rel(got, want) < TOL, rel(got, want)


@pytest.mark.parametrize("kw", COND, ids=[str(i) for i in range(len(COND))])
def test_unet_conditioned_cases(cuda_device, kw):
    torch.manual_seed(0)
    m = G.randomize_zero_params(nets().DiffusionModelUNet(**kw)).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    torch.manual_seed(1)
    x, t, ctx = torch.rand(1, 1, 16, 32), torch.randint(0, 1000, (1,)).long(), torch.rand(1, 1, 3)
    want = O.unet_forward(sd, G.unet_oracle_cfg(kw), x, t, context=ctx)
    m = m.cuda()
    got = m(x.cuda(), timesteps=t.cuda(), context=ctx.cuda())
    assert got.shape == (1, 1, 16, 32) and rel(got, want) < TOL, rel(got, want)
    with pytest.raises(ValueError):            # timesteps must be 1-D (test_timestep_with_wrong_shape)
        m(x.cuda(), timesteps=torch.randint(0, 1000, (1, 1)).cuda(), context=ctx.cuda())


def test_unet_error_paths(cuda_device):
    """test_context_with_conditioning_none / test_shape_with_additional_inputs / class-label requirements."""
    m = nets().DiffusionModelUNet(**dict(B2, attention_levels=(False, False, True), num_head_channels=4)).cuda().eval()
    with pytest.raises(ValueError):
        m(torch.rand(1, 1, 16, 32).cuda(), timesteps=torch.randint(0, 1000, (1,)).cuda(), context=torch.rand(1, 1, 3).cuda())
    kw = dict(B2, attention_levels=(False, False, True), num_head_channels=4, num_class_embeds=2)
    m = G.randomize_zero_params(nets().DiffusionModelUNet(**kw)).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    x, t, c = torch.rand(1, 1, 16, 32), torch.randint(0, 1000, (1,)).long(), torch.randint(0, 2, (1,)).long()
    m = m.cuda()
    with pytest.raises(ValueError):
        m(x.cuda(), timesteps=t.cuda())
    got = m(x.cuda(), timesteps=t.cuda(), class_labels=c.cuda())
    assert rel(got, O.unet_forward(sd, G.unet_oracle_cfg(kw), x, t, class_labels=c)) < TOL


AE = dict(spatial_dims=2, in_channels=1, out_channels=1, num_channels=(4, 4, 4), latent_channels=4, norm_num_groups=4)
AE_CASES = [
    (dict(AE, attention_levels=(False, False, False), num_res_blocks=1), (1, 1, 16, 16)),
    (dict(AE, attention_levels=(False, False, False), num_res_blocks=(1, 1, 2)), (1, 1, 16, 16)),
    (dict(AE, attention_levels=(False, False, True), num_res_blocks=1), (1, 1, 16, 16)),
    (dict(AE, attention_levels=(False, False, False), num_res_blocks=1, with_encoder_nonlocal_attn=False), (1, 1, 16, 16)),
    (dict(AE, attention_levels=(False, False, False), num_res_blocks=1, with_decoder_nonlocal_attn=False), (1, 1, 16, 16)),
    (dict(AE, attention_levels=(False, False, True), num_res_blocks=1, use_convtranspose=True), (1, 1, 16, 16)),
    (dict(AE, spatial_dims=3, attention_levels=(False, False, True), num_res_blocks=1), (1, 1, 16, 16, 16)),
]


@pytest.mark.parametrize("kw,shape", AE_CASES, ids=[str(i) for i in range(len(AE_CASES))])
def test_autoencoderkl_cases(cuda_device, kw, shape):
    torch.manual_seed(0)
    m = nets().AutoencoderKL(**kw).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    cfg = G.aekl_oracle_cfg(kw)
    torch.manual_seed(2)
    x = torch.randn(shape)
    mu_w, sig_w = O.autoencoderkl_encode(sd, cfg, x)
    m = m.cuda()
    mu, sig = m.encode(x.cuda())
    lat = tuple(s // 4 for s in shape[2:])
    assert mu.shape == (1, 4, *lat) == tuple(mu_w.shape)
    assert rel(mu, mu_w) < TOL and rel(sig, sig_w) < TOL
    rec_w = O.autoencoderkl_decode(sd, cfg, mu_w)
    assert rel(m.decode(mu_w.cuda()), rec_w) < TOL
    rec, mu2, sig2 = m(x.cuda())            # forward: (reconstruction, z_mu, z_sigma)
    assert rec.shape == shape and mu2.shape == mu.shape and sig2.shape == sig.shape
    assert m.reconstruct(x.cuda()).shape == shape
    assert m.encode_stage_2_inputs(x.cuda()).shape == mu.shape
    assert m.decode_stage_2_outputs(mu).shape == shape


VQ = dict(in_channels=1, out_channels=1, num_res_layers=1, num_embeddings=8, embedding_dim=8)
VQ_CASES = [
    (dict(VQ, spatial_dims=2, num_channels=(4, 4), num_res_channels=4, downsample_parameters=((2, 4, 1, 1),) * 2,
          upsample_parameters=((2, 4, 1, 1, 0),) * 2), (1, 1, 8, 8)),
    (dict(VQ, spatial_dims=2, num_channels=(4, 4), num_res_channels=(4, 4), downsample_parameters=(2, 4, 1, 1),
          upsample_parameters=(2, 4, 1, 1, 0)), (1, 1, 8, 8)),
    (dict(VQ, spatial_dims=3, num_channels=[4, 4], num_res_channels=[4, 4], downsample_parameters=((2, 4, 1, 1),) * 2,
          upsample_parameters=((2, 4, 1, 1, 0),) * 2), (1, 1, 8, 8, 8)),
]


@pytest.mark.parametrize("kw,shape", VQ_CASES, ids=[str(i) for i in range(len(VQ_CASES))])
def test_vqvae_cases(cuda_device, kw, shape):
    torch.manual_seed(0)
    m = nets().VQVAE(**kw).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    dp, up = kw["downsample_parameters"], kw["upsample_parameters"]
    n = len(kw["num_channels"])
    cfg = dict(downsample_parameters=dp if isinstance(dp[0], (tuple, list)) else (dp,) * n,
               upsample_parameters=up if isinstance(up[0], (tuple, list)) else (up,) * n)
    torch.manual_seed(3)
    x = torch.rand(shape)
    z_w = O.vqvae_encode(sd, cfg, x)
    m = m.cuda()
    z = m.encode(x.cuda())
    lat = tuple(s // 4 for s in shape[2:])
    assert z.shape == (1, 8, *lat) and rel(z, z_w) < TOL
    rec, loss = m(x.cuda())
    assert rec.shape == shape and loss.numel() == 1
    idx = m.index_quantize(x.cuda())
    assert idx.shape == (1, *lat) and idx.dtype == torch.int64
    assert m.decode_samples(idx).shape == shape
    q_w, _, idx_w, _ = O.vq_forward(sd["quantizer.quantizer.embedding.weight"], z_w)
    assert rel(m.decode(q_w.cuda()), O.vqvae_decode(sd, cfg, q_w)) < TOL


def test_controlnet_cases(cuda_device):
    """tests/test_controlnet.py:22-50: 6 down residuals + mid for 3 levels with 1 res block."""
    for extra in (dict(attention_levels=(False, False, False)),
                  dict(attention_levels=(False, False, True), num_head_channels=4, resblock_updown=True)):
        kw = dict(spatial_dims=2, in_channels=1, num_res_blocks=1, num_channels=(8, 8, 8), norm_num_groups=8,
                  conditioning_embedding_in_channels=1, conditioning_embedding_num_channels=(8, 8), **extra)
        torch.manual_seed(0)
        m = G.randomize_zero_params(nets().ControlNet(**kw)).eval()
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        torch.manual_seed(4)
        x, t, c = torch.rand(1, 1, 16, 16), torch.randint(0, 1000, (1,)).long(), torch.rand(1, 1, 32, 32)
        d_w, m_w = O.controlnet_forward(sd, G.unet_oracle_cfg(kw), x, t, c)
        down, mid = m.cuda()(x.cuda(), timesteps=t.cuda(), controlnet_cond=c.cuda())
        assert len(down) == 6 and mid.shape == (1, 8, 4, 4)
        for a, b in zip(down, d_w):
            assert a.shape == b.shape and rel(a, b) < TOL
        assert rel(mid, m_w) < TOL


@pytest.mark.parametrize("sd_", [2, 3])
def test_inferer_cases(cuda_device, sd_):
    """tests/test_diffusion_inferer.py:23-50 (8x8 / 8x8x8, one level, 8 channels): sample with intermediates,
    DDPM and DDIM samplers, crossattn and concat conditioning."""
    from generativemodels_b200.inferers import DiffusionInferer
    from generativemodels_b200.networks.schedulers import DDIMScheduler, DDPMScheduler
    kw = dict(spatial_dims=sd_, in_channels=1, out_channels=1, num_channels=[8], norm_num_groups=8,
              attention_levels=[True], num_res_blocks=1, num_head_channels=8)
    shape = (2, 1, 8, 8) if sd_ == 2 else (2, 1, 8, 8, 8)
    m = G.randomize_zero_params(nets().DiffusionModelUNet(**kw)).cuda().eval()
    noise = torch.randn(shape).cuda()
    s = DDPMScheduler(num_train_timesteps=10)
    s.set_timesteps(10)
    sample, inter = DiffusionInferer(s).sample(noise, m, s, save_intermediates=True, intermediate_steps=1, verbose=False)
    assert sample.shape == shape and len(inter) == 10
    s = DDIMScheduler(num_train_timesteps=1000)
    s.set_timesteps(10)
    sample, inter = DiffusionInferer(s).sample(noise, m, s, save_intermediates=True, intermediate_steps=100, verbose=False)
    assert sample.shape == shape and len(inter) == 10
    # crossattn conditioning
    kwc = dict(kw, with_conditioning=True, cross_attention_dim=3)
    mc = G.randomize_zero_params(nets().DiffusionModelUNet(**kwc)).cuda().eval()
    cond = torch.randn(2, 1, 3).cuda()
    out = DiffusionInferer(s).sample(noise, mc, s, conditioning=cond, verbose=False)
    assert out.shape == shape and torch.isfinite(out).all()
    # concat conditioning (test_sampler_conditioned_concat): model takes in_channels + condition channels
    kwk = dict(kw, in_channels=2)
    mk = G.randomize_zero_params(nets().DiffusionModelUNet(**kwk)).cuda().eval()
    out = DiffusionInferer(s).sample(noise, mk, s, conditioning=torch.randn(shape).cuda(), mode="concat", verbose=False)
    assert out.shape == shape
    # training-style call: add_noise at per-sample timesteps + forward (DiffusionInferer.__call__)
    pred = DiffusionInferer(s)(inputs=torch.randn(shape).cuda(), diffusion_model=m, noise=noise,
                               timesteps=torch.randint(0, 1000, (2,)).cuda())
    assert pred.shape == shape
    # test_get_likelihood (138-151) and test_normal_cdf (153-162)
    s = DDPMScheduler(num_train_timesteps=10)
    s.set_timesteps(10)
    lik, inter = DiffusionInferer(s).get_likelihood(inputs=torch.randn(shape).cuda(), diffusion_model=m, scheduler=s,
                                                    save_intermediates=True, verbose=False)
    assert inter[0].shape == shape and lik.shape[0] == shape[0] and torch.isfinite(lik).all() and len(inter) == 10
    from scipy.stats import norm
    x = torch.linspace(-10, 10, 20)
    torch.testing.assert_close(DiffusionInferer(s)._approx_standard_normal_cdf(x).double(),
                               torch.from_numpy(norm.cdf(x.numpy())), atol=1e-3, rtol=1e-5)


def test_latent_inferer_cases(cuda_device):
    """tests/test_latent_diffusion_inferer.py: AutoencoderKL / VQVAE stage-1 with a latent UNet, incl. the
    different-latent-shape resizer path (ldm_latent_shape / autoencoder_latent_shape, 675-730)."""
    from generativemodels_b200.inferers import LatentDiffusionInferer
    from generativemodels_b200.networks.schedulers import DDPMScheduler
    ae = nets().AutoencoderKL(spatial_dims=2, in_channels=1, out_channels=1, num_channels=(4, 4), latent_channels=3,
                              attention_levels=[False, False], num_res_blocks=1, with_encoder_nonlocal_attn=False,
                              with_decoder_nonlocal_attn=False, norm_num_groups=4).cuda().eval()
    un = G.randomize_zero_params(nets().DiffusionModelUNet(
        spatial_dims=2, in_channels=3, out_channels=3, num_channels=[4, 4], norm_num_groups=4,
        attention_levels=[False, False], num_res_blocks=1, num_head_channels=4)).cuda().eval()
    s = DDPMScheduler(num_train_timesteps=10)
    s.set_timesteps(10)
    inf = LatentDiffusionInferer(s, scale_factor=1.0)
    noise = torch.randn(1, 3, 4, 4).cuda()
    sample, inter = inf.sample(noise, ae, un, s, save_intermediates=True, intermediate_steps=1, verbose=False)
    assert sample.shape == (1, 1, 8, 8) and len(inter) == 10 and inter[0].shape == (1, 1, 8, 8)
    pred = inf(inputs=torch.randn(1, 1, 8, 8).cuda(), autoencoder_model=ae, diffusion_model=un, noise=noise,
               timesteps=torch.randint(0, 10, (1,)).cuda())
    assert pred.shape == (1, 3, 4, 4)
    # VQVAE stage 1
    vq = nets().VQVAE(spatial_dims=2, in_channels=1, out_channels=1, num_channels=[4, 4], num_res_layers=1,
                      num_res_channels=[4, 4], downsample_parameters=((2, 4, 1, 1),) * 2,
                      upsample_parameters=((2, 4, 1, 1, 0),) * 2, num_embeddings=16, embedding_dim=3).cuda().eval()
    sample = inf.sample(noise[:, :, :2, :2].contiguous(), vq, un, s, verbose=False)
    assert sample.shape == (1, 1, 8, 8)
    # latent-shape resizing: the LDM samples an 8x8 latent, the autoencoder decodes its 4x4 centre crop
    inf2 = LatentDiffusionInferer(s, scale_factor=1.0, ldm_latent_shape=[8, 8], autoencoder_latent_shape=[4, 4])
    sample = inf2.sample(torch.randn(1, 3, 8, 8).cuda(), ae, un, s, verbose=False)
    assert sample.shape == (1, 1, 8, 8)
    # test_get_likelihoods (438-489) / test_resample_likelihoods (492-545), both stage-1 families
    img = torch.randn(1, 1, 8, 8).cuda()
    for stage1, lat_shape in ((ae, (1, 3, 4, 4)), (vq, (1, 3, 2, 2))):
        for quantized in (True, False):
            lik, inter = inf.get_likelihood(inputs=img, autoencoder_model=stage1, diffusion_model=un, scheduler=s,
                                            save_intermediates=True, quantized=quantized, verbose=False)
            assert len(inter) == 10 and inter[0].shape == lat_shape and torch.isfinite(lik).all()
        lik, inter = inf.get_likelihood(inputs=img, autoencoder_model=stage1, diffusion_model=un, scheduler=s,
                                        save_intermediates=True, resample_latent_likelihoods=True, verbose=False)
        assert len(inter) == 10 and inter[0].shape[2:] == img.shape[2:]
    with pytest.raises(ValueError):
        inf.get_likelihood(inputs=img, autoencoder_model=ae, diffusion_model=un, scheduler=s, verbose=False,
                           resample_latent_likelihoods=True, resample_interpolation_mode="cubic")


@pytest.mark.parametrize("sd_", [2, 3])
def test_controlnet_inferer_cases(cuda_device, sd_):
    """tests/test_controlnet_inferers.py:30-80 (CNDM_TEST_CASES) — call, DDPM/DDIM sample with intermediates,
    crossattn / concat conditioning, get_likelihood (568-590)."""
    from generativemodels_b200.inferers import ControlNetDiffusionInferer
    from generativemodels_b200.networks.schedulers import DDIMScheduler, DDPMScheduler
    base = dict(spatial_dims=sd_, in_channels=1, num_channels=[8], norm_num_groups=8, attention_levels=[True],
                num_res_blocks=1, num_head_channels=8)
    cnkw = dict(conditioning_embedding_num_channels=[16], conditioning_embedding_in_channels=1)
    shape = (2, 1, 8, 8) if sd_ == 2 else (2, 1, 8, 8, 8)
    m = G.randomize_zero_params(nets().DiffusionModelUNet(out_channels=1, **base)).cuda().eval()
    cn = G.randomize_zero_params(nets().ControlNet(**base, **cnkw)).cuda().eval()
    x, mask, noise = (torch.randn(shape).cuda() for _ in range(3))
    s = DDPMScheduler(num_train_timesteps=10)
    s.set_timesteps(10)
    inf = ControlNetDiffusionInferer(s)
    pred = inf(inputs=x, diffusion_model=m, controlnet=cn, noise=noise, timesteps=torch.randint(0, 10, (2,)).cuda(),
               cn_cond=mask)
    assert pred.shape == shape
    sample, inter = inf.sample(input_noise=noise, diffusion_model=m, controlnet=cn, cn_cond=mask, scheduler=s,
                               save_intermediates=True, intermediate_steps=1, verbose=False)
    assert sample.shape == shape and len(inter) == 10 and torch.isfinite(sample).all()
    d = DDIMScheduler(num_train_timesteps=1000)
    d.set_timesteps(10)
    sample, inter = inf.sample(input_noise=noise, diffusion_model=m, controlnet=cn, cn_cond=mask, scheduler=d,
                               save_intermediates=True, intermediate_steps=100, verbose=False)
    assert sample.shape == shape and len(inter) == 10
    lik, inter = inf.get_likelihood(inputs=x, diffusion_model=m, controlnet=cn, cn_cond=mask, scheduler=s,
                                    save_intermediates=True, verbose=False)
    assert inter[0].shape == shape and lik.shape[0] == shape[0] and torch.isfinite(lik).all()
    # crossattn and concat conditioning (539-566, 603-640)
    ckw = dict(base, with_conditioning=True, cross_attention_dim=3)
    mc = G.randomize_zero_params(nets().DiffusionModelUNet(out_channels=1, **ckw)).cuda().eval()
    cc = G.randomize_zero_params(nets().ControlNet(**ckw, **cnkw)).cuda().eval()
    out = inf.sample(input_noise=noise, diffusion_model=mc, controlnet=cc, cn_cond=mask, scheduler=d,
                     conditioning=torch.randn(2, 1, 3).cuda(), verbose=False)
    assert out.shape == shape
    kkw = dict(base, in_channels=2)
    mk = G.randomize_zero_params(nets().DiffusionModelUNet(out_channels=1, **kkw)).cuda().eval()
    ck = G.randomize_zero_params(nets().ControlNet(**kkw, **cnkw)).cuda().eval()
    out = inf.sample(input_noise=noise, diffusion_model=mk, controlnet=ck, cn_cond=mask, scheduler=d,
                     conditioning=torch.randn(shape).cuda(), mode="concat", verbose=False)
    assert out.shape == shape
    lik = inf.get_likelihood(inputs=x, diffusion_model=mk, controlnet=ck, cn_cond=mask, scheduler=s,
                             conditioning=torch.randn(shape).cuda(), mode="concat", verbose=False)
    assert lik.shape == (2,) and torch.isfinite(lik).all()


def test_controlnet_latent_inferer_cases(cuda_device):
    """tests/test_controlnet_inferers.py:643-970 (LATENT_CNDM_TEST_CASES): prediction, sample, likelihood and
    resampled likelihood maps with an AutoencoderKL / VQVAE first stage; the ControlNet mask is given at image
    resolution and resized to the latent grid."""
    from generativemodels_b200.inferers import ControlNetLatentDiffusionInferer
    from generativemodels_b200.networks.schedulers import DDPMScheduler
    ae = nets().AutoencoderKL(spatial_dims=2, in_channels=1, out_channels=1, num_channels=(4, 4), latent_channels=3,
                              attention_levels=[False, False], num_res_blocks=1, with_encoder_nonlocal_attn=False,
                              with_decoder_nonlocal_attn=False, norm_num_groups=4).cuda().eval()
    vq = nets().VQVAE(spatial_dims=2, in_channels=1, out_channels=1, num_channels=[4, 4], num_res_layers=1,
                      num_res_channels=[4, 4], downsample_parameters=((2, 4, 1, 1),) * 2,
                      upsample_parameters=((2, 4, 1, 1, 0),) * 2, num_embeddings=16, embedding_dim=3).cuda().eval()
    base = dict(spatial_dims=2, in_channels=3, num_channels=[4, 4], norm_num_groups=4, attention_levels=[False, False],
                num_res_blocks=1, num_head_channels=4)
    un = G.randomize_zero_params(nets().DiffusionModelUNet(out_channels=3, **base)).cuda().eval()
    cn = G.randomize_zero_params(nets().ControlNet(conditioning_embedding_num_channels=[16],
                                                   conditioning_embedding_in_channels=1, **base)).cuda().eval()
    s = DDPMScheduler(num_train_timesteps=10)
    s.set_timesteps(10)
    inf = ControlNetLatentDiffusionInferer(s, scale_factor=1.0)
    img, mask = torch.randn(1, 1, 8, 8).cuda(), torch.randn(1, 1, 8, 8).cuda()
    for stage1, lat in ((ae, (1, 3, 4, 4)), (vq, (1, 3, 2, 2))):
        noise = torch.randn(lat).cuda()
        pred = inf(inputs=img, autoencoder_model=stage1, diffusion_model=un, controlnet=cn, noise=noise,
                   timesteps=torch.randint(0, 10, (1,)).cuda(), cn_cond=mask)
        assert pred.shape == lat
        sample, inter = inf.sample(input_noise=noise, autoencoder_model=stage1, diffusion_model=un, controlnet=cn,
                                   cn_cond=mask, scheduler=s, save_intermediates=True, intermediate_steps=1,
                                   verbose=False)
        assert sample.shape == img.shape and len(inter) == 10 and inter[0].shape == img.shape
        lik, inter = inf.get_likelihood(inputs=img, autoencoder_model=stage1, diffusion_model=un, controlnet=cn,
                                        cn_cond=mask, scheduler=s, save_intermediates=True, verbose=False)
        assert len(inter) == 10 and inter[0].shape == lat and torch.isfinite(lik).all()
        lik, inter = inf.get_likelihood(inputs=img, autoencoder_model=stage1, diffusion_model=un, controlnet=cn,
                                        cn_cond=mask, scheduler=s, save_intermediates=True,
                                        resample_latent_likelihoods=True, verbose=False)
        assert inter[0].shape[2:] == img.shape[2:]


def test_spade_cases(cuda_device):
    """tests/test_spade_diffusion_model_unet.py (shapes, wrong timestep / label shapes, channel validation, class and
    cross-attention conditioning) and tests/test_spade_autoencoderkl.py (shape, encode / sampling / decode)."""
    base = dict(spatial_dims=2, label_nc=3, in_channels=1, out_channels=1, num_res_blocks=1, num_channels=(8, 8, 8),
                attention_levels=(False, False, False), norm_num_groups=8)
    net = G.randomize_zero_params(nets().SPADEDiffusionModelUNet(**base)).cuda().eval()
    x, t, seg = torch.rand(1, 1, 16, 16).cuda(), torch.randint(0, 1000, (1,)).long().cuda(), torch.rand(1, 3, 16, 16).cuda()
    assert net(x, t, seg).shape == (1, 1, 16, 16)
    with pytest.raises(ValueError):
        net(x, torch.randint(0, 1000, (1, 1)).long().cuda(), seg)
    with pytest.raises(RuntimeError):
        net(x, t, torch.rand(1, 6, 16, 16).cuda())
    io = G.randomize_zero_params(nets().SPADEDiffusionModelUNet(**dict(base, in_channels=6, out_channels=3))).cuda().eval()
    assert io(torch.rand(1, 6, 16, 16).cuda(), t, seg).shape == (1, 3, 16, 16)
    for bad in (dict(num_channels=(8, 8, 12)), dict(num_channels=(8, 8), attention_levels=(False, False, False)),
                dict(num_res_blocks=(1, 1)), dict(attention_levels=(False, True, True), num_head_channels=(0, 2))):
        with pytest.raises(ValueError):
            nets().SPADEDiffusionModelUNet(**dict(base, **bad))
    with pytest.raises(ValueError):
        nets().SPADEDiffusionModelUNet(**dict(base, with_conditioning=True, cross_attention_dim=None))
    cond = G.randomize_zero_params(nets().SPADEDiffusionModelUNet(**dict(
        base, attention_levels=(False, False, True), num_head_channels=8, with_conditioning=True,
        transformer_num_layers=1, cross_attention_dim=3, num_class_embeds=2))).cuda().eval()
    out = cond(torch.rand(1, 1, 16, 32).cuda(), t, torch.rand(1, 3, 16, 32).cuda(), context=torch.rand(1, 1, 3).cuda(),
               class_labels=torch.randint(0, 2, (1,)).long().cuda())
    assert out.shape == (1, 1, 16, 32)
    with pytest.raises(ValueError):
        cond(torch.rand(1, 1, 16, 32).cuda(), t, torch.rand(1, 3, 16, 32).cuda(), context=torch.rand(1, 1, 3).cuda())
    net3 = G.randomize_zero_params(nets().SPADEDiffusionModelUNet(**dict(base, spatial_dims=3))).cuda().eval()
    assert net3(torch.rand(1, 1, 16, 16, 16).cuda(), t, torch.rand(1, 3, 16, 16, 16).cuda()).shape == (1, 1, 16, 16, 16)
    # autoencoder
    akw = dict(spatial_dims=2, label_nc=3, in_channels=1, out_channels=1, num_channels=(4, 4, 4), latent_channels=4,
               attention_levels=(False, False, False), num_res_blocks=1, norm_num_groups=4)
    ae = nets().SPADEAutoencoderKL(**akw).cuda().eval()
    img, sg = torch.randn(1, 1, 16, 16).cuda(), torch.randn(1, 3, 16, 16).cuda()
    rec, mu, sigma = ae(img, sg)
    assert rec.shape == (1, 1, 16, 16) and mu.shape == (1, 4, 4, 4) and sigma.shape == (1, 4, 4, 4)
    assert ae.sampling(mu, sigma).shape == (1, 4, 4, 4)
    assert ae.decode(torch.randn(1, 4, 4, 4).cuda(), sg).shape == (1, 1, 16, 16)
    for bad in (dict(num_channels=(24, 24, 24), norm_num_groups=16), dict(attention_levels=(False, False)),
                dict(num_res_blocks=(8, 8))):
        with pytest.raises(ValueError):
            nets().SPADEAutoencoderKL(**dict(akw, **bad))


@pytest.mark.parametrize("sd_", [2, 3])
def test_transformer_inferer_cases(cuda_device, sd_):
    """tests/test_transformer.py:20-45 and tests/test_vqvaetransformer_inferer.py:24-275: prediction shapes (also with
    max_seq_len shorter than the sequence), sampling, likelihood and resampled likelihood maps."""
    from generativemodels_b200.inferers import VQVAETransformerInferer
    from generativemodels_b200.utils.ordering import Ordering
    net = nets().DecoderOnlyTransformer(num_tokens=10, max_seq_len=16, attn_layers_dim=8, attn_layers_depth=2,
                                        attn_layers_heads=2).cuda().eval()
    assert net(torch.randint(0, 10, (1, 16)).cuda()).shape == (1, 16, 10)
    netc = nets().DecoderOnlyTransformer(num_tokens=10, max_seq_len=16, attn_layers_dim=8, attn_layers_depth=2,
                                         attn_layers_heads=2, with_cross_attention=True,
                                         embedding_dropout_rate=0).cuda().eval()
    assert netc(torch.randint(0, 10, (1, 16)).cuda(), context=torch.randn(1, 4, 8).cuda()).shape == (1, 16, 10)
    vq = nets().VQVAE(spatial_dims=sd_, in_channels=1, out_channels=1, num_channels=(8, 8), num_res_channels=(8, 8),
                      downsample_parameters=((2, 4, 1, 1),) * 2, upsample_parameters=((2, 4, 1, 1, 0),) * 2,
                      num_res_layers=1, num_embeddings=16, embedding_dim=8).cuda().eval()
    n_lat = 2 ** sd_
    shape = (2, 1) + (8,) * sd_
    lat_dims = (2,) * sd_
    ordering = Ordering(ordering_type="raster_scan", spatial_dims=sd_, dimensions=(2,) + lat_dims)
    inf = VQVAETransformerInferer()
    x = torch.randn(shape).cuda()
    for max_len in (n_lat, 2):                   # full-length and shorter-than-sequence windows
        tr = nets().DecoderOnlyTransformer(num_tokens=17, max_seq_len=max_len, attn_layers_dim=4, attn_layers_depth=2,
                                           attn_layers_heads=1).cuda().eval()
        assert inf(inputs=x, vqvae_model=vq, transformer_model=tr, ordering=ordering).shape == (2, max_len, 17)
        sample = inf.sample(latent_spatial_dim=lat_dims, starting_tokens=16 * torch.ones((2, 1)).cuda(),
                            vqvae_model=vq, transformer_model=tr, ordering=ordering, verbose=False)
        assert sample.shape == shape and torch.isfinite(sample).all()
        ll = inf.get_likelihood(inputs=x, vqvae_model=vq, transformer_model=tr, ordering=ordering)
        assert ll.shape == (2,) + lat_dims
        up = inf.get_likelihood(inputs=x, vqvae_model=vq, transformer_model=tr, ordering=ordering,
                                resample_latent_likelihoods=True, resample_interpolation_mode="nearest")
        assert up.shape == shape
