"""Pin oracle/torch_oracle.py against the UNMODIFIED reference (imported from /root/reference on the MONAI shim).
Runs only where the reference tree exists (the build container); the GPU box relies on tests/golden fixtures."""
import pytest
import torch

from oracle import ref_import
from oracle import torch_oracle as O
from tests.golden import configs as G

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    ref_import.import_reference()
    import generative.networks.nets as nets
    import generative.networks.schedulers as sch
    return nets, sch


def _close(a, b, tol=2e-5):
    err = (a - b).abs().max().item()
    assert err <= tol * max(1.0, b.abs().max().item()), f"max abs err {err:.3e}"


@pytest.mark.parametrize("name", list(G.UNET_CASES))
def test_unet(ref, name):
    nets, _ = ref
    kw = G.UNET_CASES[name]
    torch.manual_seed(0)
    m = G.randomize_zero_params(nets.DiffusionModelUNet(**kw)).eval()
    inp = G.UNET_INPUTS[name]
    torch.manual_seed(1)
    x = torch.randn(inp["shape"])
    t = torch.randint(0, 1000, (inp["shape"][0],)).long()
    ctx = torch.randn(inp["context"]) if "context" in inp else None
    cls = torch.randint(0, 5, (inp["shape"][0],)) if inp.get("classes") else None
    with torch.no_grad():
        want = m(x, t, context=ctx, class_labels=cls)
        got = O.unet_forward(m.state_dict(), G.unet_oracle_cfg(kw), x, t, context=ctx, class_labels=cls)
    _close(got, want)


def test_controlnet(ref):
    nets, _ = ref
    kw = G.CONTROLNET_CASE
    torch.manual_seed(0)
    cn = G.randomize_zero_params(nets.ControlNet(**kw)).eval()
    ukw = {k: v for k, v in kw.items() if not k.startswith("conditioning_embedding")}
    un = G.randomize_zero_params(nets.DiffusionModelUNet(out_channels=3, **ukw)).eval()
    torch.manual_seed(2)
    x, cond, ctx = torch.randn(2, 3, 16, 16), torch.rand(2, 1, 16, 16), torch.randn(2, 1, 8)
    t = torch.tensor([10, 500]).long()
    cfg = G.unet_oracle_cfg(kw)
    with torch.no_grad():
        d_w, m_w = cn(x, t, cond, conditioning_scale=0.7, context=ctx)
        d_g, m_g = O.controlnet_forward(cn.state_dict(), cfg, x, t, cond, 0.7, ctx)
        for a, b in zip(d_g, d_w):
            _close(a, b)
        _close(m_g, m_w)
        want = un(x, t, context=ctx, down_block_additional_residuals=d_w, mid_block_additional_residual=m_w)
        got = O.unet_forward(un.state_dict(), cfg, x, t, context=ctx, down_block_additional_residuals=d_g,
                             mid_block_additional_residual=m_g)
    _close(got, want)


@pytest.mark.parametrize("name", list(G.AEKL_CASES))
def test_autoencoderkl(ref, name):
    nets, _ = ref
    kw = G.AEKL_CASES[name]
    torch.manual_seed(0)
    m = nets.AutoencoderKL(**kw).eval()
    torch.manual_seed(3)
    x = torch.randn(G.AEKL_INPUTS[name])
    cfg = G.aekl_oracle_cfg(kw)
    with torch.no_grad():
        mu_w, sig_w = m.encode(x)
        mu_g, sig_g = O.autoencoderkl_encode(m.state_dict(), cfg, x)
        _close(mu_g, mu_w)
        _close(sig_g, sig_w)
        _close(O.autoencoderkl_decode(m.state_dict(), cfg, mu_w), m.decode(mu_w))


@pytest.mark.parametrize("name", list(G.VQVAE_CASES))
def test_vqvae(ref, name):
    nets, _ = ref
    kw = G.VQVAE_CASES[name]
    torch.manual_seed(0)
    m = nets.VQVAE(**kw).eval()
    torch.manual_seed(4)
    x = torch.rand(G.VQVAE_INPUTS[name])
    cfg = G.vqvae_oracle_cfg(kw)
    sd = m.state_dict()
    with torch.no_grad():
        rec_w, loss_w = m(x)
        idx_w = m.index_quantize(x)
        rec_g, loss_g, idx_g = O.vqvae_forward(sd, cfg, x)
        assert torch.equal(idx_g, idx_w)
        _close(rec_g, rec_w)
        _close(loss_g, loss_w)
        _close(O.vqvae_decode(sd, cfg, O.vq_embed(sd["quantizer.quantizer.embedding.weight"], idx_w)),
               m.decode_samples(idx_w))
        _close(O.vq_forward(sd["quantizer.quantizer.embedding.weight"], m.encode(x))[3], m.quantizer.perplexity)


SCHED = [
    ("ddim", dict(num_train_timesteps=1000, schedule="linear_beta", beta_start=0.0015, beta_end=0.0195), 50),
    ("ddim", dict(num_train_timesteps=1000, schedule="scaled_linear_beta", beta_start=0.0005, beta_end=0.0195,
                  clip_sample=False), 50),
    ("ddim", dict(num_train_timesteps=100, prediction_type="v_prediction", set_alpha_to_one=False, steps_offset=1), 10),
    # (cosine has alphas_cumprod[0] == 1 exactly: t = 0 divides by zero in the reference itself, so offset by 1)
    ("ddim", dict(num_train_timesteps=100, prediction_type="sample", schedule="cosine", steps_offset=1,
                  set_alpha_to_one=False), 7),
    ("ddpm", dict(num_train_timesteps=1000), 4),
    ("ddpm", dict(num_train_timesteps=100, variance_type="fixed_large", prediction_type="v_prediction",
                  schedule="sigmoid_beta"), 10),
    ("pndm", dict(num_train_timesteps=1000, skip_prk_steps=True), 20),
    ("pndm", dict(num_train_timesteps=1000, skip_prk_steps=False), 20),
    ("pndm", dict(num_train_timesteps=100, skip_prk_steps=False, prediction_type="v_prediction",
                  set_alpha_to_one=True, steps_offset=1), 10),
]


@pytest.mark.parametrize("kind,kw,steps", SCHED, ids=[f"{k}-{i}" for i, (k, _, _) in enumerate(SCHED)])
def test_scheduler_trajectory(ref, kind, kw, steps):
    """Identical model-output sequence through the reference scheduler and the restated one: every step equal."""
    _, sch = ref
    R = {"ddim": sch.DDIMScheduler, "ddpm": sch.DDPMScheduler, "pndm": sch.PNDMScheduler}[kind](**kw)
    M = {"ddim": O.DDIMOracle, "ddpm": O.DDPMOracle, "pndm": O.PNDMOracle}[kind](**kw)
    R.set_timesteps(steps)
    M.set_timesteps(steps)
    assert torch.equal(R.timesteps, M.timesteps)
    assert torch.equal(R.alphas_cumprod, M.alphas_cumprod)
    torch.manual_seed(5)
    xr = xm = torch.randn(2, 3, 8, 8)
    for t in R.timesteps:
        eps = torch.tanh(xr * 0.7 + 0.01 * float(t))      # any deterministic stand-in for the network
        epsm = torch.tanh(xm * 0.7 + 0.01 * float(t))
        if kind == "ddpm":
            gr, gm = torch.Generator().manual_seed(int(t)), torch.Generator().manual_seed(int(t))
            xr, _ = R.step(eps, int(t), xr, generator=gr)
            xm, _ = M.step(epsm, int(t), xm, generator=gm)
        else:
            xr, _ = R.step(eps, int(t), xr)
            xm, _ = M.step(epsm, int(t), xm)
        assert torch.equal(xr, xm), f"diverged at t={int(t)}"


def test_inferer_sample(ref):
    nets, sch = ref
    from generative.inferers import DiffusionInferer
    kw = G.UNET_CASES["unet2d_attn"]
    torch.manual_seed(0)
    m = G.randomize_zero_params(nets.DiffusionModelUNet(**kw)).eval()
    skw = dict(num_train_timesteps=1000, schedule="linear_beta", beta_start=0.0015, beta_end=0.0195)
    R, M = sch.DDIMScheduler(**skw), O.DDIMOracle(**skw)
    R.set_timesteps(5)
    M.set_timesteps(5)
    torch.manual_seed(7)
    noise = torch.randn(1, 1, 16, 16)
    cfg = G.unet_oracle_cfg(kw)
    sd = m.state_dict()
    with torch.no_grad():
        want = DiffusionInferer(R).sample(input_noise=noise, diffusion_model=m, scheduler=R, verbose=False)
        got = O.diffusion_sample(lambda x, t, c: O.unet_forward(sd, cfg, x, t, context=c), M, noise)
    _close(got, want, 1e-4)


@pytest.mark.parametrize("ptype,clip", [("epsilon", True), ("v_prediction", False), ("sample", True)])
def test_get_likelihood(ref, monkeypatch, ptype, clip):
    """Oracle get_likelihood vs the reference's DiffusionInferer.get_likelihood (inferer.py:145-277), same noise."""
    nets, sch = ref
    from generative.inferers import DiffusionInferer
    kw = G.UNET_CASES["unet2d_attn"]
    torch.manual_seed(0)
    m = G.randomize_zero_params(nets.DiffusionModelUNet(**kw)).eval()
    skw = dict(num_train_timesteps=12, prediction_type=ptype, clip_sample=clip)
    R, M = sch.DDPMScheduler(**skw), O.DDPMOracle(**skw)
    R.set_timesteps(12)
    M.set_timesteps(12)
    torch.manual_seed(5)
    x = torch.rand(2, 1, 16, 16) * 2 - 1
    x[0, 0, 0, :4] = torch.tensor([-1.0, 1.0, -0.9995, 0.9995])          # both edge bins of the discretised decoder
    noise = torch.randn(2, 1, 16, 16)
    monkeypatch.setattr(torch, "randn_like", lambda t: noise.clone())
    cfg, sd = G.unet_oracle_cfg(kw), m.state_dict()
    with torch.no_grad():
        want = DiffusionInferer(R).get_likelihood(inputs=x, diffusion_model=m, scheduler=R, verbose=False)
        got = O.get_likelihood(lambda xx, t, c: O.unet_forward(sd, cfg, xx, t, context=c), M, x, noise)
    _close(got, want, 1e-4)


@pytest.mark.parametrize("name", list(G.SPADE_UNET_CASES))
def test_spade_unet(ref, name):
    """SPADEDiffusionModelUNet (spade_diffusion_model_unet.py:612-912): SPADE norms in the up path, incl. the
    InstanceNorm that monai's Convolution puts on mlp_gamma / mlp_beta."""
    nets, _ = ref
    kw, inp = G.SPADE_UNET_CASES[name], G.SPADE_UNET_INPUTS[name]
    torch.manual_seed(0)
    m = G.randomize_zero_params(nets.SPADEDiffusionModelUNet(**kw)).eval()
    torch.manual_seed(1)
    x = torch.randn(*inp["shape"])
    t = torch.randint(0, 1000, (inp["shape"][0],)).long()
    seg = G.seg_onehot(inp["seg"])
    ctx = torch.randn(*inp["context"]) if "context" in inp else None
    with torch.no_grad():
        want = m(x, t, seg, context=ctx)
        got = O.unet_forward(m.state_dict(), G.unet_oracle_cfg(kw), x, t, context=ctx, seg=seg)
    _close(got, want)


@pytest.mark.parametrize("name", list(G.SPADE_AEKL_CASES))
def test_spade_autoencoderkl(ref, name):
    """SPADEAutoencoderKL.encode / decode(z, seg) (spade_autoencoderkl.py:425-469)."""
    nets, _ = ref
    kw = G.SPADE_AEKL_CASES[name]
    torch.manual_seed(0)
    m = nets.SPADEAutoencoderKL(**kw).eval()
    torch.manual_seed(1)
    x = torch.randn(*G.SPADE_AEKL_INPUTS[name])
    seg = G.seg_onehot((x.shape[0], kw["label_nc"], *x.shape[2:]))
    with torch.no_grad():
        mu, sigma = m.encode(x)
        rec = m.decode(mu, seg)
        gmu, gsigma = O.autoencoderkl_encode(m.state_dict(), G.aekl_oracle_cfg(kw), x)
        grec = O.autoencoderkl_decode(m.state_dict(), G.aekl_oracle_cfg(kw), gmu, seg=seg)
    _close(gmu, mu)
    _close(gsigma, sigma)
    _close(grec, rec)


def test_ordering(ref):
    """utils/ordering.py: raster / s-curve scans of transposed, rotated and reflected index grids."""
    import numpy as np
    from generative.utils.ordering import Ordering
    for kw in (dict(ordering_type="s_curve", spatial_dims=2, dimensions=(1, 3, 4), reflected_spatial_dims=(True, False)),
               dict(ordering_type="raster_scan", spatial_dims=3, dimensions=(1, 2, 3, 4), transpositions_axes=((2, 0, 1),),
                    rot90_axes=((0, 1),), reflected_spatial_dims=(False, True, True)),
               dict(ordering_type="s_curve", spatial_dims=3, dimensions=(1, 3, 2, 4),
                    transformation_order=("reflect", "transpose", "rotate_90"),
                    reflected_spatial_dims=(True, True, False), transpositions_axes=((1, 0, 2),))):
        assert np.array_equal(Ordering(**kw).get_sequence_ordering(), O.sequence_ordering(**kw))


@pytest.mark.parametrize("cross", [False, True])
def test_transformer(ref, cross):
    """DecoderOnlyTransformer.forward (nets/transformer.py:96-106)."""
    nets, _ = ref
    torch.manual_seed(0)
    m = nets.DecoderOnlyTransformer(num_tokens=11, max_seq_len=16, attn_layers_dim=32, attn_layers_depth=2,
                                    attn_layers_heads=4, with_cross_attention=cross).eval()
    x = torch.randint(0, 11, (2, 9))
    ctx = torch.randn(2, 3, 32) if cross else None
    with torch.no_grad():
        _close(O.transformer_forward(m.state_dict(), 4, x, ctx), m(x, context=ctx))


def test_transformer_greedy_sampling(ref):
    """VQVAETransformerInferer.sample with top_k = 1 (inferer.py:1183-1245) against the oracle's greedy loop, incl.
    the sliding window once the sequence outgrows max_seq_len."""
    nets, _ = ref
    from generative.inferers import VQVAETransformerInferer
    from generative.utils.ordering import Ordering
    torch.manual_seed(0)
    vq = nets.VQVAE(**G.VQVAE_CASES["vqvae2d"]).eval()
    K = G.VQVAE_CASES["vqvae2d"]["num_embeddings"]
    tr = nets.DecoderOnlyTransformer(num_tokens=K + 1, max_seq_len=10, attn_layers_dim=32, attn_layers_depth=2,
                                     attn_layers_heads=4).eval()
    ordering = Ordering("raster_scan", 2, (1, 4, 4))
    start = torch.full((2, 1), K)
    with torch.no_grad():
        want = VQVAETransformerInferer().sample((4, 4), start, vq, tr, ordering, top_k=1, verbose=False)
        seq = O.transformer_sample_greedy(tr.state_dict(), 4, 10, K, 16, 2)
        seq = seq[:, ordering.get_revert_sequence_ordering()].reshape(2, 4, 4)
        sd = vq.state_dict()
        emb = O.vq_embed(sd["quantizer.quantizer.embedding.weight"], seq)
        got = O.vqvae_decode(sd, G.vqvae_oracle_cfg(G.VQVAE_CASES["vqvae2d"]), emb)
    _close(got, want)


def test_pndm_restart_mid_runge_kutta(ref):
    """The reference keeps a stale Runge-Kutta accumulator when a loop is aborted mid-cycle and restarted on the same
    scheduler (``+=`` at phase 0, pndm.py:208; set_timesteps resets ets / counter only): the oracle follows it."""
    _, sch = ref
    torch.manual_seed(0)
    x = torch.randn(1, 1, 8, 8)
    outs = [torch.randn(1, 1, 8, 8) for _ in range(8)]

    def drive(s):
        s.set_timesteps(4)
        y = x
        for t, e in zip(s.timesteps[:2], outs):
            y, _ = s.step(e.clone(), int(t), y)
        s.set_timesteps(4)
        y = x
        for t, e in zip(s.timesteps[:8], outs):
            y, _ = s.step(e.clone(), int(t), y)
        return y
    want = drive(sch.PNDMScheduler(num_train_timesteps=1000, skip_prk_steps=False))
    got = drive(O.PNDMOracle(num_train_timesteps=1000, skip_prk_steps=False))
    fresh = sch.PNDMScheduler(num_train_timesteps=1000, skip_prk_steps=False)
    fresh.set_timesteps(4)
    y = x
    for t, e in zip(fresh.timesteps[:8], outs):
        y, _ = fresh.step(e.clone(), int(t), y)
    assert not torch.allclose(want, y)          # the stale sum really is inherited by the reference
    _close(got, want, 1e-6)
