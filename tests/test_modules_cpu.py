"""End-to-end host logic on CPU: the real modules / schedulers / inferers of generativemodels_b200 driven through the
test-only CPU stand-in for the C-ABI (tests/cpu_backend.py) and compared with the golden vectors of the unmodified
reference and with the oracle.  Verifies everything above the kernels: module trees and state_dict keys, weight
packing, virtual concat, epilogue wiring, attention plumbing, scheduler coefficient maths, inferer loops."""
from pathlib import Path

import pytest
import torch

from oracle import torch_oracle as O
from tests import cpu_backend
from tests.golden import configs as G

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(autouse=True)
def _cpu_backend(monkeypatch):
    cpu_backend.install(monkeypatch)


def load(name):
    return torch.load(GOLD / f"{name}.pt", weights_only=False)


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def nets():
    import generativemodels_b200.networks.nets as N
    return N


def test_unet_golden_cpu():
    for name in ("g_unet2d", "g_unet3d_cross"):
        fx = load(name)
        m = nets().DiffusionModelUNet(**fx["kwargs"]).eval()
        assert set(m.state_dict().keys()) == set(fx["state_dict"].keys())
        m.load_state_dict(fx["state_dict"])
        y = m(fx["x"], fx["t"], context=fx["context"])
        assert y.shape == fx["y"].shape and rel(y, fx["y"]) < 2e-2, (name, rel(y, fx["y"]))


@pytest.mark.parametrize("name", list(G.UNET_CASES))
def test_unet_vs_oracle_cpu(name):
    kw = G.UNET_CASES[name]
    torch.manual_seed(0)
    m = G.randomize_zero_params(nets().DiffusionModelUNet(**kw)).eval()
    inp = G.UNET_INPUTS[name]
    torch.manual_seed(1)
    x = torch.randn(inp["shape"])
    t = torch.Tensor((500,))
    ctx = torch.randn(inp["context"]) if "context" in inp else None
    cls = torch.randint(0, 5, (inp["shape"][0],)) if inp.get("classes") else None
    want = O.unet_forward(m.state_dict(), G.unet_oracle_cfg(kw), x, t, context=ctx, class_labels=cls)
    got = m(x, t, context=ctx, class_labels=cls)
    assert rel(got, want) < 2e-2, rel(got, want)


def test_time_embedding_projections_are_one_launch_cpu(monkeypatch):
    """All ResnetBlocks' time_emb_proj(silu(emb)) come out of one GEMV over the concatenated weights and reach the
    conv epilogues as strided column slices: same result as one launch per block, and the cache follows the weights."""
    from generativemodels_b200 import _lib
    import generativemodels_b200.networks.nets.diffusion_model_unet as U
    kw = G.UNET_CASES["unet2d_updown_class"]
    torch.manual_seed(0)
    m = G.randomize_zero_params(nets().DiffusionModelUNet(**kw)).eval()
    x, t, cls = torch.randn(2, 2, 16, 16), torch.tensor([10, 700]), torch.tensor([1, 4])
    lib = _lib.require_device()
    calls = []
    orig = lib.b200_small_linear
    monkeypatch.setattr(lib, "b200_small_linear", lambda *a: (calls.append(a[5]), orig(*a))[1])
    batched = m(x, t, class_labels=cls)
    n_res = sum(isinstance(b, U.ResnetBlock) for b in m.modules())
    assert n_res >= 5 and len(calls) == 3 and calls[-1] == sum(
        b.out_channels for b in m.modules() if isinstance(b, U.ResnetBlock))
    calls.clear()
    real = U.project_time_embedding
    monkeypatch.setattr(U, "project_time_embedding", lambda root, emb: U.TimeEmb(emb, {}))
    per_block = m(x, t, class_labels=cls)
    assert len(calls) == 2 + n_res
    assert rel(batched, per_block) < 1e-6
    monkeypatch.setattr(U, "project_time_embedding", real)
    with torch.no_grad():                                   # in-place weight update: the concatenation is rebuilt
        blk = next(b for b in m.modules() if isinstance(b, U.ResnetBlock))
        blk.time_emb_proj.bias.add_(1.0)
    want = O.unet_forward(m.state_dict(), G.unet_oracle_cfg(kw), x, t, class_labels=cls)
    assert rel(m(x, t, class_labels=cls), want) < 2e-2


def test_small_groupnorm_routing_cpu(monkeypatch):
    """ops._GN_SMALL (on by default since round 2): small GroupNorms go through ONE b200_groupnorm_fused call — also over the
    virtual concat of the up path — and the network output is unchanged."""
    from generativemodels_b200 import _lib, ops
    assert ops._GN_SMALL is True
    monkeypatch.setattr(ops, "_GN_SMALL", False)
    kw = G.UNET_CASES["unet3d_attn"]
    torch.manual_seed(0)
    m = G.randomize_zero_params(nets().DiffusionModelUNet(**kw)).eval()
    x, t = torch.randn(1, 1, 8, 12, 8), torch.Tensor((300,))
    want = m(x, t)
    lib = _lib.require_device()
    calls = {"fused": 0, "stats": 0}
    fused, stats = lib.b200_groupnorm_fused, lib.b200_groupnorm_stats
    monkeypatch.setattr(lib, "b200_groupnorm_fused", lambda *a: (calls.__setitem__("fused", calls["fused"] + 1), fused(*a))[1])
    monkeypatch.setattr(ops, "_GN_SMALL", True)
    monkeypatch.setattr(ops, "_GN_SMALL_MAX_ELEMS", 8 * 12 * 8 * 4)       # level 0 of the up path (8 ch / group) stays two-phase
    got = m(x, t)
    monkeypatch.setattr(lib, "b200_groupnorm_stats", lambda *a: (calls.__setitem__("stats", calls["stats"] + 1), stats(*a))[1])
    m(x, t)
    assert calls["fused"] >= 2 * 10 and calls["stats"] >= 1
    assert rel(got, want) < 1e-6


def test_samplers_golden_cpu():
    from generativemodels_b200.inferers import DiffusionInferer
    from generativemodels_b200.networks.schedulers import DDIMScheduler, DDPMScheduler, PNDMScheduler
    fx = load("g_unet2d")
    m = nets().DiffusionModelUNet(**fx["kwargs"]).eval()
    m.load_state_dict(fx["state_dict"])
    s = DDIMScheduler(**fx["ddim_kwargs"])
    s.set_timesteps(fx["ddim_steps"])
    assert rel(DiffusionInferer(s).sample(fx["noise"], m, s, verbose=False), fx["ddim_sample"]) < 5e-2
    p = PNDMScheduler(**fx["pndm_kwargs"])
    p.set_timesteps(fx["pndm_steps"])
    assert rel(DiffusionInferer(p).sample(fx["noise"], m, p, verbose=False), fx["pndm_sample"]) < 5e-2
    d = DDPMScheduler(**fx["ddpm_kwargs"])
    d.set_timesteps(fx["ddpm_steps"])
    torch.manual_seed(fx["ddpm_seed"])
    assert rel(DiffusionInferer(d).sample(fx["noise"], m, d, verbose=False), fx["ddpm_sample"]) < 5e-2


def test_controlnet_golden_cpu():
    fx = load("g_controlnet")
    cn = nets().ControlNet(**fx["kwargs"]).eval()
    assert set(cn.state_dict().keys()) == set(fx["cn_state_dict"].keys())
    cn.load_state_dict(fx["cn_state_dict"])
    un = nets().DiffusionModelUNet(**fx["unet_kwargs"]).eval()
    un.load_state_dict(fx["unet_state_dict"])
    down, mid = cn(fx["x"], fx["t"], fx["cond"], conditioning_scale=fx["scale"], context=fx["context"])
    for a, b in zip(down, fx["down"]):
        assert rel(a, b) < 2e-2
    assert rel(mid, fx["mid"]) < 2e-2
    y = un(fx["x"], fx["t"], context=fx["context"], down_block_additional_residuals=down,
           mid_block_additional_residual=mid)
    assert rel(y, fx["y"]) < 2e-2


def test_autoencoderkl_ldm_golden_cpu():
    from generativemodels_b200.inferers import LatentDiffusionInferer
    from generativemodels_b200.networks.schedulers import DDIMScheduler
    fx = load("g_aekl2d")
    ae = nets().AutoencoderKL(**fx["kwargs"]).eval()
    assert set(ae.state_dict().keys()) == set(fx["state_dict"].keys())
    ae.load_state_dict(fx["state_dict"])
    mu, sigma = ae.encode(fx["x"])
    assert rel(mu, fx["mu"]) < 2e-2 and rel(sigma, fx["sigma"]) < 2e-2
    assert rel(ae.decode(fx["mu"]), fx["rec"]) < 2e-2
    un = nets().DiffusionModelUNet(**fx["latent_unet_kwargs"]).eval()
    un.load_state_dict(fx["latent_unet_state_dict"])
    s = DDIMScheduler(**fx["ddim_kwargs"])
    s.set_timesteps(fx["ddim_steps"])
    img = LatentDiffusionInferer(s, scale_factor=fx["scale_factor"]).sample(fx["latent_noise"], ae, un, s, verbose=False)
    assert rel(img, fx["ldm_sample"]) < 5e-2


@pytest.mark.parametrize("name", list(G.AEKL_CASES))
def test_autoencoderkl_vs_oracle_cpu(name):
    kw = G.AEKL_CASES[name]
    torch.manual_seed(0)
    m = nets().AutoencoderKL(**kw).eval()
    torch.manual_seed(3)
    x = torch.randn(G.AEKL_INPUTS[name])
    cfg = G.aekl_oracle_cfg(kw)
    mu_w, sig_w = O.autoencoderkl_encode(m.state_dict(), cfg, x)
    mu, sig = m.encode(x)
    assert rel(mu, mu_w) < 2e-2 and rel(sig, sig_w) < 2e-2
    assert rel(m.decode(mu_w), O.autoencoderkl_decode(m.state_dict(), cfg, mu_w)) < 2e-2


def test_vqvae_golden_cpu():
    fx = load("g_vqvae3d")
    m = nets().VQVAE(**fx["kwargs"]).eval()
    assert set(m.state_dict().keys()) == set(fx["state_dict"].keys())
    m.load_state_dict(fx["state_dict"])
    assert rel(m.encode(fx["x"]), fx["z"]) < 2e-2
    q, loss, idx = m.quantizer.quantizer(fx["z"])
    assert torch.equal(idx, fx["idx"])
    assert rel(m.decode_samples(fx["idx"]), fx["dec_from_idx"]) < 2e-2
    rec, loss = m(fx["x"])
    assert rec.shape == fx["rec"].shape and rel(rec, fx["rec"]) < 8e-2
    assert m.index_quantize(fx["x"]).shape == fx["idx"].shape
    assert m.decode_stage_2_outputs(fx["z"]).shape == fx["rec"].shape
    assert m.encode_stage_2_inputs(fx["x"]).shape == fx["z"].shape


@pytest.mark.parametrize("name", list(G.VQVAE_CASES))
def test_vqvae_vs_oracle_cpu(name):
    kw = G.VQVAE_CASES[name]
    torch.manual_seed(0)
    m = nets().VQVAE(**kw).eval()
    cfg = G.vqvae_oracle_cfg(kw)
    torch.manual_seed(4)
    x = torch.rand(G.VQVAE_INPUTS[name])
    z_w = O.vqvae_encode(m.state_dict(), cfg, x)
    assert rel(m.encode(x), z_w) < 2e-2
    cb = m.state_dict()["quantizer.quantizer.embedding.weight"]
    q_w, loss_w, idx_w, perp_w = O.vq_forward(cb, z_w)
    q, loss, idx = m.quantizer.quantizer(z_w)
    assert torch.equal(idx, idx_w) and torch.equal(q, q_w)
    assert abs(float(loss) - float(loss_w)) < 1e-6
    m.quantizer(z_w)
    assert abs(float(m.quantizer.perplexity) - float(perp_w)) < 1e-4
    assert rel(m.decode(q_w), O.vqvae_decode(m.state_dict(), cfg, q_w)) < 2e-2


def test_vq_ema_training_composite():
    """The reference's known-answer test (tests/test_vector_quantizer.py:45-62): one train step with decay=0,
    epsilon=0 moves code 0 to its (shifted) input and leaves code 1 bit-identical."""
    from generativemodels_b200.networks.layers import EMAQuantizer
    torch.manual_seed(0)
    layer = EMAQuantizer(spatial_dims=2, num_embeddings=2, embedding_dim=2, epsilon=0, decay=0)
    w0, w1 = layer.embedding.weight[0].clone(), layer.embedding.weight[1].clone()
    x = torch.cat([(w0[None, :, None, None] + 0.001), w1[None, :, None, None]], dim=0)
    layer.train()
    layer(x)
    assert all(layer.embedding.weight[0] != w0)
    assert all(layer.embedding.weight[1] == w1)


def test_get_likelihood_vs_oracle_cpu(monkeypatch):
    """DiffusionInferer.get_likelihood (fused KL kernel per step) against the oracle's restatement, same noise draw."""
    from generativemodels_b200.inferers import DiffusionInferer
    from generativemodels_b200.networks.schedulers import DDIMScheduler, DDPMScheduler
    fx = load("g_unet2d")
    m = nets().DiffusionModelUNet(**fx["kwargs"]).eval()
    m.load_state_dict(fx["state_dict"])
    sd, cfg = fx["state_dict"], G.unet_oracle_cfg(fx["kwargs"])
    s, so = DDPMScheduler(num_train_timesteps=10), O.DDPMOracle(num_train_timesteps=10)
    s.set_timesteps(10)
    so.set_timesteps(10)
    torch.manual_seed(21)
    x = torch.rand(2, 1, 8, 8) * 2 - 1
    noise = torch.randn(2, 1, 8, 8)
    monkeypatch.setattr(torch, "randn_like", lambda t: noise.clone())
    got, inter = DiffusionInferer(s).get_likelihood(x, m, s, save_intermediates=True, verbose=False)
    want = O.get_likelihood(lambda xx, tt, c: O.unet_forward(sd, cfg, xx, tt, context=c), so, x, noise)
    assert got.shape == (2,) and len(inter) == 10 and inter[0].shape == x.shape
    assert torch.allclose(got, want, rtol=5e-2, atol=1e-3), (got, want)
    with pytest.raises(NotImplementedError):
        d = DDIMScheduler(num_train_timesteps=10)
        DiffusionInferer(d).get_likelihood(x, m, d, verbose=False)


def test_likelihood_variants_host_logic_cpu():
    """Latent / ControlNet / ControlNet-latent get_likelihood plumbing (inferer.py:489-562, 710-853, 1041-1124):
    encode -> scale -> shared KL loop -> optional resampling of the KL maps, through the CPU stand-in."""
    from generativemodels_b200.inferers import (ControlNetDiffusionInferer, ControlNetLatentDiffusionInferer,
                                                LatentDiffusionInferer)
    from generativemodels_b200.networks.schedulers import DDPMScheduler
    ae = nets().AutoencoderKL(spatial_dims=2, in_channels=1, out_channels=1, num_channels=(4, 4), latent_channels=3,
                              attention_levels=[False, False], num_res_blocks=1, with_encoder_nonlocal_attn=False,
                              with_decoder_nonlocal_attn=False, norm_num_groups=4).eval()
    base = dict(spatial_dims=2, in_channels=3, num_channels=[4, 4], norm_num_groups=4, attention_levels=[False, False],
                num_res_blocks=1, num_head_channels=4)
    un = G.randomize_zero_params(nets().DiffusionModelUNet(out_channels=3, **base)).eval()
    cn = G.randomize_zero_params(nets().ControlNet(conditioning_embedding_num_channels=[16],
                                                   conditioning_embedding_in_channels=1, **base)).eval()
    s = DDPMScheduler(num_train_timesteps=4)
    s.set_timesteps(4)
    img, mask = torch.randn(1, 1, 8, 8), torch.randn(1, 1, 8, 8)
    lik, inter = LatentDiffusionInferer(s, 1.0).get_likelihood(img, ae, un, s, save_intermediates=True, verbose=False)
    assert lik.shape == (1,) and len(inter) == 4 and inter[0].shape == (1, 3, 4, 4)
    lik, inter = LatentDiffusionInferer(s, 1.0).get_likelihood(img, ae, un, s, save_intermediates=True, verbose=False,
                                                               resample_latent_likelihoods=True)
    assert inter[0].shape == (1, 3, 8, 8)
    with pytest.raises(ValueError):
        LatentDiffusionInferer(s, 1.0).get_likelihood(img, ae, un, s, resample_latent_likelihoods=True,
                                                      resample_interpolation_mode="cubic", verbose=False)
    lik, inter = ControlNetLatentDiffusionInferer(s, 1.0).get_likelihood(
        img, ae, un, cn, mask, s, save_intermediates=True, resample_latent_likelihoods=True, verbose=False)
    assert torch.isfinite(lik).all() and inter[0].shape == (1, 3, 8, 8)
    lat = torch.randn(1, 3, 4, 4)
    lik = ControlNetDiffusionInferer(s).get_likelihood(lat, un, cn, torch.randn(1, 1, 4, 4), s, verbose=False)
    assert lik.shape == (1,) and torch.isfinite(lik).all()


def test_spade_golden_cpu():
    """SPADEDiffusionModelUNet / SPADEAutoencoderKL host logic (strict state_dict load, fused gamma|beta conv,
    InstanceNorm table, segmentation pyramid) against the reference's golden outputs."""
    fx = load("g_spade_unet2d")
    m = nets().SPADEDiffusionModelUNet(**fx["kwargs"]).eval()
    m.load_state_dict(fx["state_dict"])              # strict: same keys as the reference
    assert rel(m(fx["x"], fx["t"], fx["seg"]), fx["y"]) < 2e-2
    with pytest.raises(ValueError):
        m.up_blocks[0].resnets[0]([ops_cl(fx["x"], 16), ops_cl(fx["x"], 16)], torch.zeros(2, 32), None)
    fx = load("g_spade_aekl2d")
    ae = nets().SPADEAutoencoderKL(**fx["kwargs"]).eval()
    ae.load_state_dict(fx["state_dict"])
    mu, sigma = ae.encode(fx["x"])
    assert rel(mu, fx["mu"]) < 2e-2 and rel(sigma, fx["sigma"]) < 2e-2
    assert rel(ae.decode(fx["mu"], fx["seg"]), fx["rec"]) < 2e-2
    assert rel(ae.reconstruct(fx["x"], fx["seg"]), fx["rec"]) < 3e-2
    with pytest.raises(ValueError):
        nets().SPADEDiffusionModelUNet(spatial_dims=2, in_channels=1, out_channels=1, label_nc=3,
                                       num_channels=(8, 16), attention_levels=(False,), norm_num_groups=4)


def ops_cl(x, C_):
    from generativemodels_b200 import ops
    return ops.CL(torch.zeros(x.shape[0], 1, x.shape[2], x.shape[3], C_, dtype=ops.H16), C_, 2)


def _transformer_pair(cross=False, max_seq_len=16, tokens=11):
    torch.manual_seed(0)
    m = nets().DecoderOnlyTransformer(num_tokens=tokens, max_seq_len=max_seq_len, attn_layers_dim=32,
                                      attn_layers_depth=2, attn_layers_heads=4, with_cross_attention=cross).eval()
    return m, {k: v.clone() for k, v in m.state_dict().items()}


@pytest.mark.parametrize("cross", [False, True])
def test_transformer_forward_and_cache_cpu(monkeypatch, cross):
    """DecoderOnlyTransformer host logic (embedding kernel call, fused-GELU MLP, causal attention parameters, logits
    slicing) against the oracle, and the key/value-cache ``step`` against the full forward — same rows, any chunking."""
    import generativemodels_b200.networks.nets.transformer as T
    monkeypatch.setattr(T, "require_cuda", lambda x, m: None)
    m, sd = _transformer_pair(cross)
    assert set(sd) == set(m.state_dict()) and "blocks.0.attn.causal_mask" in sd
    x = torch.randint(0, 11, (2, 9))
    ctx = torch.randn(2, 3, 32) if cross else None
    want = O.transformer_forward(sd, 4, x, ctx)
    got = m(x, context=ctx)
    assert got.shape == (2, 9, 11) and rel(got, want) < 2e-2
    cache = m.new_cache(2, x.device, ctx)
    inc = torch.cat([m.step(x[:, :4], cache)] + [m.step(x[:, i:i + 1], cache) for i in range(4, 9)], 1)
    assert torch.equal(inc, got)
    with pytest.raises(IndexError):
        m.step(torch.randint(0, 11, (2, 8)), cache)
    # the position-in-device-memory form of the single-token step (what the CUDA graph captures)
    dyn = m.new_cache(2, x.device, ctx)
    rows = [m._step_dyn(x[:, i:i + 1].contiguous(), dyn) for i in range(9)]
    assert int(dyn.pos_dev) == 9 and rel(torch.cat(rows, 1), got) < 1e-2      # GEMV path: other rounding points
    if cross:
        with pytest.raises(ValueError):
            m(x)


def test_vqvae_transformer_inferer_cpu(monkeypatch):
    """VQVAETransformerInferer: greedy sampling (top_k = 1; key/value cache, then the sliding window once the sequence
    outgrows max_seq_len) against the oracle's loop, teacher-forced __call__, get_likelihood and the Ordering class."""
    import numpy as np
    import generativemodels_b200.networks.nets.transformer as T
    from generativemodels_b200.inferers import VQVAETransformerInferer
    from generativemodels_b200.utils.ordering import Ordering
    monkeypatch.setattr(T, "require_cuda", lambda x, m: None)
    kw = G.VQVAE_CASES["vqvae2d"]
    K = kw["num_embeddings"]
    torch.manual_seed(0)
    vq = nets().VQVAE(**kw).eval()
    vsd = {k: v.clone() for k, v in vq.state_dict().items()}
    tr, tsd = _transformer_pair(max_seq_len=10, tokens=K + 1)
    ordering = Ordering("s_curve", 2, (1, 4, 4), reflected_spatial_dims=(True, False))
    assert np.array_equal(ordering.get_sequence_ordering(),
                          O.sequence_ordering("s_curve", 2, (1, 4, 4), reflected_spatial_dims=(True, False)))
    inf = VQVAETransformerInferer()
    start = torch.full((2, 1), K)
    got = inf.sample((4, 4), start, vq, tr, ordering, top_k=1, verbose=False)
    seq = O.transformer_sample_greedy(tsd, 4, 10, K, 16, 2)
    seq = seq[:, ordering.get_revert_sequence_ordering()].reshape(2, 4, 4)
    want = O.vqvae_decode(vsd, G.vqvae_oracle_cfg(kw), O.vq_embed(vsd["quantizer.quantizer.embedding.weight"], seq))
    assert got.shape == want.shape and rel(got, want) < 3e-2
    # teacher forcing and likelihood: shapes and agreement with the oracle's logits
    x = torch.randn(2, 1, 16, 16)
    pred, target, sdim = inf(x, vq, tr, ordering, return_latent=True)
    assert sdim == (4, 4) and pred.shape == (2, 10, K + 1) and target.shape == (2, 10)
    ll = inf.get_likelihood(x, vq, tr, ordering)
    assert ll.shape == (2, 4, 4) and torch.isfinite(ll).all() and (ll <= 0).all()
    ll_up = inf.get_likelihood(x, vq, tr, ordering, resample_latent_likelihoods=True)
    assert ll_up.shape == (2, 1, 16, 16)
    with pytest.raises(ValueError):
        Ordering("hilbert", 2, (1, 4, 4))


def test_rank1_rank3_reference_fixtures_cpu(monkeypatch):
    """get_likelihood and the transformer sampler against fixtures written by the unmodified reference."""
    import generativemodels_b200.networks.nets.transformer as T
    from tests import fixture_checks
    monkeypatch.setattr(T, "require_cuda", lambda x, m: None)
    fixture_checks.check_likelihood_fixture("cpu", monkeypatch)
    fixture_checks.check_transformer_fixture("cpu")


def test_c1_reference_fixture_cpu():
    """BASELINE.json configs[0] at its real size through the modules on the CPU stand-in (~25 s of numpy GEMMs)."""
    from tests import fixture_checks
    fixture_checks.check_c1_fixture("cpu")


def test_c2_reference_fixture_cpu():
    """BASELINE.json configs[1] at its real size on the CPU stand-in: two probe steps of the reference trajectory (the
    benign first step and the ill-conditioned t = 500 one) and the 256 x 256 decoder."""
    from tests import fixture_checks
    report = fixture_checks.check_c2_fixture("cpu", probes=(0, 24))
    assert report[24][0] > report[0][0]          # the peaked-softmax probe is the harder one, as documented


def test_pndm_restart_mid_runge_kutta_cpu():
    """A sampling loop aborted inside a Runge-Kutta cycle and restarted on the same scheduler: the reference keeps the
    stale accumulator (``+=`` at phase 0, pndm.py:208; set_timesteps resets ets / counter only, 160-161) — the
    round-1 advisor read it as an assignment; the reference's code is the contract, so the call sequence must give the
    oracle's (= the reference's) numbers, stale sum included."""
    from generativemodels_b200.networks.schedulers import PNDMScheduler
    torch.manual_seed(0)
    x = torch.randn(1, 1, 8, 8)
    outs = [torch.randn(1, 1, 8, 8) for _ in range(8)]

    def drive(s):
        s.set_timesteps(4)
        y = x
        for t, e in zip(s.timesteps[:2], outs):      # abort after phases 0 and 1 of the first cycle
            y, _ = s.step(e, int(t), y)
        s.set_timesteps(4)
        y = x
        for t, e in zip(s.timesteps[:8], outs):
            y, _ = s.step(e, int(t), y)
        return y
    got = drive(PNDMScheduler(num_train_timesteps=1000, skip_prk_steps=False))
    want = drive(O.PNDMOracle(num_train_timesteps=1000, skip_prk_steps=False))
    assert rel(got, want) < 1e-5, rel(got, want)


def test_invalidate_packed_cpu():
    import generativemodels_b200 as B
    kw = G.UNET_CASES["unet2d_attn"]
    torch.manual_seed(0)
    m = G.randomize_zero_params(nets().DiffusionModelUNet(**kw)).eval()
    x, t = torch.randn(1, 1, 16, 16), torch.Tensor((500,))
    y0 = m(x, t)
    with torch.no_grad():
        m.conv_in.conv.weight.data.mul_(2.0)          # invisible to the version-keyed cache ...
    assert torch.equal(m(x, t), y0)
    B.invalidate_packed(m)                               # ... until the caches are dropped
    y1 = m(x, t)
    want = O.unet_forward(m.state_dict(), G.unet_oracle_cfg(kw), x, t)
    assert rel(y1, want) < 2e-2 and not torch.equal(y0, y1)


@pytest.mark.parametrize("act", ["tanh", "sigmoid", "relu"])
def test_vqvae_output_act_cpu(act):
    """VQVAE(output_act=...) (vqvae.py:263-264): the activation rides in the last transposed convolution's epilogue."""
    kw = dict(G.VQVAE_CASES["vqvae2d"], output_act=act)
    torch.manual_seed(0)
    m = nets().VQVAE(**kw).eval()
    cfg = G.vqvae_oracle_cfg(kw)
    torch.manual_seed(4)
    z = torch.randn(2, kw["embedding_dim"], 8, 8)
    assert rel(m.decode(z), O.vqvae_decode(m.state_dict(), cfg, z)) < 2e-2
    with pytest.raises(NotImplementedError):
        nets().VQVAE(**dict(kw, output_act="mish"))
