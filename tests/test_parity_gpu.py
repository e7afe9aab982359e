"""Parity of the CUDA path (through the reference-facing nn.Module / Inferer API) against
  (1) the committed golden vectors = outputs of the UNMODIFIED reference (tests/golden/*.pt), and
  (2) the CPU oracle (oracle/torch_oracle.py) on seeded models at sizes it finishes in seconds.

Tolerances (stated once, tests/fixture_checks.py and DESIGN.md section 3): the kernels multiply 16-bit operands (fp16 by
default) into fp32 accumulators with 16-bit activations between ops, the checker is fp32 end to end, so a network
forward is held to a relative L2 error of TOL_REL = 2e-2 AND a largest pointwise deviation of TOL_MAX = 4e-2 of the
largest reference magnitude; a multi-step sampler trajectory to TOL_TRAJ = 5e-2 (max-abs 1e-1); scheduler steps are
fp32 elementwise and held to 1e-5; VQ indices are bit-exact except at fp32 near-ties (gap between best and second-best
code below 1e-4 relative), which are counted.
"""
import math
from pathlib import Path

import pytest
import torch

from oracle import torch_oracle as O
from tests.fixture_checks import TOL_MAX, TOL_REL, TOL_TRAJ, relmax
from tests.golden import configs as G

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
FWD_TOL, TRAJ_TOL = TOL_REL, TOL_TRAJ


def load(name):
    return torch.load(GOLD / f"{name}.pt", weights_only=False)


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def check(a, b, tol=FWD_TOL, what="", tol_max=None):
    """relative L2 < tol and normalised max-abs < tol_max (default 2 x tol: TOL_MAX for forwards)."""
    assert tuple(a.shape) == tuple(b.shape), f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    r, m = rel(a, b), relmax(a.detach(), b.detach())
    tol_max = 2 * tol if tol_max is None else tol_max
    assert math.isfinite(r) and r < tol, f"{what}: relative L2 error {r:.3e} >= {tol}"
    assert math.isfinite(m) and m < tol_max, f"{what}: max-abs error {m:.3e} of the reference's peak >= {tol_max}"
    return r


def nets():
    import generativemodels_b200.networks.nets as N
    return N


def cuda(t):
    return None if t is None else t.cuda()


# ------------------------------------------------------------------------------------------------ golden (reference)
def test_unet_golden(cuda_device):
    for name in ("g_unet2d", "g_unet3d_cross"):
        fx = load(name)
        m = nets().DiffusionModelUNet(**fx["kwargs"]).cuda().eval()
        m.load_state_dict(fx["state_dict"])
        y = m(cuda(fx["x"]), cuda(fx["t"]), context=cuda(fx["context"]))
        check(y, fx["y"], FWD_TOL, name)


def test_samplers_golden(cuda_device):
    """DiffusionInferer.sample with DDIM / PNDM / DDPM against the reference's own trajectories."""
    from generativemodels_b200.inferers import DiffusionInferer
    from generativemodels_b200.networks.schedulers import DDIMScheduler, DDPMScheduler, PNDMScheduler
    fx = load("g_unet2d")
    m = nets().DiffusionModelUNet(**fx["kwargs"]).cuda().eval()
    m.load_state_dict(fx["state_dict"])
    noise = fx["noise"].cuda()
    s = DDIMScheduler(**fx["ddim_kwargs"])
    s.set_timesteps(fx["ddim_steps"])
    check(DiffusionInferer(s).sample(noise, m, s, verbose=False), fx["ddim_sample"], TRAJ_TOL, "ddim sample")
    p = PNDMScheduler(**fx["pndm_kwargs"])
    p.set_timesteps(fx["pndm_steps"])
    check(DiffusionInferer(p).sample(noise, m, p, verbose=False), fx["pndm_sample"], TRAJ_TOL, "pndm sample")
    d = DDPMScheduler(**fx["ddpm_kwargs"])
    d.set_timesteps(fx["ddpm_steps"])
    torch.manual_seed(fx["ddpm_seed"])     # same CPU noise stream as the reference run
    check(DiffusionInferer(d).sample(noise, m, d, verbose=False), fx["ddpm_sample"], TRAJ_TOL, "ddpm sample")


def test_controlnet_golden(cuda_device):
    from generativemodels_b200.inferers import ControlNetDiffusionInferer
    fx = load("g_controlnet")
    cn = nets().ControlNet(**fx["kwargs"]).cuda().eval()
    cn.load_state_dict(fx["cn_state_dict"])
    un = nets().DiffusionModelUNet(**fx["unet_kwargs"]).cuda().eval()
    un.load_state_dict(fx["unet_state_dict"])
    x, t, cond, ctx = (cuda(fx[k]) for k in ("x", "t", "cond", "context"))
    down, mid = cn(x, t, cond, conditioning_scale=fx["scale"], context=ctx)
    assert len(down) == len(fx["down"])
    for i, (a, b) in enumerate(zip(down, fx["down"])):
        check(a, b, FWD_TOL, f"controlnet down[{i}]")
    check(mid, fx["mid"], FWD_TOL, "controlnet mid")
    y = un(x, t, context=ctx, down_block_additional_residuals=down, mid_block_additional_residual=mid)
    check(y, fx["y"], FWD_TOL, "unet + controlnet residuals")


def test_autoencoderkl_and_ldm_golden(cuda_device):
    from generativemodels_b200.inferers import LatentDiffusionInferer
    from generativemodels_b200.networks.schedulers import DDIMScheduler
    fx = load("g_aekl2d")
    ae = nets().AutoencoderKL(**fx["kwargs"]).cuda().eval()
    ae.load_state_dict(fx["state_dict"])
    mu, sigma = ae.encode(fx["x"].cuda())
    check(mu, fx["mu"], FWD_TOL, "aekl mu")
    check(sigma, fx["sigma"], FWD_TOL, "aekl sigma")
    check(ae.decode(fx["mu"].cuda()), fx["rec"], FWD_TOL, "aekl decode")
    un = nets().DiffusionModelUNet(**fx["latent_unet_kwargs"]).cuda().eval()
    un.load_state_dict(fx["latent_unet_state_dict"])
    s = DDIMScheduler(**fx["ddim_kwargs"])
    s.set_timesteps(fx["ddim_steps"])
    inf = LatentDiffusionInferer(s, scale_factor=fx["scale_factor"])
    img = inf.sample(fx["latent_noise"].cuda(), ae, un, s, verbose=False)
    check(img, fx["ldm_sample"], TRAJ_TOL, "latent diffusion sample")


def _index_mismatch_report(idx, idx_ref, codebook, flat):
    """exact outside fp32 near-ties: returns (#hard mismatches, #near-tie flips)."""
    bad = (idx.cpu().flatten() != idx_ref.cpu().flatten())
    if not bad.any():
        return 0, 0
    margin = O.vq_index_margin(codebook.cpu(), flat.cpu())
    scale = (flat.cpu().double() ** 2).sum(1) + 1e-12
    near = (margin / scale) < 1e-4
    return int((bad & ~near).sum()), int((bad & near).sum())


def test_vqvae_golden(cuda_device):
    fx = load("g_vqvae3d")
    m = nets().VQVAE(**fx["kwargs"]).cuda().eval()
    m.load_state_dict(fx["state_dict"])
    x = fx["x"].cuda()
    z = m.encode(x)
    check(z, fx["z"], FWD_TOL, "vqvae encode")
    cb = fx["state_dict"]["quantizer.quantizer.embedding.weight"]
    # quantiser on the reference's own z: indices must be bit-exact
    q, loss, idx = m.quantizer.quantizer(fx["z"].cuda())
    flat = fx["z"].permute(0, 2, 3, 4, 1).reshape(-1, cb.shape[1])
    hard, near = _index_mismatch_report(idx, fx["idx"], cb, flat)
    assert hard == 0, f"{hard} VQ indices differ outside near-ties ({near} near-tie flips)"
    check(m.decode_samples(fx["idx"].cuda()), fx["dec_from_idx"], FWD_TOL, "vqvae decode_samples")
    rec, loss = m(x)
    # encoder runs in bf16, so a few vectors may cross a cell boundary: compare the decode statistically
    check(rec, fx["rec"], 6e-2, "vqvae forward")
    assert abs(float(loss) - float(fx["loss"])) < 0.05 * abs(float(fx["loss"])) + 1e-4


def test_vq_reference_known_answer(cuda_device):
    """tests/test_vector_quantizer.py:45-62 of the reference, through the CUDA kernel."""
    from generativemodels_b200.networks.layers import EMAQuantizer
    fx = load("g_vq_ema_case")
    q = EMAQuantizer(spatial_dims=2, num_embeddings=2, embedding_dim=2, epsilon=0, decay=0).cuda().eval()
    q.embedding.weight.data.copy_(fx["codebook"])
    _, _, idx = q(fx["x"].cuda())
    assert torch.equal(idx.cpu(), fx["idx"])


# ------------------------------------------------------------------------------------------------ oracle (seeded)
@pytest.mark.parametrize("name", list(G.UNET_CASES))
def test_unet_vs_oracle(cuda_device, name):
    kw = G.UNET_CASES[name]
    torch.manual_seed(0)
    m = G.randomize_zero_params(nets().DiffusionModelUNet(**kw)).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    inp = G.UNET_INPUTS[name]
    torch.manual_seed(1)
    x = torch.randn(inp["shape"])
    t = torch.randint(0, 1000, (inp["shape"][0],)).long()
    ctx = torch.randn(inp["context"]) if "context" in inp else None
    cls = torch.randint(0, 5, (inp["shape"][0],)) if inp.get("classes") else None
    want = O.unet_forward(sd, G.unet_oracle_cfg(kw), x, t, context=ctx, class_labels=cls)
    got = m.cuda()(x.cuda(), t.cuda(), context=cuda(ctx), class_labels=cuda(cls))
    check(got, want, FWD_TOL, name)
    # single float timestep broadcast over the batch, as DiffusionInferer.sample passes it (inferer.py:129)
    t1 = torch.Tensor((500,))
    want1 = O.unet_forward(sd, G.unet_oracle_cfg(kw), x, t1, context=ctx, class_labels=cls)
    got1 = m(x.cuda(), t1.cuda(), context=cuda(ctx), class_labels=cuda(cls))
    check(got1, want1, FWD_TOL, name + " (broadcast timestep)")


@pytest.mark.parametrize("name", list(G.AEKL_CASES))
def test_autoencoderkl_vs_oracle(cuda_device, name):
    kw = G.AEKL_CASES[name]
    torch.manual_seed(0)
    m = nets().AutoencoderKL(**kw).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    torch.manual_seed(3)
    x = torch.randn(G.AEKL_INPUTS[name])
    cfg = G.aekl_oracle_cfg(kw)
    mu_w, sig_w = O.autoencoderkl_encode(sd, cfg, x)
    m = m.cuda()
    mu, sig = m.encode(x.cuda())
    check(mu, mu_w, FWD_TOL, name + " mu")
    check(sig, sig_w, FWD_TOL, name + " sigma")
    check(m.decode(mu_w.cuda()), O.autoencoderkl_decode(sd, cfg, mu_w), FWD_TOL, name + " decode")


@pytest.mark.parametrize("name", list(G.VQVAE_CASES))
def test_vqvae_vs_oracle(cuda_device, name):
    kw = G.VQVAE_CASES[name]
    torch.manual_seed(0)
    m = nets().VQVAE(**kw).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    cfg = G.vqvae_oracle_cfg(kw)
    torch.manual_seed(4)
    x = torch.rand(G.VQVAE_INPUTS[name])
    z_w = O.vqvae_encode(sd, cfg, x)
    m = m.cuda()
    check(m.encode(x.cuda()), z_w, FWD_TOL, name + " encode")
    cb = sd["quantizer.quantizer.embedding.weight"]
    flat, idx_w = O.vq_quantize(cb, z_w)
    q, loss, idx = m.quantizer.quantizer(z_w.cuda())
    hard, near = _index_mismatch_report(idx, idx_w, cb, flat)
    assert hard == 0, f"{hard} VQ indices differ outside near-ties ({near} near-tie flips)"
    q_w, loss_w, _, perp_w = O.vq_forward(cb, z_w, kw.get("commitment_cost", 0.25))
    if near == 0:
        assert torch.equal(q.cpu(), q_w), "straight-through values must be bit-identical when indices agree"
    assert abs(float(loss) - float(loss_w)) < 1e-5 * max(1.0, abs(float(loss_w)))
    loss2, _ = m.quantizer(z_w.cuda())
    assert abs(float(m.quantizer.perplexity) - float(perp_w)) < 1e-3 * float(perp_w) + 1e-4
    check(m.decode(q_w.cuda()), O.vqvae_decode(sd, cfg, q_w), FWD_TOL, name + " decode")
    check(m.decode_stage_2_outputs(z_w.cuda()), O.vqvae_decode(sd, cfg, q_w), FWD_TOL, name + " stage-2 decode")


def test_vq_large_exact(cuda_device):
    """C4-sized quantiser problem (M = 32768, K = 256, D = 32): bit-exact indices vs the oracle outside near-ties."""
    from generativemodels_b200.networks.layers import EMAQuantizer
    torch.manual_seed(9)
    q = EMAQuantizer(spatial_dims=3, num_embeddings=256, embedding_dim=32).eval()
    cb = q.embedding.weight.detach().clone()
    z = torch.randn(1, 32, 32, 32, 32) * 0.7
    flat, idx_w = O.vq_quantize(cb, z)
    _, _, idx = q.cuda()(z.cuda())
    hard, near = _index_mismatch_report(idx, idx_w, cb, flat)
    assert hard == 0 and near <= 8, f"hard={hard} near-tie flips={near} of {idx_w.numel()}"


# ------------------------------------------------------------------------------------------------ schedulers
def _sched_pairs():
    from generativemodels_b200.networks import schedulers as S
    return {"ddim": (S.DDIMScheduler, O.DDIMOracle), "ddpm": (S.DDPMScheduler, O.DDPMOracle),
            "pndm": (S.PNDMScheduler, O.PNDMOracle)}


SCHED = [
    ("ddim", dict(num_train_timesteps=1000, schedule="linear_beta", beta_start=0.0015, beta_end=0.0195), 50),
    ("ddim", dict(num_train_timesteps=1000, schedule="scaled_linear_beta", beta_start=0.0005, beta_end=0.0195,
                  clip_sample=False), 50),
    ("ddim", dict(num_train_timesteps=100, prediction_type="v_prediction", set_alpha_to_one=False, steps_offset=1), 10),
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
def test_scheduler_trajectory(cuda_device, kind, kw, steps):
    """Same stand-in model output through the fused-kernel scheduler and the oracle: every step within 1e-5."""
    P, M = _sched_pairs()[kind]
    p, m = P(**kw), M(**kw)
    p.set_timesteps(steps)
    m.set_timesteps(steps)
    assert torch.equal(p.timesteps.cpu(), m.timesteps)
    assert torch.equal(p.alphas_cumprod, m.alphas_cumprod)
    torch.manual_seed(5)
    xm = torch.randn(2, 3, 8, 8)
    xp = xm.cuda()
    for t in m.timesteps:
        epsm = torch.tanh(xm * 0.7 + 0.01 * float(t))
        epsp = torch.tanh(xp * 0.7 + 0.01 * float(t))
        if kind == "ddpm":
            gm, gp = torch.Generator().manual_seed(int(t)), torch.Generator().manual_seed(int(t))
            xm, x0m = m.step(epsm, int(t), xm, generator=gm)
            xp, x0p = p.step(epsp, int(t), xp, generator=gp)
        else:
            xm, x0m = m.step(epsm, int(t), xm)
            xp, x0p = p.step(epsp, int(t), xp)
        err = (xp.cpu() - xm).abs().max().item()
        assert err < 1e-5 * max(1.0, xm.abs().max().item()), f"{kind} diverged at t={int(t)}: {err:.3e}"
        if x0m is not None and x0p is not None:
            assert (x0p.cpu() - x0m).abs().max().item() < 1e-5 * max(1.0, x0m.abs().max().item())


def test_ddim_eta_and_reversed_and_noise_ops(cuda_device):
    from generativemodels_b200.networks.schedulers import DDIMScheduler
    kw = dict(num_train_timesteps=1000)
    p, m = DDIMScheduler(**kw), O.DDIMOracle(**kw)
    p.set_timesteps(10)
    m.set_timesteps(10)
    torch.manual_seed(6)
    x, eps = torch.randn(2, 1, 8, 8), torch.randn(2, 1, 8, 8)
    want, _ = m.step(eps, 500, x, eta=0.5, generator=torch.Generator().manual_seed(3))
    got, _ = p.step(eps.cuda(), 500, x.cuda(), eta=0.5, generator=torch.Generator().manual_seed(3))
    assert (got.cpu() - want).abs().max().item() < 1e-5
    # add_noise / get_velocity (scheduler.py:169-200)
    t = torch.tensor([10, 900])
    acp = m.alphas_cumprod
    sa, sb = (acp[t] ** 0.5)[:, None, None, None], ((1 - acp[t]) ** 0.5)[:, None, None, None]
    assert (p.add_noise(x.cuda(), eps.cuda(), t).cpu() - (sa * x + sb * eps)).abs().max().item() < 1e-6
    assert (p.get_velocity(x.cuda(), eps.cuda(), t).cpu() - (sa * eps - sb * x)).abs().max().item() < 1e-6
    # reversed step: closed form (ddim.py:239-301)
    nt = 500 + 100
    a_t, a_n = acp[500], acp[nt]
    x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
    x0 = torch.clamp(x0, -1, 1)
    want_r = a_n ** 0.5 * x0 + (1 - a_n) ** 0.5 * eps
    got_r, got_x0 = p.reversed_step(eps.cuda(), 500, x.cuda())
    assert (got_r.cpu() - want_r).abs().max().item() < 1e-5
    assert (got_x0.cpu() - x0).abs().max().item() < 1e-5


# ------------------------------------------------------------------------------------------------ inferers vs oracle
def test_inferers_vs_oracle(cuda_device):
    from generativemodels_b200.inferers import ControlNetDiffusionInferer, DiffusionInferer
    from generativemodels_b200.networks.schedulers import DDIMScheduler
    # concat conditioning through DiffusionInferer
    kw = dict(spatial_dims=2, in_channels=3, out_channels=1, num_res_blocks=1, num_channels=(32, 64),
              attention_levels=(False, True), norm_num_groups=8, num_head_channels=(0, 64))
    torch.manual_seed(0)
    m = G.randomize_zero_params(nets().DiffusionModelUNet(**kw)).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    cfg = G.unet_oracle_cfg(kw)
    skw = dict(num_train_timesteps=1000, schedule="scaled_linear_beta", beta_start=0.0005, beta_end=0.0195,
               clip_sample=False)
    so, sp = O.DDIMOracle(**skw), DDIMScheduler(**skw)
    so.set_timesteps(4)
    sp.set_timesteps(4)
    torch.manual_seed(2)
    noise, cond = torch.randn(2, 1, 16, 16), torch.randn(2, 2, 16, 16)
    want = O.diffusion_sample(lambda x, t, c: O.unet_forward(sd, cfg, x, t, context=c), so, noise, cond, mode="concat")
    got = DiffusionInferer(sp).sample(noise.cuda(), m.cuda(), sp, conditioning=cond.cuda(), mode="concat", verbose=False)
    check(got, want, TRAJ_TOL, "DiffusionInferer concat")
    with pytest.raises(NotImplementedError):
        DiffusionInferer(sp).sample(noise.cuda(), m, sp, mode="nope", verbose=False)

    # ControlNet inferer (crossattn), 3 DDIM steps
    ckw = G.CONTROLNET_CASE
    torch.manual_seed(0)
    cn = G.randomize_zero_params(nets().ControlNet(**ckw)).eval()
    ukw = {k: v for k, v in ckw.items() if not k.startswith("conditioning_embedding")}
    un = G.randomize_zero_params(nets().DiffusionModelUNet(out_channels=3, **ukw)).eval()
    csd = {k: v.clone() for k, v in cn.state_dict().items()}
    usd = {k: v.clone() for k, v in un.state_dict().items()}
    ccfg = G.unet_oracle_cfg(ckw)
    so.set_timesteps(3)
    sp.set_timesteps(3)
    torch.manual_seed(3)
    noise, cnc, ctx = torch.randn(2, 3, 16, 16), torch.rand(2, 1, 16, 16), torch.randn(2, 1, 8)
    want = O.controlnet_sample(
        lambda x, t, c, d, mm: O.unet_forward(usd, ccfg, x, t, context=c, down_block_additional_residuals=d,
                                              mid_block_additional_residual=mm),
        lambda x, t, cc, c: O.controlnet_forward(csd, ccfg, x, t, cc, 1.0, c), so, noise, cnc, ctx)
    got = ControlNetDiffusionInferer(sp).sample(noise.cuda(), un.cuda(), cn.cuda(), cnc.cuda(), sp,
                                                conditioning=ctx.cuda(), verbose=False)
    check(got, want, TRAJ_TOL, "ControlNetDiffusionInferer")


@pytest.mark.parametrize("ptype,clip", [("epsilon", True), ("v_prediction", False)])
def test_get_likelihood_vs_oracle(cuda_device, monkeypatch, ptype, clip):
    """DiffusionInferer.get_likelihood (inferer.py:145-277): network forward + fused b200_ddpm_kl per timestep."""
    from generativemodels_b200.inferers import DiffusionInferer
    from generativemodels_b200.networks.schedulers import DDPMScheduler
    kw = G.UNET_CASES["unet2d_attn"]
    torch.manual_seed(0)
    m = G.randomize_zero_params(nets().DiffusionModelUNet(**kw)).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    cfg = G.unet_oracle_cfg(kw)
    skw = dict(num_train_timesteps=12, prediction_type=ptype, clip_sample=clip)
    so, sp = O.DDPMOracle(**skw), DDPMScheduler(**skw)
    so.set_timesteps(12)
    sp.set_timesteps(12)
    torch.manual_seed(5)
    x = torch.rand(2, 1, 16, 16) * 2 - 1
    x[0, 0, 0, :4] = torch.tensor([-1.0, 1.0, -0.9995, 0.9995])
    noise = torch.randn(2, 1, 16, 16)
    monkeypatch.setattr(torch, "randn_like", lambda t: noise.clone().to(t.device))
    want = O.get_likelihood(lambda xx, t, c: O.unet_forward(sd, cfg, xx, t, context=c), so, x, noise)
    got, inter = DiffusionInferer(sp).get_likelihood(x.cuda(), m.cuda(), sp, save_intermediates=True, verbose=False)
    assert len(inter) == 12 and inter[0].shape == x.shape
    # bf16 network output inside a KL whose late-timestep terms scale the mean error by 1/variance
    assert torch.allclose(got.cpu(), want, rtol=5e-2, atol=2e-3), (got, want)


def test_ddpm_kl_kernel_exact(cuda_device):
    """b200_ddpm_kl alone (fp32 in, fp32 out) against the oracle's closed forms, all three prediction types, t>0 / t=0."""
    import ctypes as C
    from generativemodels_b200 import _lib
    lib = _lib.require_device()
    torch.manual_seed(1)
    N, per = 3, 1000
    x0 = (torch.rand(N, per) * 2 - 1)
    x0[0, :4] = torch.tensor([-1.0, 1.0, -0.9995, 0.9995])
    xt, mo = torch.randn(N, per), torch.randn(N, per)
    for pt, name in ((_lib.PRED_EPSILON, "eps"), (_lib.PRED_SAMPLE, "sample"), (_lib.PRED_V, "v")):
        for t0 in (0, 1):
            c = _lib.KlCoef()
            c.sqrt_alpha_prod_t, c.sqrt_beta_prod_t, c.coef_x0, c.coef_xt = 0.9, 0.43589, 0.3, 0.68
            c.log_pred_var = c.log_post_var = -3.2
            c.bin_width, c.prediction_type, c.clip, c.is_t0 = 1.0 / 255, pt, 1, t0
            if pt == _lib.PRED_EPSILON:
                p0 = (xt - c.sqrt_beta_prod_t * mo) / c.sqrt_alpha_prod_t
            elif pt == _lib.PRED_SAMPLE:
                p0 = mo
            else:
                p0 = c.sqrt_alpha_prod_t * xt - c.sqrt_beta_prod_t * mo
            p0 = p0.clamp(-1, 1)
            pred, post = c.coef_x0 * p0 + c.coef_xt * xt, c.coef_x0 * x0 + c.coef_xt * xt
            if t0:
                want = -O.decoder_log_likelihood(x0, pred, torch.tensor(0.5 * c.log_pred_var))
            else:
                want = 0.5 * (-1.0 + 1.0 + (post - pred) ** 2 * torch.exp(torch.tensor(-c.log_pred_var)))
            kl = torch.empty(N, per, device="cuda")
            ss = torch.zeros(N, dtype=torch.float64, device="cuda")
            gx0, gxt, gmo = x0.cuda(), xt.cuda(), mo.cuda()
            _lib.check(lib.b200_ddpm_kl(gx0.data_ptr(), gxt.data_ptr(), gmo.data_ptr(), C.byref(c),
                                        kl.data_ptr(), ss.data_ptr(), N, per, None), "b200_ddpm_kl")
            torch.cuda.synchronize()
            err = (kl.cpu() - want).abs() / (1 + want.abs())
            assert err.max().item() < 2e-3, (name, t0, err.max().item())   # log of a difference of two nearby fp32 CDF values: ill-conditioned in the tails on both sides
            assert torch.allclose(ss.cpu(), want.double().sum(1), rtol=1e-4), (name, t0)


# ------------------------------------------------------------------------------------------------ SPADE (§8f rank 2)
def test_spade_golden(cuda_device):
    """SPADEDiffusionModelUNet / SPADEAutoencoderKL against the UNMODIFIED reference's outputs (fixtures from
    tests/golden/make_golden_spade.py): state_dict loads strictly, forward(x, t, seg) and decode(z, seg) match."""
    fx = load("g_spade_unet2d")
    m = nets().SPADEDiffusionModelUNet(**fx["kwargs"]).eval()
    m.load_state_dict(fx["state_dict"])
    check(m.cuda()(fx["x"].cuda(), fx["t"].cuda(), fx["seg"].cuda()), fx["y"], FWD_TOL, "SPADE UNet golden")
    fx = load("g_spade_aekl2d")
    ae = nets().SPADEAutoencoderKL(**fx["kwargs"]).eval()
    ae.load_state_dict(fx["state_dict"])
    ae.cuda()
    mu, sigma = ae.encode(fx["x"].cuda())
    check(mu, fx["mu"], FWD_TOL, "SPADE AE mu")
    check(sigma, fx["sigma"], FWD_TOL, "SPADE AE sigma")
    check(ae.decode(fx["mu"].cuda(), fx["seg"].cuda()), fx["rec"], FWD_TOL, "SPADE AE decode golden")


@pytest.mark.parametrize("name", list(G.SPADE_UNET_CASES))
def test_spade_unet_vs_oracle(cuda_device, name):
    kw, inp = G.SPADE_UNET_CASES[name], G.SPADE_UNET_INPUTS[name]
    torch.manual_seed(0)
    m = G.randomize_zero_params(nets().SPADEDiffusionModelUNet(**kw)).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    torch.manual_seed(1)
    x = torch.randn(*inp["shape"])
    t = torch.randint(0, 1000, (inp["shape"][0],)).long()
    seg = G.seg_onehot(inp["seg"])
    ctx = torch.randn(*inp["context"]) if "context" in inp else None
    want = O.unet_forward(sd, G.unet_oracle_cfg(kw), x, t, context=ctx, seg=seg)
    got = m.cuda()(x.cuda(), t.cuda(), seg.cuda(), context=None if ctx is None else ctx.cuda())
    check(got, want, FWD_TOL, f"SPADE UNet {name}")


@pytest.mark.parametrize("name", list(G.SPADE_AEKL_CASES))
def test_spade_autoencoderkl_vs_oracle(cuda_device, name):
    kw = G.SPADE_AEKL_CASES[name]
    torch.manual_seed(0)
    m = nets().SPADEAutoencoderKL(**kw).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    torch.manual_seed(1)
    x = torch.randn(*G.SPADE_AEKL_INPUTS[name])
    seg = G.seg_onehot((x.shape[0], kw["label_nc"], *x.shape[2:]))
    cfg = G.aekl_oracle_cfg(kw)
    mu_w, sig_w = O.autoencoderkl_encode(sd, cfg, x)
    rec_w = O.autoencoderkl_decode(sd, cfg, mu_w, seg=seg)
    m.cuda()
    mu, sig = m.encode(x.cuda())
    check(mu, mu_w, FWD_TOL, "SPADE AE mu")
    check(m.decode(mu_w.cuda(), seg.cuda()), rec_w, FWD_TOL, f"SPADE AE decode {name}")


def test_spade_latent_sampling_vs_oracle(cuda_device):
    """LatentDiffusionInferer.sample with SPADE networks on both stages and ``seg=`` (inferer.py:431-474): 3 DDIM steps
    of the SPADE UNet in latent space, then SPADEAutoencoderKL.decode(z / scale, seg)."""
    from generativemodels_b200.inferers import LatentDiffusionInferer
    from generativemodels_b200.networks.schedulers import DDIMScheduler
    akw = G.SPADE_AEKL_CASES["spade_aekl2d"]
    ukw = dict(G.SPADE_UNET_CASES["spade_unet2d"], in_channels=akw["latent_channels"],
               out_channels=akw["latent_channels"])
    torch.manual_seed(0)
    ae = nets().SPADEAutoencoderKL(**akw).eval()
    un = G.randomize_zero_params(nets().SPADEDiffusionModelUNet(**ukw)).eval()
    asd = {k: v.clone() for k, v in ae.state_dict().items()}
    usd = {k: v.clone() for k, v in un.state_dict().items()}
    skw = dict(num_train_timesteps=1000, schedule="scaled_linear_beta", beta_start=0.0015, beta_end=0.0205,
               clip_sample=False)
    so, sp = O.DDIMOracle(**skw), DDIMScheduler(**skw)
    so.set_timesteps(3)
    sp.set_timesteps(3)
    torch.manual_seed(4)
    noise = torch.randn(2, akw["latent_channels"], 8, 8)
    seg = G.seg_onehot((2, 3, 32, 32))
    ucfg = G.unet_oracle_cfg(ukw)
    lat = O.diffusion_sample(lambda x, t, c: O.unet_forward(usd, ucfg, x, t, context=c, seg=seg), so, noise)
    want = O.autoencoderkl_decode(asd, G.aekl_oracle_cfg(akw), lat / 0.7, seg=seg)
    got = LatentDiffusionInferer(sp, scale_factor=0.7).sample(noise.cuda(), ae.cuda(), un.cuda(), sp, seg=seg.cuda(),
                                                             verbose=False)
    check(got, want, TRAJ_TOL, "SPADE latent sampling")
    other = nets().SPADEDiffusionModelUNet(**dict(ukw, label_nc=5)).cuda().eval()
    with pytest.raises(ValueError):
        LatentDiffusionInferer(sp, scale_factor=0.7).sample(noise.cuda(), ae, other, sp, seg=seg.cuda(), verbose=False)


# ------------------------------------------------------------------------------------------------ transformer (§8f rank 3)
@pytest.mark.parametrize("cross", [False, True])
def test_transformer_vs_oracle(cuda_device, cross):
    """DecoderOnlyTransformer forward (logits) vs the oracle, and the key/value-cache ``step`` vs the full forward."""
    torch.manual_seed(0)
    m = nets().DecoderOnlyTransformer(num_tokens=70, max_seq_len=40, attn_layers_dim=64, attn_layers_depth=3,
                                      attn_layers_heads=4, with_cross_attention=cross).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    x = torch.randint(0, 70, (2, 33))
    ctx = torch.randn(2, 5, 64) if cross else None
    want = O.transformer_forward(sd, 4, x, ctx)
    m.cuda()
    got = m(x.cuda(), context=None if ctx is None else ctx.cuda())
    check(got, want, FWD_TOL, "transformer logits")
    cache = m.new_cache(2, torch.device("cuda"), None if ctx is None else ctx.cuda())
    inc = torch.cat([m.step(x[:, :7].cuda(), cache)] + [m.step(x[:, i:i + 1].cuda(), cache) for i in range(7, 33)], 1)
    check(inc, got.float().cpu(), 5e-3, "incremental decoding vs full forward")
    # single-token steps replayed from ONE captured CUDA graph (prefix length in device memory), after a 7-token prompt
    gc = m.new_cache(2, torch.device("cuda"), None if ctx is None else ctx.cuda(), graph=True)
    rows = [m.step(x[:, :7].cuda(), gc)] + [m.step(x[:, i:i + 1].cuda(), gc).clone() for i in range(7, 33)]
    assert gc.graph is not None and int(gc.pos_dev) == 33
    check(torch.cat(rows, 1), got.float().cpu(), 1e-2, "graph-replayed decoding vs full forward")
    check(torch.cat(rows, 1), want, FWD_TOL, "graph-replayed decoding vs oracle")


def test_vqvae_transformer_inferer_vs_oracle(cuda_device):
    """VQVAETransformerInferer.sample (top_k = 1 -> deterministic; cache, then sliding window), __call__ and
    get_likelihood (inferer.py:1126-1330) with the CUDA VQVAE + transformer against the oracle."""
    from generativemodels_b200.inferers import VQVAETransformerInferer
    from generativemodels_b200.utils.ordering import Ordering
    kw = G.VQVAE_CASES["vqvae2d"]
    K = kw["num_embeddings"]
    torch.manual_seed(0)
    vq = nets().VQVAE(**kw).eval()
    vsd = {k: v.clone() for k, v in vq.state_dict().items()}
    tr = nets().DecoderOnlyTransformer(num_tokens=K + 1, max_seq_len=12, attn_layers_dim=64, attn_layers_depth=2,
                                       attn_layers_heads=4).eval()
    tsd = {k: v.clone() for k, v in tr.state_dict().items()}
    ordering = Ordering("s_curve", 2, (1, 4, 4))
    vq.cuda(), tr.cuda()
    inf = VQVAETransformerInferer()
    # likelihood of an image: token log-probabilities vs the oracle's logits on the same (exact) indices
    torch.manual_seed(3)
    x = torch.randn(2, 1, 16, 16)
    ll = inf.get_likelihood(x.cuda(), vq, tr, ordering).cpu()
    z = O.vqvae_encode(vsd, G.vqvae_oracle_cfg(kw), x)
    idx = O.vq_quantize(vsd["quantizer.quantizer.embedding.weight"], z)[1] if False else vq.index_quantize(x.cuda()).cpu()
    lat = idx.reshape(2, -1)[:, ordering.get_sequence_ordering()]
    seq = torch.nn.functional.pad(lat, (1, 0), "constant", K).long()
    lp = torch.log_softmax(O.transformer_forward(tsd, 4, seq[:, :12]), -1)
    want = torch.gather(lp, 2, seq[:, 1:13].unsqueeze(2)).squeeze(2)          # first max_seq_len targets
    got = ll.reshape(2, -1)[:, ordering.get_sequence_ordering()][:, :12]
    assert (got - want).abs().max().item() < 5e-2, (got - want).abs().max().item()
    pred = inf(x.cuda(), vq, tr, ordering)
    assert pred.shape == (2, 12, K + 1)
    # greedy sampling: a flipped argmax (bf16 logits) changes everything after it, so compare the token sequences and
    # only require the decoded images to match when they agree
    start = torch.full((2, 1), K).cuda()
    img = inf.sample((4, 4), start, vq, tr, ordering, top_k=1, verbose=False)
    assert img.shape == (2, 1, 16, 16) and torch.isfinite(img).all()
    seq_w = O.transformer_sample_greedy(tsd, 4, 12, K, 16, 2)
    lat_w = seq_w[:, ordering.get_revert_sequence_ordering()].reshape(2, 4, 4)
    want_img = O.vqvae_decode(vsd, G.vqvae_oracle_cfg(kw), O.vq_embed(vsd["quantizer.quantizer.embedding.weight"], lat_w))
    agree = rel(img, want_img)
    if agree > FWD_TOL:      # near-tie somewhere: at least the first tokens must coincide
        first = tr(torch.full((2, 1), K).cuda())[:, -1, :K].argmax(-1).cpu()
        assert torch.equal(first, seq_w[:, 0])
    else:
        assert agree <= FWD_TOL


def test_no_cpu_path(cuda_device):
    m = nets().DiffusionModelUNet(2, 1, 1, num_res_blocks=1, num_channels=(8, 8), attention_levels=(False, False),
                                  norm_num_groups=4).cuda()
    with pytest.raises(RuntimeError):
        m(torch.randn(1, 1, 8, 8), torch.tensor([1]))


def test_cuda_graph_replay_matches_eager(cuda_device):
    """generativemodels_b200.cuda_graph.graphed(): captured replay is bit-identical to the eager forward and can be
    driven by DiffusionInferer.sample unchanged."""
    from generativemodels_b200.cuda_graph import graphed
    from generativemodels_b200.inferers import DiffusionInferer
    from generativemodels_b200.networks.schedulers import DDIMScheduler
    kw = G.UNET_CASES["unet2d_attn"]
    torch.manual_seed(0)
    m = G.randomize_zero_params(nets().DiffusionModelUNet(**kw)).cuda().eval()
    g = graphed(m)
    torch.manual_seed(1)
    x = torch.randn(2, 1, 16, 16).cuda()
    for t in (900.0, 20.0):
        ts = torch.Tensor((t,)).cuda()
        assert torch.equal(g(x, timesteps=ts).clone(), m(x, timesteps=ts))
    s = DDIMScheduler(num_train_timesteps=1000)
    s.set_timesteps(4)
    a = DiffusionInferer(s).sample(x, m, s, verbose=False)
    b = DiffusionInferer(s).sample(x, g, s, verbose=False)
    assert torch.equal(a, b)
    # new weights invalidate the captured graphs (they bake in the packed weights' addresses)
    with torch.no_grad():
        m.conv_in.conv.weight.mul_(1.5)
    ts = torch.Tensor((500.0,)).cuda()
    assert torch.equal(g(x, timesteps=ts).clone(), m(x, timesteps=ts))


def test_cuda_graph_with_pndm_history(cuda_device):
    """A graph-replayed network hands out a fresh tensor per call: PNDMScheduler keeps up to four past model outputs
    (pndm.py:230-291), which a shared static output buffer would silently alias (all PLMS weights on the current eps).
    Graphed and eager sampling must agree bit for bit, with and without the Runge-Kutta warm-up."""
    from generativemodels_b200.cuda_graph import graphed
    from generativemodels_b200.inferers import DiffusionInferer
    from generativemodels_b200.networks.schedulers import PNDMScheduler
    kw = G.UNET_CASES["unet2d_attn"]
    torch.manual_seed(0)
    m = G.randomize_zero_params(nets().DiffusionModelUNet(**kw)).cuda().eval()
    g = graphed(m)
    torch.manual_seed(1)
    x = torch.randn(2, 1, 16, 16).cuda()
    for skip in (True, False):
        outs = []
        for net in (m, g):
            s = PNDMScheduler(num_train_timesteps=1000, skip_prk_steps=skip)
            s.set_timesteps(8)
            outs.append(DiffusionInferer(s).sample(x, net, s, verbose=False))
        assert torch.equal(outs[0], outs[1]), f"graphed PNDM sampling differs from eager (skip_prk_steps={skip})"
    ts = torch.Tensor((500.0,)).cuda()
    a, b = g(x, timesteps=ts), g(x + 1, timesteps=ts)
    assert a.data_ptr() != b.data_ptr() and not torch.equal(a, b)


def test_param_data_surgery_needs_invalidate(cuda_device):
    """Writes through ``param.data`` (EMA swaps) do not bump the version counter the packed-weight caches key on;
    ``generativemodels_b200.invalidate_packed`` is the documented way to make the next forward see them."""
    import generativemodels_b200 as B
    kw = G.UNET_CASES["unet2d_attn"]
    torch.manual_seed(0)
    m = G.randomize_zero_params(nets().DiffusionModelUNet(**kw)).cuda().eval()
    torch.manual_seed(1)
    x, ts = torch.randn(2, 1, 16, 16).cuda(), torch.Tensor((300.0,)).cuda()
    y0 = m(x, timesteps=ts)
    new = {k: v.clone() for k, v in m.state_dict().items()}
    for k in new:
        if k.endswith("conv1.conv.weight"):
            new[k] = new[k] * 1.25
    with torch.no_grad():
        for k, p_ in m.named_parameters():
            p_.data.copy_(new[k])
    B.invalidate_packed(m)
    y1 = m(x, timesteps=ts)
    want = O.unet_forward({k: v.cpu() for k, v in new.items()}, G.unet_oracle_cfg(kw), x.cpu(), ts.cpu())
    check(y1, want)
    assert not torch.equal(y0, y1)


def test_cuda_graph_nested_inputs_outputs(cuda_device):
    """graphed() with a tuple of tensors as OUTPUT (ControlNet residuals) and a list of tensors + a scalar as INPUTS
    (UNet with down_block_additional_residuals, ControlNet conditioning_scale): replay equals the eager call, outputs
    are fresh tensors per call, a different scalar is a different graph."""
    from generativemodels_b200.cuda_graph import graphed
    torch.manual_seed(0)
    cn = G.randomize_zero_params(nets().ControlNet(**G.CONTROLNET_CASE)).cuda().eval()
    ukw = {k: v for k, v in G.CONTROLNET_CASE.items() if not k.startswith("conditioning_embedding")}
    un = G.randomize_zero_params(nets().DiffusionModelUNet(out_channels=3, **ukw)).cuda().eval()
    gcn, gun = graphed(cn), graphed(un)
    torch.manual_seed(1)
    x, cond, ctx = torch.randn(2, 3, 16, 16).cuda(), torch.rand(2, 1, 16, 16).cuda(), torch.randn(2, 3, 8).cuda()
    for t, scale in ((700.0, 1.0), (100.0, 1.0), (100.0, 0.5)):
        ts = torch.Tensor((t,)).cuda()
        down, mid = cn(x, ts, cond, conditioning_scale=scale, context=ctx)
        gdown, gmid = gcn(x, ts, cond, conditioning_scale=scale, context=ctx)
        assert len(gdown) == len(down) and all(torch.equal(a, b) for a, b in zip(gdown, down)) and torch.equal(gmid, mid)
        want = un(x, ts, context=ctx, down_block_additional_residuals=down, mid_block_additional_residual=mid)
        got = gun(x, ts, context=ctx, down_block_additional_residuals=gdown, mid_block_additional_residual=gmid)
        assert torch.equal(got, want)
    assert len(gcn._entries) == 2            # two conditioning scales -> two captured graphs
    a = gcn(x, ts, cond, conditioning_scale=0.5, context=ctx)[0][0]
    b = gcn(x + 1, ts, cond, conditioning_scale=0.5, context=ctx)[0][0]
    assert a.data_ptr() != b.data_ptr() and not torch.equal(a, b)


@pytest.mark.parametrize("act", ["tanh", "sigmoid"])
def test_vqvae_output_act(cuda_device, act):
    """VQVAE(output_act=...) (vqvae.py:263-264) against the oracle: the activation is applied in the epilogue of the last
    transposed convolution's phase launches."""
    kw = dict(G.VQVAE_CASES["vqvae3d"], output_act=act)
    torch.manual_seed(0)
    m = nets().VQVAE(**kw).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    torch.manual_seed(4)
    z = torch.randn(1, kw["embedding_dim"], 4, 4, 4)
    want = O.vqvae_decode(sd, G.vqvae_oracle_cfg(kw), z)
    check(m.cuda().decode(z.cuda()), want, FWD_TOL, f"vqvae decode with output_act={act}")
