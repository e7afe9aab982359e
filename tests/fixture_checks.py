"""Checks of the product against the reference-generated fixtures of SURVEY.md §8f ranks 1 and 3
(tests/golden/g_likelihood.pt, g_transformer.pt; tests/golden/make_golden_next.py).  One body per check, run on the CPU
stand-in by tests/test_modules_cpu.py and on the CUDA path by tests/test_parity_gpu.py."""
from pathlib import Path

import torch

GOLD = Path(__file__).resolve().parent / "golden"


# The tolerances of the whole suite, stated once.  16-bit operands (fp16 by default), fp32
# accumulation, against the fp32 reference: relative L2 of a network forward <= TOL_REL, largest pointwise deviation
# <= TOL_MAX of the largest reference magnitude (a mis-strided plane or phase is O(1) there while staying small in
# L2), integer results (code indices, timestep tables) exact, fp32 scheduler arithmetic <= 1e-5.
TOL_REL = 2e-2
TOL_MAX = 4e-2
TOL_TRAJ = 5e-2          # sampler trajectories of several steps (rounding is amplified by a random-weight network)


def rel(a, b):
    return ((a.float().cpu() - b.float().cpu()).norm() / (b.float().cpu().norm() + 1e-12)).item()


def relmax(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def close(a, b, what, tol_rel=TOL_REL, tol_max=TOL_MAX):
    """Shape, relative-L2 and normalised max-abs check of one tensor against the reference's; returns both errors."""
    assert tuple(a.shape) == tuple(b.shape), (what, tuple(a.shape), tuple(b.shape))
    assert bool(torch.isfinite(a.float()).all()), (what, "non-finite values")
    r, m = rel(a, b), relmax(a, b)
    assert r < tol_rel and m < tol_max, (what, f"rel-L2 {r:.3e} (tol {tol_rel:g}), max-abs {m:.3e} (tol {tol_max:g})")
    return r, m


def check_likelihood_fixture(device, monkeypatch):
    from generativemodels_b200.inferers import DiffusionInferer
    from generativemodels_b200.networks.nets import DiffusionModelUNet
    from generativemodels_b200.networks.schedulers import DDPMScheduler
    fx = torch.load(GOLD / "g_likelihood.pt", weights_only=False)
    net_fx = torch.load(GOLD / fx["unet_fixture"], weights_only=False)
    m = DiffusionModelUNet(**net_fx["kwargs"]).eval()
    m.load_state_dict(net_fx["state_dict"])
    m = m.to(device)
    x, noise = fx["x"].to(device), fx["noise"].to(device)
    monkeypatch.setattr(torch, "randn_like", lambda t: noise.clone())
    for ptype, case in fx["cases"].items():
        s = DDPMScheduler(**case["scheduler_kwargs"])
        s.set_timesteps(case["scheduler_kwargs"]["num_train_timesteps"])
        lik, inter = DiffusionInferer(s).get_likelihood(inputs=x, diffusion_model=m, scheduler=s,
                                                        save_intermediates=True, verbose=False)
        want = case["likelihood"]
        assert lik.shape == want.shape and len(inter) == len(case["intermediates"]), ptype
        # bf16 network, fp32 KL: the per-sample bound agrees to a few percent, every per-step map in relative L2
        assert torch.allclose(lik.cpu(), want, rtol=5e-2, atol=1e-3), (ptype, lik.cpu(), want)
        worst = max(rel(a, b) for a, b in zip(inter, case["intermediates"]))
        assert worst < 8e-2, (ptype, worst)


def check_transformer_fixture(device):
    from generativemodels_b200.inferers import VQVAETransformerInferer
    from generativemodels_b200.networks.nets import VQVAE, DecoderOnlyTransformer
    from generativemodels_b200.utils.ordering import Ordering
    fx = torch.load(GOLD / "g_transformer.pt", weights_only=False)
    for name in ("plain", "cross"):
        c = fx[name]
        tr = DecoderOnlyTransformer(**c["kwargs"]).eval()
        assert set(tr.state_dict()) == set(c["state_dict"])
        tr.load_state_dict(c["state_dict"])
        tr = tr.to(device)
        ctx = None if c["context"] is None else c["context"].to(device)
        got = tr(c["tokens"].to(device), context=ctx)
        assert got.shape == c["logits"].shape and rel(got, c["logits"]) < 2e-2, (name, rel(got, c["logits"]))
    s = fx["sampler"]
    vq = VQVAE(**s["vqvae_kwargs"]).eval()
    vq.load_state_dict(s["vqvae_state"])
    tr = DecoderOnlyTransformer(**s["transformer_kwargs"]).eval()
    tr.load_state_dict(s["transformer_state"])
    vq, tr = vq.to(device), tr.to(device)
    ordering = Ordering(**s["ordering_kwargs"])
    got = VQVAETransformerInferer().sample((4, 4), s["start"].to(device), vq, tr, ordering, top_k=1, verbose=False)
    # greedy decoding is discrete: a token flipped by bf16 logits would change whole codebook vectors, so the decoded
    # image either matches to bf16 round-off or is grossly off
    assert got.shape == s["sample"].shape and rel(got, s["sample"]) < 3e-2, rel(got, s["sample"])


def check_c1_fixture(device, sample_tol=5e-2):
    """BASELINE.json configs[0]: the tutorial's 2-D UNet, DDPM with 4 inference steps, batch 2 of 1x64x64, against the
    unmodified reference's run (tests/golden/make_golden_c1.py); weights from the shared recipe."""
    from generativemodels_b200.inferers import DiffusionInferer
    from generativemodels_b200.networks.nets import DiffusionModelUNet
    from generativemodels_b200.networks.schedulers import DDPMScheduler
    from tests.golden import configs as G
    fx = torch.load(GOLD / "g_c1.pt", weights_only=False)
    m = DiffusionModelUNet(**G.C1_UNET).eval()
    G.recipe_state_dict(m)
    assert sum(p.numel() for p in m.parameters()) == fx["n_params"]
    m = m.to(device)
    s = DDPMScheduler(num_train_timesteps=1000)
    s.set_timesteps(G.C1_STEPS)
    assert [int(t) for t in s.timesteps] == [750, 500, 250, 0]
    noise = fx["noise"].to(device)
    y = m(noise, torch.tensor([500, 500], device=device))
    # every weight of this recipe is O(1/sqrt(fan_in)) — also the convolutions a trained / freshly initialised network
    # keeps near zero — so bf16 rounding of the activations shows more than in the other fixtures: 1.0e-2 relative L2
    # on the bf16 stand-in for the forward (tolerance 3e-2), 1e-4 for the 4-step DDPM sample
    assert y.shape == fx["y500"].shape and rel(y, fx["y500"]) < 3e-2, rel(y, fx["y500"])
    torch.manual_seed(fx["ddpm_seed"])
    sample = DiffusionInferer(s).sample(input_noise=noise, diffusion_model=m, scheduler=s, verbose=False)
    assert sample.shape == fx["sample"].shape and rel(sample, fx["sample"]) < sample_tol, rel(sample, fx["sample"])


# Probe 24 (t = 500) of the random-weight C2 trajectory is ill-conditioned (a peaked softmax amplifies every upstream
# rounding): an all-bf16 data path is 7.9e-2 off the fp32 reference there, an all-fp16 one 7.8e-3 (1.2e-2 / 1.5e-3 at
# the other probes) — measured by rounding every weight and activation inside the fp32 oracle.  Making only the
# attention arithmetic fp32 leaves 7.2e-2, so this is what decided the library's 16-bit format.
# All four probes are held to the suite's tolerance on every backend; there is no exemption.
def check_c2_fixture(device, probes=(0, 1, 24, 49)):
    """BASELINE.json configs[1]: LDM-tutorial AutoencoderKL + latent UNet, DDIM-50, pinned teacher-forced along the
    unmodified reference's trajectory (tests/golden/make_golden_c2.py): network output and scheduler step at the probe
    steps from the reference's x_k, and the decoder on the reference's final latent."""
    from generativemodels_b200.networks.nets import AutoencoderKL, DiffusionModelUNet
    from generativemodels_b200.networks.schedulers import DDIMScheduler
    from tests.golden import configs as G
    fx = torch.load(GOLD / "g_c2.pt", weights_only=False)
    unet = DiffusionModelUNet(**G.C2_UNET).eval()
    G.recipe_state_dict(unet, 12)
    unet = unet.to(device)
    s = DDIMScheduler(**G.C2_SCHEDULER)
    s.set_timesteps(50)
    report = {}
    for k in probes:
        p = fx["probes"][k]
        assert int(s.timesteps[k]) == p["t"]
        x = p["x"].to(device)
        eps = unet(x, timesteps=torch.Tensor((p["t"],)).to(device))
        nxt, _ = s.step(eps, p["t"], x)
        exact, _ = s.step(p["eps"].to(device), p["t"], x)            # the scheduler alone: fp32 on both sides
        assert rel(exact, p["nxt"]) < 1e-5, (k, rel(exact, p["nxt"]))
        report[k] = (close(eps, p["eps"], f"C2 probe {k} eps"), close(nxt, p["nxt"], f"C2 probe {k} next latent"))
    ae = AutoencoderKL(**G.C2_AEKL).eval()
    G.recipe_state_dict(ae, 11)
    ae = ae.to(device)
    img = ae.decode_stage_2_outputs(fx["latent"].to(device))
    report["decode"] = close(img, fx["image"], "C2 decoder")
    return report


def check_c3_fixture(device, steps=True):
    """BASELINE.json configs[2] — the bench's headline model: 3-D UNet (256, 256, 512), attention (F, F, T), head 512,
    on the tutorial volume 32x40x32 (T = 640 tokens of 512 channels at the attention level, 512 -> 256 virtual-concat
    convolutions, stride-2 and folded-upsample convolutions): one forward and a DDIM-5 sample against the unmodified
    reference's CPU run (tests/golden/make_golden_c345.py)."""
    from generativemodels_b200.inferers import DiffusionInferer
    from generativemodels_b200.networks.nets import DiffusionModelUNet
    from generativemodels_b200.networks.schedulers import DDIMScheduler
    from tests.golden import configs as G
    fx = torch.load(GOLD / "g_c3.pt", weights_only=False)
    m = DiffusionModelUNet(**G.C3_UNET).eval()
    G.recipe_state_dict(m, 13)
    assert sum(p.numel() for p in m.parameters()) == fx["n_params"]
    m = m.to(device)
    noise = fx["noise"].to(device)
    report = {"forward": close(m(noise, torch.Tensor((500,)).to(device)), fx["y500"], "C3 forward t=500")}
    if steps:
        s = DDIMScheduler(**G.C3_SCHEDULER)
        s.set_timesteps(G.C3_STEPS)
        assert [int(t) for t in s.timesteps] == fx["timesteps"]
        sample, inter = DiffusionInferer(s).sample(input_noise=noise, diffusion_model=m, scheduler=s, verbose=False,
                                                   save_intermediates=True, intermediate_steps=1)
        assert len(inter) == len(fx["intermediates"])
        for k, (a, b) in enumerate(zip(inter, fx["intermediates"])):
            report[f"x_{k + 1}"] = close(a, b, f"C3 DDIM-5 after step {k + 1}", TOL_TRAJ, 2 * TOL_TRAJ)
        report["sample"] = close(sample, fx["sample"], "C3 DDIM-5 sample", TOL_TRAJ, 2 * TOL_TRAJ)
    return report


def check_c4_fixture(device):
    """BASELINE.json configs[3]: 3-D VQVAE (256, 256), 256 codes x 32, encode -> quantise -> decode on 1x64^3 against
    the unmodified reference.  Code indices: (1) the quantiser on the reference's own encoder output must reproduce
    the reference's indices bit for bit; (2) through this encoder (16-bit convolutions) an index may only differ where
    the reference's own best / second-best distance gap is below what the encoder's rounding can move, and those are
    counted."""
    from generativemodels_b200.networks.nets import VQVAE
    from tests.golden import configs as G
    fx = torch.load(GOLD / "g_c4.pt", weights_only=False)
    m = VQVAE(**G.C4_VQVAE).eval()
    G.recipe_state_dict(m, 14)
    assert sum(p.numel() for p in m.parameters()) == fx["n_params"]
    m = m.to(device)
    x = fx["x"].to(device)
    report = {}
    # (1) quantiser alone, on the reference's z: exact
    q_in = fx["z"].to(device)
    idx_ref_z = m.quantizer.quantize(q_in)
    n_bad = int((idx_ref_z.cpu() != fx["indices"]).sum())
    assert n_bad == 0, f"C4: {n_bad} code indices differ on the reference's own encoder output"
    # (2) whole pipeline
    z = m.encode(x)
    report["z"] = close(z, fx["z"], "C4 encoder output")
    idx = m.index_quantize(x).cpu()
    assert idx.shape == fx["indices"].shape and idx.dtype == torch.int64
    flipped = idx != fx["indices"]
    dz = (z.cpu() - fx["z"]).norm(dim=1)                       # per-vector encoder deviation
    # a flip needs |z - z_ref| to bridge the gap: d2 - d1 <= 2 |dz| (|e1 - e2|) <= 2 |dz| * diam(codebook)
    cb = m.quantizer.quantizer.embedding.weight.detach().float().cpu()
    diam = torch.cdist(cb, cb).max()
    allowed = fx["margin"] <= 2.0 * dz * diam + 1e-6
    assert bool((~flipped | allowed).all()), "C4: a code index differs where the reference's margin is not a near-tie"
    report["index_flips"] = (int(flipped.sum()), int(flipped.numel()))
    assert report["index_flips"][0] <= 0.01 * flipped.numel(), report["index_flips"]
    # decoder on the reference's indices (no dependence on near-ties), and the full reconstruction
    report["decode"] = close(m.decode_samples(fx["indices"].to(device)), fx["recon_from_idx"], "C4 decoder")
    recon, loss = m(x)
    report["recon"] = (rel(recon, fx["recon"]), relmax(recon, fx["recon"]))
    assert tuple(recon.shape) == tuple(fx["recon"].shape)
    if report["index_flips"][0] == 0:
        close(recon, fx["recon"], "C4 reconstruction")
    assert abs(float(loss) - float(fx["loss"])) <= 2e-2 * abs(float(fx["loss"])) + 1e-6, (float(loss), float(fx["loss"]))
    return report


def check_c5_fixture(device):
    """BASELINE.json configs[4]: ControlNet + conditioned UNet (128, 256, 256) at 3x256x256, one classifier-free-guidance
    DDIM step as the tutorials run it (batch doubled, context [-1, +1], guidance 7) against the unmodified reference:
    T = 16 384 self-attention (head 256) at 128^2, S = 1 cross-attention, zero-conv residual adds."""
    from generativemodels_b200.networks.nets import ControlNet, DiffusionModelUNet
    from generativemodels_b200.networks.schedulers import DDIMScheduler
    from tests.golden import configs as G
    fx = torch.load(GOLD / "g_c5.pt", weights_only=False)
    unet = DiffusionModelUNet(**G.C5_UNET).eval()
    cn = ControlNet(**G.C5_CONTROLNET).eval()
    G.recipe_state_dict(unet, 15)
    G.recipe_state_dict(cn, 16)
    assert (sum(p.numel() for p in unet.parameters()), sum(p.numel() for p in cn.parameters())) == tuple(fx["n_params"])
    unet, cn = unet.to(device), cn.to(device)
    s = DDIMScheduler(num_train_timesteps=1000)
    s.set_timesteps(50)
    t = int(s.timesteps[G.C5_T_INDEX])
    assert t == fx["t"]
    x = fx["x"].to(device)
    ctx = torch.cat([-1 * torch.ones(1, 1, 1), torch.ones(1, 1, 1)], dim=0).to(device)
    x2 = torch.cat([x] * 2)
    ts = torch.Tensor((t,)).to(device)
    down, mid = cn(x=x2, timesteps=ts, controlnet_cond=torch.cat([G.c5_mask()] * 2).to(device), context=ctx)
    report = {"cn_mid_mean": close(mid.mean((2, 3)), fx["mid_mean"], "C5 ControlNet mid residual (channel means)")}
    for i, (d, w) in enumerate(zip(down, fx["down_means"])):
        close(d.mean((2, 3)), w, f"C5 ControlNet down residual {i} (channel means)")
    eps2 = unet(x2, timesteps=ts, context=ctx, down_block_additional_residuals=down, mid_block_additional_residual=mid)
    report["eps2"] = close(eps2, fx["eps2"], "C5 UNet output (uncond | cond)")
    eu, et = eps2.chunk(2)
    eps = eu + G.C5_GUIDANCE * (et - eu)
    # guidance 7 amplifies the (cond - uncond) difference of two nearly equal outputs: the guided eps gets the
    # trajectory tolerance
    report["eps"] = close(eps, fx["eps"], "C5 guided eps", TOL_TRAJ, 2 * TOL_TRAJ)
    nxt, _ = s.step(eps, t, x)
    report["nxt"] = close(nxt, fx["nxt"], "C5 next sample")
    return report
