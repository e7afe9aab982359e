"""Checks of the product against the reference-generated fixtures of SURVEY.md §8f ranks 1 and 3
(tests/golden/g_likelihood.pt, g_transformer.pt; tests/golden/make_golden_next.py).  One body per check, run on the CPU
stand-in by tests/test_modules_cpu.py and on the CUDA path by tests/test_parity_gpu.py."""
from pathlib import Path

import torch

GOLD = Path(__file__).resolve().parent / "golden"


def rel(a, b):
    return ((a.float().cpu() - b.float().cpu()).norm() / (b.float().cpu().norm() + 1e-12)).item()


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


# Probe 24 (t = 500) of the random-weight C2 trajectory sits at a peaked softmax: rounding ONLY q and k to bf16 inside
# the fp32 oracle already moves the network output by 1.2e-2 there (1.5e-4 at the other probes), and the full bf16
# data path by 7e-2 on the stand-in.  It is kept as a loosely bounded, documented case, not hidden.
C2_PROBE_TOL = {0: 3e-2, 1: 3e-2, 24: 2e-1, 49: 3e-2}


def check_c2_fixture(device, probes=(0, 1, 24, 49), strict_ill_conditioned=True):
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
        report[k] = (rel(eps, p["eps"]), rel(nxt, p["nxt"]))
        if k != 24 or strict_ill_conditioned:
            assert max(report[k]) < C2_PROBE_TOL[k], (k, report[k])
    ae = AutoencoderKL(**G.C2_AEKL).eval()
    G.recipe_state_dict(ae, 11)
    ae = ae.to(device)
    img = ae.decode_stage_2_outputs(fx["latent"].to(device))
    assert img.shape == fx["image"].shape and rel(img, fx["image"]) < 3e-2, rel(img, fx["image"])
    return report
