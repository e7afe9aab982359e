"""BASELINE.json configs[1] (C2) through the UNMODIFIED reference on CPU fp32:  python -m tests.golden.make_golden_c2

AutoencoderKL (128, 128, 256) + latent UNet (128, 256, 512) of the 2-D LDM tutorial, DDIMScheduler(linear_beta
0.0015..0.0195) with 50 steps, LatentDiffusionInferer(scale_factor=1).sample on one 3x64x64 latent -> 1x256x256.
Weights from tests.golden.configs.recipe_state_dict (seeds 11 / 12, not committed).

A network with random weights is not a contraction: over 50 steps the bf16 rounding of any implementation is amplified
until the end point is unrelated (measured: relative error 1.5 after 50 steps from 1.7e-2 after two), so the fixture
pins the trajectory *teacher-forced*: at probe steps k it stores the reference's x_k, the network output there and
x_{k+1}; plus the final latent and the image decoded from it."""
from pathlib import Path

import torch

from tests.golden import configs as G      # before the reference import: /root/reference has its own `tests` package
from oracle import ref_import

OUT = Path(__file__).resolve().parent


def main():
    ref_import.import_reference()
    from generative.inferers import DiffusionInferer, LatentDiffusionInferer
    from generative.networks.nets import AutoencoderKL, DiffusionModelUNet
    from generative.networks.schedulers import DDIMScheduler
    ae = AutoencoderKL(**G.C2_AEKL).eval()
    unet = DiffusionModelUNet(**G.C2_UNET).eval()
    G.recipe_state_dict(ae, 11)
    G.recipe_state_dict(unet, 12)
    s = DDIMScheduler(**G.C2_SCHEDULER)
    s.set_timesteps(50)
    torch.manual_seed(4321)
    noise = torch.randn(G.C2_LATENT)
    probes = {}
    with torch.no_grad():
        x = noise
        for k, t in enumerate(s.timesteps):
            eps = unet(x, timesteps=torch.Tensor((t,)))
            nxt, _ = s.step(eps, t, x)
            if k in G.C2_PROBES:
                probes[k] = dict(t=int(t), x=x.clone(), eps=eps.clone(), nxt=nxt.clone())
            x = nxt
        latent = DiffusionInferer(s).sample(input_noise=noise, diffusion_model=unet, scheduler=s, verbose=False)
        assert torch.equal(latent, x)
        image = LatentDiffusionInferer(s, scale_factor=1.0).sample(input_noise=noise, autoencoder_model=ae,
                                                                  diffusion_model=unet, scheduler=s, verbose=False)
    torch.save(dict(noise=noise, probes=probes, latent=latent, image=image), OUT / "g_c2.pt")
    print("g_c2.pt", (OUT / "g_c2.pt").stat().st_size, float(latent.abs().mean()), float(image.abs().mean()))


if __name__ == "__main__":
    main()
