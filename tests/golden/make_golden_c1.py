"""BASELINE.json configs[0] (C1) through the UNMODIFIED reference on CPU fp32:  python -m tests.golden.make_golden_c1

2-D DiffusionModelUNet (128, 256, 256), attention (F, T, T), one res block, head 256 (2d_ddpm_tutorial.py:166-174),
DDPMScheduler(1000).set_timesteps(4) -> t = [750, 500, 250, 0], DiffusionInferer.sample on a batch of 2 x 1x64x64.
Weights come from tests.golden.configs.recipe_state_dict (not committed); the fixture holds the input noise, the seed
of the DDPM noise draws, one forward at t = 500 and the final sample."""
from pathlib import Path

import torch

from tests.golden import configs as G      # before the reference import: /root/reference has its own `tests` package
from oracle import ref_import

OUT = Path(__file__).resolve().parent


def main():
    ref_import.import_reference()
    from generative.inferers import DiffusionInferer
    from generative.networks.nets import DiffusionModelUNet
    from generative.networks.schedulers import DDPMScheduler
    m = DiffusionModelUNet(**G.C1_UNET).eval()
    G.recipe_state_dict(m)
    s = DDPMScheduler(num_train_timesteps=1000)
    s.set_timesteps(G.C1_STEPS)
    assert [int(t) for t in s.timesteps] == [750, 500, 250, 0]
    torch.manual_seed(1234)
    noise = torch.randn(G.C1_SHAPE)
    with torch.no_grad():
        y500 = m(noise, torch.tensor([500, 500]))
        torch.manual_seed(77)
        sample = DiffusionInferer(s).sample(input_noise=noise, diffusion_model=m, scheduler=s, verbose=False)
    torch.save(dict(noise=noise, ddpm_seed=77, y500=y500, sample=sample, n_params=sum(p.numel() for p in m.parameters())),
               OUT / "g_c1.pt")
    print("g_c1.pt", (OUT / "g_c1.pt").stat().st_size, float(y500.abs().mean()), float(sample.abs().mean()))


if __name__ == "__main__":
    main()
