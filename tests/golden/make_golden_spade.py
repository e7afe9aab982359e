"""Generate the SPADE fixtures (tests/golden/g_spade_*.pt) by running the UNMODIFIED reference
(/root/reference, CPU fp32, MONAI shim):   python -m tests.golden.make_golden_spade

SPADEDiffusionModelUNet (spade_diffusion_model_unet.py:612-912) and SPADEAutoencoderKL (spade_autoencoderkl.py:292-484)
— the §8f rank-2 widening of the sampling path; same fixture layout as make_golden.py.
"""
from pathlib import Path

import torch

from tests.golden import configs as G      # before the reference import: /root/reference has its own `tests` package
from oracle import ref_import

OUT = Path(__file__).resolve().parent

SPADE_UNET = dict(spatial_dims=2, in_channels=1, out_channels=1, label_nc=3, num_res_blocks=1, num_channels=(8, 16),
                  attention_levels=(False, True), norm_num_groups=4, num_head_channels=8,
                  spade_intermediate_channels=16)
SPADE_AEKL = dict(spatial_dims=2, label_nc=3, in_channels=1, out_channels=1, num_channels=(8, 16), latent_channels=3,
                  num_res_blocks=(1, 1), norm_num_groups=4, attention_levels=(False, True),
                  with_encoder_nonlocal_attn=True, with_decoder_nonlocal_attn=True, spade_intermediate_channels=8)


def main():
    ref_import.import_reference()
    from generative.networks.nets import SPADEAutoencoderKL, SPADEDiffusionModelUNet
    torch.manual_seed(0)
    m = G.randomize_zero_params(SPADEDiffusionModelUNet(**SPADE_UNET)).eval()
    torch.manual_seed(1)
    x, t = torch.randn(2, 1, 8, 8), torch.tensor([3, 800])
    seg = torch.nn.functional.one_hot(torch.randint(0, 3, (2, 16, 16)), 3).permute(0, 3, 1, 2).float()
    with torch.no_grad():
        y = m(x, t, seg)
    torch.save(dict(kwargs=SPADE_UNET, state_dict=m.state_dict(), x=x, t=t, seg=seg, y=y), OUT / "g_spade_unet2d.pt")

    torch.manual_seed(2)
    ae = SPADEAutoencoderKL(**SPADE_AEKL).eval()
    x = torch.randn(2, 1, 16, 16)
    seg = torch.nn.functional.one_hot(torch.randint(0, 3, (2, 16, 16)), 3).permute(0, 3, 1, 2).float()
    with torch.no_grad():
        mu, sigma = ae.encode(x)
        rec = ae.decode(mu, seg)
    torch.save(dict(kwargs=SPADE_AEKL, state_dict=ae.state_dict(), x=x, seg=seg, mu=mu, sigma=sigma, rec=rec),
               OUT / "g_spade_aekl2d.pt")
    for f in sorted(OUT.glob("g_spade*.pt")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
