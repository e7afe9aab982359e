"""Fixtures for SURVEY.md §8f ranks 1 and 3 from the UNMODIFIED reference (CPU fp32, MONAI shim):
    python -m tests.golden.make_golden_next

* g_likelihood.pt   — DiffusionInferer.get_likelihood (inferer.py:145-277) on the g_unet2d.pt network (weights are not
                      duplicated): the noise the reference drew (captured from torch.randn_like), the per-sample
                      likelihood and the per-step KL maps, for two prediction types.
* g_transformer.pt  — DecoderOnlyTransformer.forward logits (nets/transformer.py:96-106) with and without
                      cross-attention, and VQVAETransformerInferer.sample with top_k = 1 (inferer.py:1183-1245) on a
                      small VQVAE + transformer, running past max_seq_len so the sliding window is exercised.
"""
from pathlib import Path

import torch

from tests.golden import configs as G      # before the reference import: /root/reference has its own `tests` package
from oracle import ref_import

OUT = Path(__file__).resolve().parent
TR_KW = dict(num_tokens=11, max_seq_len=16, attn_layers_dim=32, attn_layers_depth=2, attn_layers_heads=4)


def main():
    ref_import.import_reference()
    from generative.inferers import DiffusionInferer, VQVAETransformerInferer
    from generative.networks.nets import VQVAE, DecoderOnlyTransformer, DiffusionModelUNet
    from generative.networks.schedulers import DDPMScheduler
    from generative.utils.ordering import Ordering

    # ---- rank 1: get_likelihood -------------------------------------------------------------------------------
    fx = torch.load(OUT / "g_unet2d.pt", weights_only=False)
    m = DiffusionModelUNet(**fx["kwargs"]).eval()
    m.load_state_dict(fx["state_dict"])
    torch.manual_seed(5)
    x = torch.rand(2, 1, 16, 16) * 2 - 1
    x[0, 0, 0, :4] = torch.tensor([-1.0, 1.0, -0.9995, 0.9995])          # both edge bins of the discretised decoder
    noise = torch.randn(2, 1, 16, 16)
    real_randn_like = torch.randn_like
    cases = {}
    for ptype, clip in (("epsilon", True), ("v_prediction", False)):
        skw = dict(num_train_timesteps=12, prediction_type=ptype, clip_sample=clip)
        s = DDPMScheduler(**skw)
        s.set_timesteps(12)
        torch.randn_like = lambda t: noise.clone()
        try:
            with torch.no_grad():
                lik, inter = DiffusionInferer(s).get_likelihood(inputs=x, diffusion_model=m, scheduler=s,
                                                                save_intermediates=True, verbose=False)
        finally:
            torch.randn_like = real_randn_like
        cases[ptype] = dict(scheduler_kwargs=skw, likelihood=lik, intermediates=[i.clone() for i in inter])
    torch.save(dict(unet_fixture="g_unet2d.pt", x=x, noise=noise, cases=cases), OUT / "g_likelihood.pt")

    # ---- rank 3: transformer forward and greedy sampling ----------------------------------------------------------
    out = {}
    for cross in (False, True):
        torch.manual_seed(0)
        tr = DecoderOnlyTransformer(with_cross_attention=cross, **TR_KW).eval()
        torch.manual_seed(1)
        tok = torch.randint(0, 11, (2, 9))
        ctx = torch.randn(2, 3, 32) if cross else None
        with torch.no_grad():
            logits = tr(tok, context=ctx)
        out["cross" if cross else "plain"] = dict(kwargs=dict(TR_KW, with_cross_attention=cross),
                                                  state_dict=tr.state_dict(), tokens=tok, context=ctx, logits=logits)
    vkw = G.VQVAE_CASES["vqvae2d"]
    K = vkw["num_embeddings"]
    torch.manual_seed(0)
    vq = VQVAE(**vkw).eval()
    skw = dict(num_tokens=K + 1, max_seq_len=10, attn_layers_dim=32, attn_layers_depth=2, attn_layers_heads=4)
    tr = DecoderOnlyTransformer(**skw).eval()
    okw = dict(ordering_type="s_curve", spatial_dims=2, dimensions=(1, 4, 4), reflected_spatial_dims=(True, False))
    ordering = Ordering(**okw)
    start = torch.full((2, 1), K)
    with torch.no_grad():
        sample = VQVAETransformerInferer().sample((4, 4), start, vq, tr, ordering, top_k=1, verbose=False)
    out["sampler"] = dict(vqvae_kwargs=vkw, vqvae_state=vq.state_dict(), transformer_kwargs=skw,
                          transformer_state=tr.state_dict(), ordering_kwargs=okw, start=start, sample=sample)
    torch.save(out, OUT / "g_transformer.pt")
    for f in ("g_likelihood.pt", "g_transformer.pt"):
        print(f, (OUT / f).stat().st_size)


if __name__ == "__main__":
    main()
