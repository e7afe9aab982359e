"""CPU restatement (plain PyTorch fp32, functional) of the reference's diffusion-sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under generativemodels_b200/ imports this file; only tests/, bench.py's
cpu_baseline / --impl reference leg and __graft_entry__.smoke() may, and there only as the checker.

Every function restates one reference function over a ``state_dict`` with the reference's own key names, citing the
file:line it follows (paths relative to /root/reference/generative).  The arithmetic lives in PyTorch ATen exactly as
it does for the reference on CPU (the reference is pure Python over torch; MONAI contributes only thin
wrappers, restated here from their documented semantics).

Pinning: the reference's tests hold no golden tensors for this path, so parity is pinned by
running the UNMODIFIED reference in the build container (oracle/ref_import.py, on the MONAI shim) and
  (a) asserting this restatement reproduces it bit-for-bit / to fp32 round-off on seeded inputs
      (tests/test_oracle_vs_reference.py — runs wherever /root/reference exists), and
  (b) committing the reference's outputs as fixtures under tests/golden/ (tests/golden/make_golden.py) which this
      restatement must reproduce everywhere (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np
import torch
import torch.nn.functional as F

# ======================================================================================================
# helpers
# ======================================================================================================


def _conv(sd, prefix, x, stride=1, padding=0, dilation=1):
    """monai Convolution(conv_only=True): child `conv` = nn.Conv{2,3}d."""
    w = sd[prefix + ".weight"]
    b = sd.get(prefix + ".bias")
    fn = F.conv2d if w.dim() == 4 else F.conv3d
    return fn(x, w, b, stride=stride, padding=padding, dilation=dilation)


def _gn(sd, prefix, x, groups, eps):
    return F.group_norm(x, groups, sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def _linear(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def _ln(sd, prefix, x):
    w = sd[prefix + ".weight"]
    return F.layer_norm(x, (w.shape[0],), w, sd[prefix + ".bias"], 1e-5)


def _has(sd, prefix):
    return any(k.startswith(prefix) for k in sd)


def _count(sd, prefix):
    """number of consecutive integer-indexed children under `prefix` (e.g. 'down_blocks.0.resnets.')."""
    n = 0
    while _has(sd, f"{prefix}{n}."):
        n += 1
    return n


# ======================================================================================================
# DiffusionModelUNet  (networks/nets/diffusion_model_unet.py)
# ======================================================================================================


def get_timestep_embedding(timesteps: torch.Tensor, embedding_dim: int, max_period: int = 10000) -> torch.Tensor:
    """diffusion_model_unet.py:461-485."""
    if timesteps.ndim != 1:
        raise ValueError("Timesteps should be a 1d-array")
    half_dim = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(start=0, end=half_dim, dtype=torch.float32)
    freqs = torch.exp(exponent / half_dim)
    args = timesteps[:, None].float() * freqs[None, :]
    embedding = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if embedding_dim % 2 == 1:
        embedding = F.pad(embedding, (0, 1, 0, 0))
    return embedding


def _heads_to_batch(x, heads):
    b, t, d = x.shape
    return x.reshape(b, t, heads, d // heads).permute(0, 2, 1, 3).reshape(b * heads, t, d // heads)


def _batch_to_heads(x, heads):
    b, t, d = x.shape
    return x.reshape(b // heads, heads, t, d).permute(0, 2, 1, 3).reshape(b // heads, t, d * heads)


def _attention(q, k, v, scale):
    """baddbmm(alpha=scale) -> softmax -> bmm (diffusion_model_unet.py:143-153, 406-416)."""
    scores = torch.baddbmm(torch.empty(q.shape[0], q.shape[1], k.shape[1], dtype=q.dtype), q, k.transpose(-1, -2),
                           beta=0, alpha=scale)
    return torch.bmm(scores.softmax(dim=-1), v)


def attention_block(sd, p, x, groups, eps, num_head_channels):
    """AttentionBlock.forward (diffusion_model_unet.py:418-458; autoencoderkl.py:271-312).  NB proj_attn exists in
    the state_dict but is never applied."""
    residual = x
    b, c = x.shape[:2]
    heads = c // num_head_channels if num_head_channels is not None else 1
    scale = 1 / math.sqrt(c / heads)
    h = _gn(sd, p + ".norm", x, groups, eps)
    h = h.view(b, c, -1).transpose(1, 2)
    q, k, v = (_heads_to_batch(_linear(sd, f"{p}.to_{n}", h), heads) for n in "qkv")
    o = _batch_to_heads(_attention(q, k, v, scale), heads)
    return o.transpose(-1, -2).reshape(x.shape) + residual


def cross_attention(sd, p, x, context, heads, num_head_channels):
    """CrossAttention.forward (diffusion_model_unet.py:155-175); to_q/k/v have no bias, to_out.0 has."""
    scale = 1 / math.sqrt(num_head_channels)
    ctx = x if context is None else context
    q = _heads_to_batch(_linear(sd, p + ".to_q", x), heads)
    k = _heads_to_batch(_linear(sd, p + ".to_k", ctx), heads)
    v = _heads_to_batch(_linear(sd, p + ".to_v", ctx), heads)
    o = _batch_to_heads(_attention(q, k, v, scale), heads)
    return _linear(sd, p + ".to_out.0", o)


def spatial_transformer(sd, p, x, context, groups, eps, num_head_channels):
    """SpatialTransformer.forward (diffusion_model_unet.py:316-342) with BasicTransformerBlock (230-234) and the
    MONAI MLPBlock(act="GEGLU"): linear1 -> a * gelu(gate) -> linear2."""
    residual = x
    h = _gn(sd, p + ".norm", x, groups, eps)
    h = _conv(sd, p + ".proj_in.conv", h)
    inner = h.shape[1]
    heads = inner // num_head_channels
    shape = h.shape
    h = h.view(shape[0], inner, -1).transpose(1, 2)
    for i in range(_count(sd, p + ".transformer_blocks.")):
        bp = f"{p}.transformer_blocks.{i}"
        h = cross_attention(sd, bp + ".attn1", _ln(sd, bp + ".norm1", h), None, heads, num_head_channels) + h
        h = cross_attention(sd, bp + ".attn2", _ln(sd, bp + ".norm2", h), context, heads, num_head_channels) + h
        f = _linear(sd, bp + ".ff.linear1", _ln(sd, bp + ".norm3", h))
        a, gate = f.chunk(2, dim=-1)
        h = _linear(sd, bp + ".ff.linear2", a * F.gelu(gate)) + h
    h = h.transpose(1, 2).reshape(shape)
    h = _conv(sd, p + ".proj_out.conv", h)
    return h + residual


def _upsample_nearest(x):
    return F.interpolate(x, scale_factor=2.0, mode="nearest")


def spade_norm(sd, p, x, seg, groups, eps):
    """SPADE.forward (blocks/spade_norm.py:78-96) with the GROUP base norm the diffusion / autoencoder blocks select.
    ``mlp_gamma`` / ``mlp_beta`` are monai Convolutions built with act=None but the DEFAULT norm="INSTANCE", so each
    is conv -> InstanceNorm (no affine); ``mlp_shared`` is conv -> LeakyReLU(0.01)."""
    normalized = F.group_norm(x, groups, sd.get(p + ".param_free_norm.N.weight"), sd.get(p + ".param_free_norm.N.bias"),
                              eps)
    seg = F.interpolate(seg, size=x.shape[2:], mode="nearest")
    k = sd[p + ".mlp_shared.conv.weight"].shape[-1]
    actv = F.leaky_relu(_conv(sd, p + ".mlp_shared.conv", seg, padding=k // 2), 0.01)
    gamma = F.instance_norm(_conv(sd, p + ".mlp_gamma.conv", actv, padding=k // 2))
    beta = F.instance_norm(_conv(sd, p + ".mlp_beta.conv", actv, padding=k // 2))
    return normalized * (1 + gamma) + beta


def _norm_any(sd, p, x, groups, eps, seg):
    if _has(sd, p + ".mlp_shared."):
        if seg is None:
            raise ValueError("a SPADE block needs the segmentation map")
        return spade_norm(sd, p, x, seg, groups, eps)
    return _gn(sd, p, x, groups, eps)


def resnet_block(sd, p, x, emb, groups, eps, up=False, down=False, seg=None):
    """ResnetBlock.forward (diffusion_model_unet.py:669-696); with SPADE norms (keys ``norm1.mlp_shared...``) it is
    SPADEResnetBlock.forward (spade_diffusion_model_unet.py:173-200)."""
    sdims = x.dim() - 2
    h = F.silu(_norm_any(sd, p + ".norm1", x, groups, eps, seg))
    if up:
        x, h = _upsample_nearest(x), _upsample_nearest(h)
    elif down:
        pool = F.avg_pool2d if sdims == 2 else F.avg_pool3d
        x, h = pool(x, 2, 2), pool(h, 2, 2)
    h = _conv(sd, p + ".conv1.conv", h, padding=1)
    temb = _linear(sd, p + ".time_emb_proj", F.silu(emb))
    h = h + temb[(...,) + (None,) * sdims]
    h = F.silu(_norm_any(sd, p + ".norm2", h, groups, eps, seg))
    h = _conv(sd, p + ".conv2.conv", h, padding=1)
    if _has(sd, p + ".skip_connection."):
        x = _conv(sd, p + ".skip_connection.conv", x)
    return x + h


def _attn_any(sd, p, h, context, groups, eps, nhc):
    if _has(sd, p + ".transformer_blocks."):
        return spatial_transformer(sd, p, h, context, groups, eps, nhc)
    return attention_block(sd, p, h, groups, eps, nhc)


def _down_or_mid_path(sd, cfg, h, emb, context, prefix=""):
    """conv_in output -> (h after mid block, skip list); shared by the UNet and the ControlNet copy of its encoder."""
    groups, eps = cfg.get("norm_num_groups", 32), cfg.get("norm_eps", 1e-6)
    nhcs = cfg["num_head_channels"]
    skips = [h]
    for i in range(_count(sd, prefix + "down_blocks.")):
        bp = f"{prefix}down_blocks.{i}"
        for j in range(_count(sd, bp + ".resnets.")):
            h = resnet_block(sd, f"{bp}.resnets.{j}", h, emb, groups, eps)
            if _has(sd, f"{bp}.attentions.{j}."):
                h = _attn_any(sd, f"{bp}.attentions.{j}", h, context, groups, eps, nhcs[i])
            skips.append(h)
        if _has(sd, bp + ".downsampler.op."):
            h = _conv(sd, bp + ".downsampler.op.conv", h, stride=2, padding=1)   # Downsample (488-531)
            skips.append(h)
        elif _has(sd, bp + ".downsampler."):
            h = resnet_block(sd, bp + ".downsampler", h, emb, groups, eps, down=True)
            skips.append(h)
    mp = prefix + "middle_block"
    h = resnet_block(sd, mp + ".resnet_1", h, emb, groups, eps)
    h = _attn_any(sd, mp + ".attention", h, context, groups, eps, nhcs[-1])
    h = resnet_block(sd, mp + ".resnet_2", h, emb, groups, eps)
    return h, skips


def _time_embedding(sd, cfg, x, timesteps, class_labels, prefix=""):
    """diffusion_model_unet.py:1888-1902."""
    c0 = sd[prefix + "time_embed.0.weight"].shape[1]
    t_emb = get_timestep_embedding(timesteps, c0).to(dtype=x.dtype)
    emb = _linear(sd, prefix + "time_embed.2", F.silu(_linear(sd, prefix + "time_embed.0", t_emb)))
    if (prefix + "class_embedding.weight") in sd:
        if class_labels is None:
            raise ValueError("class_labels should be provided when num_class_embeds > 0")
        emb = emb + F.embedding(class_labels, sd[prefix + "class_embedding.weight"]).to(dtype=x.dtype)
    return emb


def unet_forward(sd, cfg, x, timesteps, context=None, class_labels=None, down_block_additional_residuals=None,
                 mid_block_additional_residual=None, seg=None):
    """DiffusionModelUNet.forward (diffusion_model_unet.py:1869-1943); with ``seg`` and a state_dict whose up-path
    ResnetBlocks carry SPADE norms, SPADEDiffusionModelUNet.forward (spade_diffusion_model_unet.py:836-912).

    cfg: {"num_head_channels": per-level tuple, "norm_num_groups", "norm_eps", "with_conditioning"}.
    """
    groups, eps = cfg.get("norm_num_groups", 32), cfg.get("norm_eps", 1e-6)
    nhcs = cfg["num_head_channels"]
    emb = _time_embedding(sd, cfg, x, timesteps, class_labels)
    h = _conv(sd, "conv_in.conv", x, padding=1)
    if context is not None and not cfg.get("with_conditioning", False):
        raise ValueError("model should have with_conditioning = True if context is provided")
    # down + mid, with ControlNet residuals inserted between them (1917-1932)
    skips = [h]
    nlev = _count(sd, "down_blocks.")
    for i in range(nlev):
        bp = f"down_blocks.{i}"
        for j in range(_count(sd, bp + ".resnets.")):
            h = resnet_block(sd, f"{bp}.resnets.{j}", h, emb, groups, eps)
            if _has(sd, f"{bp}.attentions.{j}."):
                h = _attn_any(sd, f"{bp}.attentions.{j}", h, context, groups, eps, nhcs[i])
            skips.append(h)
        if _has(sd, bp + ".downsampler.op."):
            h = _conv(sd, bp + ".downsampler.op.conv", h, stride=2, padding=1)
            skips.append(h)
        elif _has(sd, bp + ".downsampler."):
            h = resnet_block(sd, bp + ".downsampler", h, emb, groups, eps, down=True)
            skips.append(h)
    if down_block_additional_residuals is not None:
        skips = [s + r for s, r in zip(skips, down_block_additional_residuals)]
    h = resnet_block(sd, "middle_block.resnet_1", h, emb, groups, eps)
    h = _attn_any(sd, "middle_block.attention", h, context, groups, eps, nhcs[-1])
    h = resnet_block(sd, "middle_block.resnet_2", h, emb, groups, eps)
    if mid_block_additional_residual is not None:
        h = h + mid_block_additional_residual
    # up (1935-1938; blocks 1226-1237, 1330-1345, 1451-1466)
    rev_nhc = list(reversed(nhcs))
    for i in range(_count(sd, "up_blocks.")):
        bp = f"up_blocks.{i}"
        for j in range(_count(sd, bp + ".resnets.")):
            h = torch.cat([h, skips.pop()], dim=1)
            h = resnet_block(sd, f"{bp}.resnets.{j}", h, emb, groups, eps, seg=seg)
            if _has(sd, f"{bp}.attentions.{j}."):
                h = _attn_any(sd, f"{bp}.attentions.{j}", h, context, groups, eps, rev_nhc[i])
        if _has(sd, bp + ".upsampler.conv."):
            h = _conv(sd, bp + ".upsampler.conv.conv", _upsample_nearest(h), padding=1)    # Upsample (534-586)
        elif _has(sd, bp + ".upsampler."):
            h = resnet_block(sd, bp + ".upsampler", h, emb, groups, eps, up=True)
    h = F.silu(_gn(sd, "out.0", h, groups, eps))
    return _conv(sd, "out.2.conv", h, padding=1)


# ======================================================================================================
# ControlNet  (networks/nets/controlnet.py)
# ======================================================================================================


def controlnet_forward(sd, cfg, x, timesteps, controlnet_cond, conditioning_scale=1.0, context=None,
                       class_labels=None):
    """ControlNet.forward (controlnet.py:367-436) -> (down residual list, mid residual)."""
    emb = _time_embedding(sd, cfg, x, timesteps, class_labels)
    h = _conv(sd, "conv_in.conv", x, padding=1)
    # ControlNetConditioningEmbedding (controlnet.py:45-116): conv_in, SiLU, [conv s1, SiLU, conv s2, SiLU]*, conv_out
    c = F.silu(_conv(sd, "controlnet_cond_embedding.conv_in.conv", controlnet_cond, padding=1))
    nb = _count(sd, "controlnet_cond_embedding.blocks.")
    for i in range(nb):
        stride = 2 if i % 2 == 1 else 1
        c = F.silu(_conv(sd, f"controlnet_cond_embedding.blocks.{i}.conv", c, stride=stride, padding=1))
    c = _conv(sd, "controlnet_cond_embedding.conv_out.conv", c, padding=1)
    h = h + c
    h, skips = _down_or_mid_path(sd, cfg, h, emb, context)
    outs = []
    for i, s in enumerate(skips):
        key = f"controlnet_down_blocks.{i}.conv" if f"controlnet_down_blocks.{i}.conv.weight" in sd \
            else f"controlnet_down_blocks.{i}"
        outs.append(_conv(sd, key, s) * conditioning_scale)
    mid = _conv(sd, "controlnet_mid_block.conv", h) * conditioning_scale
    return outs, mid


# ======================================================================================================
# AutoencoderKL  (networks/nets/autoencoderkl.py)
# ======================================================================================================


def _ae_resblock(sd, p, x, groups, eps, seg=None):
    """ResBlock.forward (autoencoderkl.py:179-193); SPADEResBlock.forward (spade_autoencoderkl.py:122-134) when the
    norms are SPADE blocks — those build their GroupNorm from {"num_groups", "affine": False} only, so it runs with
    PyTorch's default eps = 1e-5 rather than the network's norm_eps."""
    if _has(sd, p + ".norm1.mlp_shared."):
        if seg is None:
            raise ValueError("a SPADE block needs the segmentation map")
        h = _conv(sd, p + ".conv1.conv", F.silu(spade_norm(sd, p + ".norm1", x, seg, groups, 1e-5)), padding=1)
        h = _conv(sd, p + ".conv2.conv", F.silu(spade_norm(sd, p + ".norm2", h, seg, groups, 1e-5)), padding=1)
    else:
        h = _conv(sd, p + ".conv1.conv", F.silu(_gn(sd, p + ".norm1", x, groups, eps)), padding=1)
        h = _conv(sd, p + ".conv2.conv", F.silu(_gn(sd, p + ".norm2", h, groups, eps)), padding=1)
    if _has(sd, p + ".nin_shortcut."):
        x = _conv(sd, p + ".nin_shortcut.conv", x)
    return x + h


def _ae_blocks(sd, prefix, x, groups, eps, decoder: bool, seg=None):
    """Encoder.forward / Decoder.forward: walk `blocks` by the kind of parameters each holds (315-452, 455-597)."""
    sdims = x.dim() - 2
    n = _count(sd, prefix + "blocks.")
    for i in range(n):
        p = f"{prefix}blocks.{i}"
        if _has(sd, p + ".norm1.") or _has(sd, p + ".conv1."):
            x = _ae_resblock(sd, p, x, groups, eps, seg)
        elif _has(sd, p + ".to_q."):
            x = attention_block(sd, p, x, groups, eps, None)
        elif (p + ".weight") in sd and sd[p + ".weight"].dim() == 1:
            x = _gn(sd, p, x, groups, eps)                                  # bare GroupNorm, no activation
        elif (p + ".conv.conv.weight") in sd:
            w = sd[p + ".conv.conv.weight"]
            if decoder:                                                    # Upsample (autoencoderkl.py:41-93)
                if cfg_is_transposed(sd, p, sdims):
                    fn = F.conv_transpose2d if sdims == 2 else F.conv_transpose3d
                    x = fn(x, w, sd.get(p + ".conv.conv.bias"), stride=2, padding=1, output_padding=1)
                else:
                    x = _conv(sd, p + ".conv.conv", _upsample_nearest(x), padding=1)
            else:                                                          # Downsample (96-122): pad (0,1) then s2 p0
                x = _conv(sd, p + ".conv.conv", F.pad(x, (0, 1) * sdims), stride=2, padding=0)
        else:
            x = _conv(sd, p + ".conv", x, padding=1)                       # plain k3 conv
    return x


def cfg_is_transposed(sd, p, sdims):
    return bool(sd.get("__use_convtranspose__", False))


def autoencoderkl_encode(sd, cfg, x):
    """AutoencoderKL.encode (autoencoderkl.py:718-736) -> (z_mu, z_sigma)."""
    groups, eps = cfg.get("norm_num_groups", 32), cfg.get("norm_eps", 1e-6)
    h = _ae_blocks(sd, "encoder.", x, groups, eps, decoder=False)
    z_mu = _conv(sd, "quant_conv_mu.conv", h)
    z_log_var = torch.clamp(_conv(sd, "quant_conv_log_sigma.conv", h), -30.0, 20.0)
    return z_mu, torch.exp(z_log_var / 2)


def autoencoderkl_decode(sd, cfg, z, seg=None):
    """AutoencoderKL.decode (autoencoderkl.py:769-784); with ``seg``, SPADEAutoencoderKL.decode
    (spade_autoencoderkl.py:457-469)."""
    groups, eps = cfg.get("norm_num_groups", 32), cfg.get("norm_eps", 1e-6)
    sd = dict(sd)
    sd["__use_convtranspose__"] = cfg.get("use_convtranspose", False)
    return _ae_blocks(sd, "decoder.", _conv(sd, "post_quant_conv.conv", z), groups, eps, decoder=True, seg=seg)


# ======================================================================================================
# VQVAE + VectorQuantizer  (networks/nets/vqvae.py, networks/layers/vector_quantizer.py)
# ======================================================================================================


def _vq_res_unit(sd, p, x):
    """VQVAEResidualUnit.forward (vqvae.py:79-80): relu(x + conv2(relu(conv1(x)))) (dropout p=0 in eval)."""
    h = F.relu(_conv(sd, p + ".conv1.conv", x, padding=1))
    return F.relu(x + _conv(sd, p + ".conv2.conv", h, padding=1))


def vqvae_encode(sd, cfg, x):
    """Encoder.forward (vqvae.py:83-170); cfg["downsample_parameters"] = ((stride, k, dilation, pad), ...)."""
    n = _count(sd, "encoder.blocks.")
    lvl = 0
    for i in range(n):
        p = f"encoder.blocks.{i}"
        if _has(sd, p + ".conv1."):
            x = _vq_res_unit(sd, p, x)
        elif i == n - 1:
            x = _conv(sd, p + ".conv", x, padding=1)
        else:
            s, k, d, pad = cfg["downsample_parameters"][lvl]
            x = F.relu(_conv(sd, p + ".conv", x, stride=s, padding=pad, dilation=d))
            lvl += 1
    return x


def vqvae_decode(sd, cfg, x):
    """Decoder.forward (vqvae.py:173-271); cfg["upsample_parameters"] = ((stride, k, dilation, pad, out_pad), ...)."""
    n = _count(sd, "decoder.blocks.")
    sdims = x.dim() - 2
    ups = cfg["upsample_parameters"]
    lvl = 0
    for i in range(n):
        p = f"decoder.blocks.{i}"
        if _has(sd, p + ".conv1."):
            x = _vq_res_unit(sd, p, x)
        elif i == 0:
            x = _conv(sd, p + ".conv", x, padding=1)
        else:
            s, k, d, pad, opad = ups[lvl]
            fn = F.conv_transpose2d if sdims == 2 else F.conv_transpose3d
            x = fn(x, sd[p + ".conv.weight"], sd.get(p + ".conv.bias"), stride=s, padding=pad, output_padding=opad,
                   dilation=d)
            if lvl != len(ups) - 1:
                x = F.relu(x)
            lvl += 1
    if cfg.get("output_act"):
        x = {"relu": F.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid}[str(cfg["output_act"]).lower()](x)
    return x


def vq_quantize(codebook: torch.Tensor, z: torch.Tensor):
    """EMAQuantizer.quantize (vector_quantizer.py:86-122): -> (flat_input [M, D], indices [B, *spatial])."""
    sdims = z.dim() - 2
    perm = [0] + list(range(2, sdims + 2)) + [1]
    z = z.float()
    flat = z.permute(perm).contiguous().view(-1, codebook.shape[1])
    distances = ((flat ** 2).sum(dim=1, keepdim=True) + (codebook.t() ** 2).sum(dim=0, keepdim=True)
                 - 2 * torch.mm(flat, codebook.t()))
    idx = torch.max(-distances, dim=1)[1]
    view = list(z.shape)
    del view[1]
    return flat, idx.view(view)


def vq_embed(codebook: torch.Tensor, idx: torch.Tensor):
    """EMAQuantizer.embed (vector_quantizer.py:124-138)."""
    sdims = idx.dim() - 1
    perm = [0, sdims + 1] + list(range(1, sdims + 1))
    return F.embedding(idx, codebook).permute(perm).contiguous()


def vq_forward(codebook, z, commitment_cost=0.25):
    """EMAQuantizer.forward in eval mode (vector_quantizer.py:161-188) + VectorQuantizer.forward perplexity
    (208-220): -> (quantized (straight-through values), loss, indices, perplexity)."""
    _, idx = vq_quantize(codebook, z)
    q = vq_embed(codebook, idx)
    loss = commitment_cost * F.mse_loss(q, z)
    q_st = z + (q - z)
    K = codebook.shape[0]
    avg = torch.histc(idx.float(), bins=K, max=K).float().div(idx.numel())
    perplexity = torch.exp(-torch.sum(avg * torch.log(avg + 1e-10)))
    return q_st, loss, idx, perplexity


def vq_index_margin(codebook: torch.Tensor, flat: torch.Tensor) -> torch.Tensor:
    """fp64 gap between the best and second-best code of each vector — used by the parity tests to tell a genuine
    mismatch from an fp32 near-tie (SURVEY.md §7 'VQ index bit-exactness')."""
    d = torch.cdist(flat.double(), codebook.double()) ** 2
    s = torch.sort(d, dim=1)[0]
    return s[:, 1] - s[:, 0] if d.shape[1] > 1 else torch.full((d.shape[0],), float("inf"), dtype=torch.float64)


def vqvae_forward(sd, cfg, x):
    """VQVAE.forward (vqvae.py:436-439) -> (reconstruction, loss, indices)."""
    cb = sd["quantizer.quantizer.embedding.weight"]
    z = vqvae_encode(sd, cfg, x)
    q, loss, idx, _ = vq_forward(cb, z, cfg.get("commitment_cost", 0.25))
    return vqvae_decode(sd, cfg, q), loss, idx


# ======================================================================================================
# Schedulers  (networks/schedulers/{scheduler,ddpm,ddim,pndm}.py)
# ======================================================================================================


def noise_schedule(name: str, num_train_timesteps: int, **kw) -> torch.Tensor:
    """NoiseSchedules (scheduler.py:40-110)."""
    if name == "linear_beta":
        return torch.linspace(kw.get("beta_start", 1e-4), kw.get("beta_end", 2e-2), num_train_timesteps,
                              dtype=torch.float32)
    if name == "scaled_linear_beta":
        return torch.linspace(kw.get("beta_start", 1e-4) ** 0.5, kw.get("beta_end", 2e-2) ** 0.5,
                              num_train_timesteps, dtype=torch.float32) ** 2
    if name == "sigmoid_beta":
        sig_range = kw.get("sig_range", 6)
        betas = torch.linspace(-sig_range, sig_range, num_train_timesteps)
        return torch.sigmoid(betas) * (kw.get("beta_end", 2e-2) - kw.get("beta_start", 1e-4)) + kw.get("beta_start", 1e-4)
    if name == "cosine":
        s = kw.get("s", 8e-3)
        x = torch.linspace(0, num_train_timesteps, num_train_timesteps + 1)
        ac = torch.cos(((x / num_train_timesteps) + s) / (1 + s) * torch.pi * 0.5) ** 2
        ac /= ac[0].item()
        alphas = torch.clip(ac[1:] / ac[:-1], 0.0001, 0.9999)
        return 1.0 - alphas, alphas, ac[:-1]          # the cosine schedule returns the triple (scheduler.py:83-89)
    raise ValueError(f"unknown schedule {name}")


class SchedulerTables:
    """Scheduler.__init__ (scheduler.py:149-167)."""

    def __init__(self, num_train_timesteps=1000, schedule="linear_beta", **schedule_args):
        self.num_train_timesteps = num_train_timesteps
        ns = noise_schedule(schedule, num_train_timesteps, **schedule_args)
        if isinstance(ns, tuple):
            self.betas, self.alphas, self.alphas_cumprod = ns
        else:
            self.betas = ns
            self.alphas = 1.0 - self.betas
            self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)


class DDIMOracle(SchedulerTables):
    """DDIMScheduler (ddim.py:78-237)."""

    def __init__(self, num_train_timesteps=1000, schedule="linear_beta", clip_sample=True, set_alpha_to_one=True,
                 steps_offset=0, prediction_type="epsilon", clip_sample_min=-1, clip_sample_max=1, **schedule_args):
        super().__init__(num_train_timesteps, schedule, **schedule_args)
        self.prediction_type = prediction_type
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.clip_sample = clip_sample
        self.clip_sample_values = [clip_sample_min, clip_sample_max]
        self.steps_offset = steps_offset
        self.set_timesteps(num_train_timesteps)

    def set_timesteps(self, n):
        if n > self.num_train_timesteps:
            raise ValueError("num_inference_steps cannot be larger than num_train_timesteps")
        self.num_inference_steps = n
        ratio = self.num_train_timesteps // n
        ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts) + self.steps_offset

    def step(self, model_output, timestep, sample, eta=0.0, generator=None):
        prev_t = timestep - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        if self.prediction_type == "epsilon":
            x0 = (sample - (b_t ** 0.5) * model_output) / (a_t ** 0.5)
            eps = model_output
        elif self.prediction_type == "sample":
            x0 = model_output
            eps = (sample - (a_t ** 0.5) * x0) / (b_t ** 0.5)
        else:
            x0 = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
            eps = (a_t ** 0.5) * model_output + (b_t ** 0.5) * sample
        if self.clip_sample:
            x0 = torch.clamp(x0, self.clip_sample_values[0], self.clip_sample_values[1])
        variance = ((1 - a_prev) / (1 - a_t)) * (1 - a_t / a_prev)
        std = eta * variance ** 0.5
        prev = a_prev ** 0.5 * x0 + (1 - a_prev - std ** 2) ** 0.5 * eps
        if eta > 0:
            noise = torch.randn(model_output.shape, dtype=model_output.dtype, generator=generator)
            prev = prev + variance ** 0.5 * eta * noise
        return prev, x0


class DDPMOracle(SchedulerTables):
    """DDPMScheduler (ddpm.py:66-252), fixed variance types."""

    def __init__(self, num_train_timesteps=1000, schedule="linear_beta", variance_type="fixed_small",
                 clip_sample=True, prediction_type="epsilon", clip_sample_min=-1, clip_sample_max=1, **schedule_args):
        super().__init__(num_train_timesteps, schedule, **schedule_args)
        self.variance_type, self.clip_sample, self.prediction_type = variance_type, clip_sample, prediction_type
        self.clip_sample_values = [clip_sample_min, clip_sample_max]
        self.set_timesteps(num_train_timesteps)

    def set_timesteps(self, n):
        if n > self.num_train_timesteps:
            raise ValueError("num_inference_steps cannot be larger than num_train_timesteps")
        self.num_inference_steps = n
        ratio = self.num_train_timesteps // n
        self.timesteps = torch.from_numpy((np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64))

    def _variance(self, t):
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[t - 1] if t > 0 else self.one
        variance = (1 - a_prev) / (1 - a_t) * self.betas[t]
        if self.variance_type == "fixed_small":
            variance = torch.clamp(variance, min=1e-20)
        elif self.variance_type == "fixed_large":
            variance = self.betas[t]
        return variance

    def step(self, model_output, timestep, sample, generator=None):
        t = timestep
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[t - 1] if t > 0 else self.one
        b_t, b_prev = 1 - a_t, 1 - a_prev
        if self.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        elif self.prediction_type == "sample":
            x0 = model_output
        else:
            x0 = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
        if self.clip_sample:
            x0 = torch.clamp(x0, self.clip_sample_values[0], self.clip_sample_values[1])
        c0 = (a_prev ** 0.5 * self.betas[t]) / b_t
        ct = self.alphas[t] ** 0.5 * b_prev / b_t
        prev = c0 * x0 + ct * sample
        if t > 0:
            noise = torch.randn(model_output.size(), dtype=model_output.dtype, generator=generator)
            prev = prev + (self._variance(t) ** 0.5) * noise
        return prev, x0


class PNDMOracle(SchedulerTables):
    """PNDMScheduler (pndm.py:78-317)."""

    def __init__(self, num_train_timesteps=1000, schedule="linear_beta", skip_prk_steps=False,
                 set_alpha_to_one=False, prediction_type="epsilon", steps_offset=0, **schedule_args):
        super().__init__(num_train_timesteps, schedule, **schedule_args)
        self.prediction_type = prediction_type
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.pndm_order, self.skip_prk_steps, self.steps_offset = 4, skip_prk_steps, steps_offset
        self.cur_model_output, self.counter, self.cur_sample, self.ets = 0, 0, None, []
        self.set_timesteps(num_train_timesteps)

    def set_timesteps(self, n):
        if n > self.num_train_timesteps:
            raise ValueError("num_inference_steps cannot be larger than num_train_timesteps")
        self.num_inference_steps = n
        ratio = self.num_train_timesteps // n
        self._timesteps = (np.arange(0, n) * ratio).round().astype(np.int64) + self.steps_offset
        if self.skip_prk_steps:
            self.prk_timesteps = np.array([])
            self.plms_timesteps = self._timesteps[::-1]
        else:
            prk = np.array(self._timesteps[-self.pndm_order:]).repeat(2) + np.tile(
                np.array([0, self.num_train_timesteps // n // 2]), self.pndm_order)
            self.prk_timesteps = (prk[:-1].repeat(2)[1:-1])[::-1].copy()
            self.plms_timesteps = self._timesteps[:-3][::-1].copy()
        ts = np.concatenate([self.prk_timesteps, self.plms_timesteps]).astype(np.int64)
        self.timesteps = torch.from_numpy(ts)
        self.num_inference_steps = len(self.timesteps)
        self.ets, self.counter = [], 0

    def step(self, model_output, timestep, sample):
        if self.counter < len(self.prk_timesteps) and not self.skip_prk_steps:
            return self.step_prk(model_output, timestep, sample), None
        return self.step_plms(model_output, timestep, sample), None

    def step_prk(self, model_output, timestep, sample):
        diff = 0 if self.counter % 2 else self.num_train_timesteps // self.num_inference_steps // 2
        prev_t = timestep - diff
        timestep = self.prk_timesteps[self.counter // 4 * 4]
        if self.counter % 4 == 0:
            self.cur_model_output += 1 / 6 * model_output
            self.ets.append(model_output)
            self.cur_sample = sample
        elif (self.counter - 1) % 4 == 0:
            self.cur_model_output += 1 / 3 * model_output
        elif (self.counter - 2) % 4 == 0:
            self.cur_model_output += 1 / 3 * model_output
        elif (self.counter - 3) % 4 == 0:
            model_output = self.cur_model_output + 1 / 6 * model_output
            self.cur_model_output = 0
        cur = self.cur_sample if self.cur_sample is not None else sample
        prev = self._prev(cur, timestep, prev_t, model_output)
        self.counter += 1
        return prev

    def step_plms(self, model_output, timestep, sample):
        prev_t = timestep - self.num_train_timesteps // self.num_inference_steps
        if self.counter != 1:
            self.ets = self.ets[-3:]
            self.ets.append(model_output)
        else:
            prev_t = timestep
            timestep = timestep + self.num_train_timesteps // self.num_inference_steps
        if len(self.ets) == 1 and self.counter == 0:
            self.cur_sample = sample
        elif len(self.ets) == 1 and self.counter == 1:
            model_output = (model_output + self.ets[-1]) / 2
            sample = self.cur_sample
            self.cur_sample = None
        elif len(self.ets) == 2:
            model_output = (3 * self.ets[-1] - self.ets[-2]) / 2
        elif len(self.ets) == 3:
            model_output = (23 * self.ets[-1] - 16 * self.ets[-2] + 5 * self.ets[-3]) / 12
        else:
            model_output = (1 / 24) * (55 * self.ets[-1] - 59 * self.ets[-2] + 37 * self.ets[-3] - 9 * self.ets[-4])
        prev = self._prev(sample, timestep, prev_t, model_output)
        self.counter += 1
        return prev

    def _prev(self, sample, timestep, prev_t, model_output):
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t, b_prev = 1 - a_t, 1 - a_prev
        if self.prediction_type == "v_prediction":
            model_output = (a_t ** 0.5) * model_output + (b_t ** 0.5) * sample
        sample_coeff = (a_prev / a_t) ** 0.5
        denom = a_t * b_prev ** 0.5 + (a_t * b_t * a_prev) ** 0.5
        return sample_coeff * sample - (a_prev - a_t) * model_output / denom


# ======================================================================================================
# Inferers  (inferers/inferer.py)
# ======================================================================================================


def diffusion_sample(model_fn, scheduler, input_noise, conditioning=None, mode="crossattn", generator=None):
    """DiffusionInferer.sample (inferer.py:83-143) over a ``model_fn(x, timesteps, context)`` closure."""
    if mode not in ("crossattn", "concat"):
        raise NotImplementedError(f"{mode} condition is not supported")
    image = input_noise
    for t in scheduler.timesteps:
        ts = torch.Tensor((t,))
        if mode == "concat":
            out = model_fn(torch.cat([image, conditioning], dim=1), ts, None)
        else:
            out = model_fn(image, ts, conditioning)
        if isinstance(scheduler, DDPMOracle):
            image, _ = scheduler.step(out, int(t), image, generator=generator)
        else:
            image, _ = scheduler.step(out, int(t), image)
    return image


def controlnet_sample(unet_fn, controlnet_fn, scheduler, input_noise, cn_cond, conditioning=None):
    """ControlNetDiffusionInferer.sample (inferer.py:632-707), crossattn mode."""
    image = input_noise
    for t in scheduler.timesteps:
        ts = torch.Tensor((t,))
        down, mid = controlnet_fn(image, ts, cn_cond, conditioning)
        out = unet_fn(image, ts, conditioning, down, mid)
        image, _ = scheduler.step(out, int(t), image)
    return image


def approx_standard_normal_cdf(x):
    """inferer.py:279-283."""
    return 0.5 * (1.0 + torch.tanh(torch.sqrt(torch.Tensor([2.0 / math.pi])) * (x + 0.044715 * torch.pow(x, 3))))


def decoder_log_likelihood(inputs, means, log_scales, original_input_range=(0, 255), scaled_input_range=(0, 1)):
    """DiffusionInferer._get_decoder_log_likelihood (inferer.py:285-321)."""
    bin_width = (scaled_input_range[1] - scaled_input_range[0]) / (original_input_range[1] - original_input_range[0])
    centered_x = inputs - means
    inv_stdv = torch.exp(-log_scales)
    cdf_plus = approx_standard_normal_cdf(inv_stdv * (centered_x + bin_width / 2))
    cdf_min = approx_standard_normal_cdf(inv_stdv * (centered_x - bin_width / 2))
    log_cdf_plus = torch.log(cdf_plus.clamp(min=1e-12))
    log_one_minus_cdf_min = torch.log((1.0 - cdf_min).clamp(min=1e-12))
    cdf_delta = cdf_plus - cdf_min
    return torch.where(inputs < -0.999, log_cdf_plus,
                       torch.where(inputs > 0.999, log_one_minus_cdf_min, torch.log(cdf_delta.clamp(min=1e-12))))


def get_likelihood(model_fn, scheduler: DDPMOracle, inputs, noise, conditioning=None,
                   original_input_range=(0, 255), scaled_input_range=(0, 1)):
    """DiffusionInferer.get_likelihood (inferer.py:145-277), crossattn mode, fixed variance; ``noise`` is passed in
    (the reference draws torch.randn_like(inputs)) so that both sides of a parity test see the same draw."""
    total_kl = torch.zeros(inputs.shape[0])
    acp = scheduler.alphas_cumprod
    for t in scheduler.timesteps:
        t = int(t)
        ts = torch.full(inputs.shape[:1], t).long()
        sa = (acp[ts] ** 0.5)[(...,) + (None,) * (inputs.dim() - 1)]
        sb = ((1 - acp[ts]) ** 0.5)[(...,) + (None,) * (inputs.dim() - 1)]
        noisy = sa * inputs + sb * noise                                  # Scheduler.add_noise (scheduler.py:169-189)
        out = model_fn(noisy, ts, conditioning)
        a_t = acp[t]
        a_prev = acp[t - 1] if t > 0 else scheduler.one
        b_t, b_prev = 1 - a_t, 1 - a_prev
        if scheduler.prediction_type == "epsilon":
            x0 = (noisy - b_t ** 0.5 * out) / a_t ** 0.5
        elif scheduler.prediction_type == "sample":
            x0 = out
        else:
            x0 = (a_t ** 0.5) * noisy - (b_t ** 0.5) * out
        if scheduler.clip_sample:
            x0 = torch.clamp(x0, -1, 1)
        c0 = (a_prev ** 0.5 * scheduler.betas[t]) / b_t
        ct = scheduler.alphas[t] ** 0.5 * b_prev / b_t
        predicted_mean = c0 * x0 + ct * noisy
        posterior_mean = c0 * inputs + ct * noisy                           # DDPMScheduler._get_mean (ddpm.py:133-156)
        log_post = torch.log(scheduler._variance(t))
        log_pred = log_post
        if t == 0:
            kl = -decoder_log_likelihood(inputs, predicted_mean, 0.5 * log_pred, original_input_range,
                                         scaled_input_range)
        else:
            kl = 0.5 * (-1.0 + log_pred - log_post + torch.exp(log_post - log_pred)
                        + ((posterior_mean - predicted_mean) ** 2) * torch.exp(-log_pred))
        total_kl += kl.view(kl.shape[0], -1).mean(axis=1)
    return total_kl


# ======================================================================================================
# DecoderOnlyTransformer + VQVAETransformerInferer  (nets/transformer.py, blocks/{selfattention,transformerblock}.py,
# utils/ordering.py, inferers/inferer.py:1126-1330) — SURVEY.md §8f rank 3
# ======================================================================================================
def _sa_block(sd, p, x, heads, causal, context=None):
    """SABlock.forward (blocks/selfattention.py:101-148), non-xformers branch."""
    b, t, c = x.shape
    kv = context if context is not None else x
    q = F.linear(x, sd[p + ".to_q.weight"], sd.get(p + ".to_q.bias"))
    k = F.linear(kv, sd[p + ".to_k.weight"], sd.get(p + ".to_k.bias"))
    v = F.linear(kv, sd[p + ".to_v.weight"], sd.get(p + ".to_v.bias"))
    kv_t = kv.shape[1]
    hs = c // heads
    q = q.view(b, t, heads, hs).transpose(1, 2) * (1.0 / math.sqrt(hs))
    k = k.view(b, kv_t, heads, hs).transpose(1, 2)
    v = v.view(b, kv_t, heads, hs).transpose(1, 2)
    s = q @ k.transpose(-2, -1)
    if causal:
        mask = torch.tril(torch.ones(t, kv_t)).view(1, 1, t, kv_t)
        s = s.masked_fill(mask == 0, float("-inf"))
    y = (F.softmax(s, dim=-1) @ v).transpose(1, 2).contiguous().view(b, t, c)
    return F.linear(y, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


def transformer_forward(sd, heads, x, context=None):
    """DecoderOnlyTransformer.forward (nets/transformer.py:96-106) over TransformerBlocks
    (blocks/transformerblock.py:87-92): pre-LN causal self-attention, optional cross-attention, GELU MLP."""
    b, t = x.shape
    h = F.embedding(x, sd["token_embeddings.weight"]) + F.embedding(torch.arange(t).repeat(b, 1),
                                                                  sd["position_embeddings.embedding.weight"])
    for i in range(_count(sd, "blocks.")):
        p = f"blocks.{i}"
        h = h + _sa_block(sd, p + ".attn", _ln(sd, p + ".norm1", h), heads, True)
        if _has(sd, p + ".cross_attn."):
            h = h + _sa_block(sd, p + ".cross_attn", _ln(sd, p + ".norm2", h), heads, False, context=context)
        m = _ln(sd, p + ".norm3", h)
        m = _linear(sd, p + ".mlp.linear2", F.gelu(_linear(sd, p + ".mlp.linear1", m)))
        h = h + m
    return _linear(sd, "to_logits", h)


def sequence_ordering(ordering_type, spatial_dims, dimensions, reflected_spatial_dims=(), transpositions_axes=(),
                      rot90_axes=(), transformation_order=("transpose", "rotate_90", "reflect")):
    """utils/ordering.py: index template -> transformations in the requested order -> raster / s-curve scan
    (``random`` draws from numpy's global generator and is not restated)."""
    import numpy as np
    template = np.arange(int(np.prod(dimensions[1:]))).reshape(*dimensions[1:])
    for tr in transformation_order:
        if tr == "transpose":
            for axes in transpositions_axes:
                template = np.transpose(template, axes=axes)
        elif tr == "rotate_90":
            for axes in rot90_axes:
                template = np.rot90(template, axes=axes)
        elif tr == "reflect":
            for axis, flag in enumerate(reflected_spatial_dims):
                template = np.flip(template, axis=axis) if flag else template
    shp = template.shape
    out = []
    for r in range(shp[0]):
        cols = range(shp[1]) if (ordering_type == "raster_scan" or r % 2 == 0) else range(shp[1] - 1, -1, -1)
        for c in cols:
            if spatial_dims == 3:
                deps = range(shp[2]) if (ordering_type == "raster_scan" or c % 2 == 0) else range(shp[2] - 1, -1, -1)
                out.extend(template[r, c, d] for d in deps)
            else:
                out.append(template[r, c])
    return np.array(out)


def transformer_sample_greedy(sd, heads, max_seq_len, bos, seq_len, batch, context=None):
    """VQVAETransformerInferer.sample (inferer.py:1183-1245) with top_k = 1 (the multinomial draw over a one-hot
    distribution is deterministic): returns the token sequence without the BOS, before the ordering is reverted."""
    seq = torch.full((batch, 1), bos, dtype=torch.long)
    for _ in range(seq_len):
        cond = seq if seq.shape[1] <= max_seq_len else seq[:, -max_seq_len:]
        logits = transformer_forward(sd, heads, cond, context)[:, -1, :]
        logits[:, bos] = -float("inf")            # probs[:, num_embeddings] = 0 in the reference
        seq = torch.cat([seq, logits.argmax(-1, keepdim=True)], dim=1)
    return seq[:, 1:]


# ======================================================================================================
# brain-LDM bundle scripts  (model-zoo/models/brain_image_synthesis_latent_diffusion_model/scripts)
# ======================================================================================================


def bundle_sampling_fn(model_fn, decode_fn, scheduler, input_noise, conditioning):
    """Sampler.sampling_fn (scripts/sampler.py:17-52): the conditioning vector is both broadcast to planes that are
    concatenated to the latent and passed as cross-attention context; long timesteps; decode_stage_2_outputs at the
    end (the CUDA autocast around it is a no-op on the CPU)."""
    image = input_noise
    cond_concat = conditioning.squeeze(1).unsqueeze(-1).unsqueeze(-1).unsqueeze(-1)
    cond_concat = cond_concat.expand(list(cond_concat.shape[0:2]) + list(input_noise.shape[2:]))
    for t in scheduler.timesteps:
        out = model_fn(torch.cat((image, cond_concat), dim=1), torch.Tensor((t,)).long(), conditioning)
        image, _ = scheduler.step(out, int(t), image)
    return decode_fn(image)


def bundle_nifti_quantise(image_data: np.ndarray) -> np.ndarray:
    """NiftiSaver.save up to the nibabel call (scripts/saver.py:22-25)."""
    image_data = image_data[0, 0, 5:-5, 5:-5, :-15]
    image_data = (image_data - image_data.min()) / (image_data.max() - image_data.min())
    return (image_data * 255).astype(np.uint8)
