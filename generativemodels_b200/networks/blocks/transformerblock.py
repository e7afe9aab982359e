"""TransformerBlock — ``generative/networks/blocks/transformerblock.py:19-92``: pre-LayerNorm causal self-attention,
optional cross-attention, GELU MLP (monai ``MLPBlock``: linear1 -> GELU -> linear2), residual around each.  The GELU
runs in the first GEMM's epilogue and every residual add in the epilogue of the GEMM that produces the branch."""
from __future__ import annotations

import torch.nn as nn

from ... import ops
from ...ops import ACT_GELU, CL
from .._holders import f32, packed_linear
from .selfattention import SABlock


class MLPBlock(nn.Module):
    """Key layout of ``monai.networks.blocks.MLPBlock(hidden, mlp_dim, act="GELU")``."""

    def __init__(self, hidden_size: int, mlp_dim: int, dropout_rate: float = 0.0) -> None:
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        mlp_dim = mlp_dim or hidden_size
        self.linear1 = nn.Linear(hidden_size, mlp_dim)
        self.linear2 = nn.Linear(mlp_dim, hidden_size)

    def forward(self, x: CL, residual: CL | None = None) -> CL:
        h = ops.linear(x, packed_linear(self, "linear1"), act1=ACT_GELU)
        return ops.linear(h, packed_linear(self, "linear2"), residual=residual)


class TransformerBlock(nn.Module):
    def __init__(self, hidden_size: int, mlp_dim: int, num_heads: int, dropout_rate: float = 0.0,
                 qkv_bias: bool = False, causal: bool = False, sequence_length: int | None = None,
                 with_cross_attention: bool = False, use_flash_attention: bool = False) -> None:
        self.with_cross_attention = with_cross_attention
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise ValueError("hidden_size should be divisible by num_heads.")
        self.norm1 = nn.LayerNorm(hidden_size)
        self.attn = SABlock(hidden_size, num_heads, dropout_rate, qkv_bias, causal, sequence_length,
                            use_flash_attention=use_flash_attention)
        self.norm2 = None
        self.cross_attn = None
        if self.with_cross_attention:
            self.norm2 = nn.LayerNorm(hidden_size)
            self.cross_attn = SABlock(hidden_size, num_heads, dropout_rate, qkv_bias,
                                      with_cross_attention=with_cross_attention, causal=False,
                                      use_flash_attention=use_flash_attention)
        self.norm3 = nn.LayerNorm(hidden_size)
        self.mlp = MLPBlock(hidden_size, mlp_dim, dropout_rate)

    @staticmethod
    def _ln(norm: nn.LayerNorm, x: CL) -> CL:
        return ops.layernorm(x, f32(norm.weight), f32(norm.bias), norm.eps)

    def forward(self, x: CL, B: int, T: int, context: CL | None = None, context_len: int = 0) -> CL:
        x = self.attn(self._ln(self.norm1, x), B, T, residual=x)
        if self.with_cross_attention:
            x = self.cross_attn(self._ln(self.norm2, x), B, T, context=context, context_len=context_len, residual=x)
        return self.mlp(self._ln(self.norm3, x), residual=x)
