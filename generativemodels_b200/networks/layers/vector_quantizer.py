"""EMAQuantizer / VectorQuantizer with the reference's interface (generative/networks/layers/vector_quantizer.py).

Eval mode (the sampling path): one CUDA kernel does the nearest-code search, gather, commitment-loss numerator and
code histogram (vq.cu) — no M x K distance or one-hot tensors.  Training mode keeps the reference's PyTorch composite
(EMA codebook update and the optional all_reduce, vector_quantizer.py:140-180): SURVEY.md §8(a15) leaves it there.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
from torch import nn

from ... import _lib, ops
from ...ops import CL

__all__ = ["VectorQuantizer", "EMAQuantizer"]


class EMAQuantizer(nn.Module):
    """vector_quantizer.py:20-188."""

    def __init__(self, spatial_dims: int, num_embeddings: int, embedding_dim: int, commitment_cost: float = 0.25,
                 decay: float = 0.99, epsilon: float = 1e-5, embedding_init: str = "normal", ddp_sync: bool = True):
        super().__init__()
        self.spatial_dims: int = spatial_dims
        self.embedding_dim: int = embedding_dim
        self.num_embeddings: int = num_embeddings
        assert self.spatial_dims in [2, 3], ValueError(
            f"EMAQuantizer only supports 4D and 5D tensor inputs but received spatial dims {spatial_dims}.")
        self.embedding: torch.nn.Embedding = torch.nn.Embedding(self.num_embeddings, self.embedding_dim)
        if embedding_init == "kaiming_uniform":
            torch.nn.init.kaiming_uniform_(self.embedding.weight.data, mode="fan_in", nonlinearity="linear")
        self.embedding.weight.requires_grad = False
        self.commitment_cost: float = commitment_cost
        self.register_buffer("ema_cluster_size", torch.zeros(self.num_embeddings))
        self.register_buffer("ema_w", self.embedding.weight.data.clone())
        self.decay: float = decay
        self.epsilon: float = epsilon
        self.ddp_sync: bool = ddp_sync
        self.flatten_permutation: Sequence[int] = [0] + list(range(2, self.spatial_dims + 2)) + [1]
        self.quantization_permutation: Sequence[int] = [0, self.spatial_dims + 1] + list(
            range(1, self.spatial_dims + 1))

    # ---------------------------------------------------------------- CUDA path (eval)
    def quantize_cl(self, z: torch.Tensor, want_f32: bool = True, ste: bool = True) -> dict:
        """``z``: fp32 channels-last ``[N, D, H, W, pitch]`` encoder output.  Returns indices ``[N, D, H, W]``,
        the gathered rows as a h16 :class:`CL` (decoder input) and fp32 channels-last, the commitment loss and the
        code histogram."""
        lib = _lib.require_device()
        N, D, H, W, P = z.shape
        M, Dm, K = N * D * H * W, self.embedding_dim, self.num_embeddings
        cb = self.embedding.weight.detach()
        cb = cb if cb.dtype == torch.float32 and cb.is_contiguous() else cb.float().contiguous()
        idx = torch.empty((N, D, H, W), dtype=torch.int64, device=z.device)
        q = ops.new_cl(N, (D, H, W), Dm, z.device, self.spatial_dims)
        q32 = torch.empty((N, D, H, W, Dm), dtype=torch.float32, device=z.device) if want_f32 else None
        sq = torch.zeros((), dtype=torch.float64, device=z.device)
        hist = torch.zeros((K,), dtype=torch.int32, device=z.device)
        _lib.check(lib.b200_vq_argmin_gather(z.data_ptr(), M, Dm, P, cb.data_ptr(), K, idx.data_ptr(), q.t.data_ptr(),
                                             q.pitch, None if q32 is None else q32.data_ptr(), 1 if ste else 0,
                                             sq.data_ptr(), hist.data_ptr(), ops._stream()),
                   "b200_vq_argmin_gather")
        loss = (self.commitment_cost * sq / float(M * Dm)).float()
        return dict(indices=idx, q=q, q_f32=q32, loss=loss, hist=hist, count=M)

    def _z_channels_last(self, inputs: torch.Tensor) -> torch.Tensor:
        """NC[D]HW -> fp32 [N, D, H, W, C] (the reference's permute().contiguous(), vector_quantizer.py:105)."""
        x = inputs.float().permute(self.flatten_permutation).contiguous()
        return x.unsqueeze(1) if self.spatial_dims == 2 else x

    def _indices_view(self, idx: torch.Tensor) -> torch.Tensor:
        return idx.squeeze(1) if self.spatial_dims == 2 else idx

    def _to_channel_first(self, q32: torch.Tensor) -> torch.Tensor:
        return ops.from_cl_f32(q32.contiguous(), self.embedding_dim, self.spatial_dims)

    # ---------------------------------------------------------------- reference interface
    def quantize(self, inputs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """vector_quantizer.py:86-122 -> (flat_input, one-hot encodings, indices)."""
        if self.training or not inputs.is_cuda:
            return self._quantize_composite(inputs)
        z = self._z_channels_last(inputs)
        r = self.quantize_cl(z, want_f32=False)
        idx = self._indices_view(r["indices"])
        flat = z.view(-1, self.embedding_dim)
        return flat, torch.nn.functional.one_hot(idx.reshape(-1), self.num_embeddings).float(), idx

    def _quantize_composite(self, inputs: torch.Tensor):
        if not self.training:
            raise RuntimeError("EMAQuantizer: eval-mode quantisation runs on the CUDA kernel only (no CPU path)")
        view = list(inputs.shape)
        del view[1]
        inputs = inputs.float()
        flat = inputs.permute(self.flatten_permutation).contiguous().view(-1, self.embedding_dim)
        w = self.embedding.weight
        distances = (flat ** 2).sum(dim=1, keepdim=True) + (w.t() ** 2).sum(dim=0, keepdim=True) - 2 * torch.mm(flat, w.t())
        idx = torch.max(-distances, dim=1)[1]
        enc = torch.nn.functional.one_hot(idx, self.num_embeddings).float()
        return flat, enc, idx.view(view)

    def embed(self, embedding_indices: torch.Tensor) -> torch.Tensor:
        """vector_quantizer.py:124-138 -> NC[D]HW fp32 rows of the codebook."""
        return self.embedding(embedding_indices).permute(self.quantization_permutation).contiguous()

    @torch.jit.unused
    def distributed_synchronization(self, encodings_sum: torch.Tensor, dw: torch.Tensor) -> None:
        if self.ddp_sync and torch.distributed.is_initialized():
            torch.distributed.all_reduce(tensor=encodings_sum, op=torch.distributed.ReduceOp.SUM)
            torch.distributed.all_reduce(tensor=dw, op=torch.distributed.ReduceOp.SUM)

    def forward(self, inputs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        if not self.training:
            with torch.no_grad():
                r = self.quantize_cl(self._z_channels_last(inputs), want_f32=True, ste=True)
                self._last_hist = (r["hist"], r["count"])
                return self._to_channel_first(r["q_f32"]), r["loss"], self._indices_view(r["indices"])
        # training: the reference's composite (161-188)
        flat_input, encodings, encoding_indices = self._quantize_composite(inputs)
        quantized = self.embed(encoding_indices)
        with torch.no_grad():
            encodings_sum = encodings.sum(0)
            dw = torch.mm(encodings.t(), flat_input)
            if self.ddp_sync:
                self.distributed_synchronization(encodings_sum, dw)
            self.ema_cluster_size.data.mul_(self.decay).add_(torch.mul(encodings_sum, 1 - self.decay))
            n = self.ema_cluster_size.sum()
            weights = (self.ema_cluster_size + self.epsilon) / (n + self.num_embeddings * self.epsilon) * n
            self.ema_w.data.mul_(self.decay).add_(torch.mul(dw, 1 - self.decay))
            self.embedding.weight.data.copy_(self.ema_w / weights.unsqueeze(1))
        loss = self.commitment_cost * torch.nn.functional.mse_loss(quantized.detach(), inputs)
        quantized = inputs + (quantized - inputs).detach()
        self._last_hist = None
        return quantized, loss, encoding_indices


class VectorQuantizer(torch.nn.Module):
    """vector_quantizer.py:191-228 (keeps the ``perplexity`` side effect)."""

    def __init__(self, quantizer: torch.nn.Module = None):
        super().__init__()
        self.quantizer: torch.nn.Module = quantizer
        self.perplexity: torch.Tensor = torch.rand(1)

    def _perplexity_from_hist(self, hist: torch.Tensor, count: int) -> torch.Tensor:
        avg_probs = hist.float().div(count)
        return torch.exp(-torch.sum(avg_probs * torch.log(avg_probs + 1e-10)))

    def forward(self, inputs: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        quantized, loss, encoding_indices = self.quantizer(inputs)
        last = getattr(self.quantizer, "_last_hist", None)
        if last is not None:
            self.perplexity = self._perplexity_from_hist(*last)
        else:
            avg_probs = (torch.histc(encoding_indices.float(), bins=self.quantizer.num_embeddings,
                                     max=self.quantizer.num_embeddings).float().div(encoding_indices.numel()))
            self.perplexity = torch.exp(-torch.sum(avg_probs * torch.log(avg_probs + 1e-10)))
        return loss, quantized

    def forward_cl(self, z: torch.Tensor, want_f32: bool) -> dict:
        """Channels-last fast path used by VQVAE (fp32 encoder output in, h16 decoder input out)."""
        r = self.quantizer.quantize_cl(z, want_f32=want_f32)
        self.perplexity = self._perplexity_from_hist(r["hist"], r["count"])
        return r

    def embed(self, embedding_indices: torch.Tensor) -> torch.Tensor:
        return self.quantizer.embed(embedding_indices=embedding_indices)

    def quantize(self, encodings: torch.Tensor) -> torch.Tensor:
        _, _, encoding_indices = self.quantizer(encodings)
        return encoding_indices
