"""monai.networks.blocks.Convolution / ADN semantics: `conv` child = nn.Conv / nn.ConvTranspose with
same_padding(k, dilation) when padding is None and output_padding = stride - 1 when None for transposed; unless
conv_only, an `adn` child adds only the letters present in `ordering` (A = activation, D = dropout, N = norm)."""
import numpy as np
import torch.nn as nn

from ..layers.factories import Conv, get_act_layer, get_norm_layer


def same_padding(kernel_size, dilation=1):
    k = np.atleast_1d(kernel_size)
    d = np.atleast_1d(dilation)
    p = (k - 1) // 2 * d
    p = tuple(int(v) for v in p)
    return p if len(p) > 1 else p[0]


def stride_minus_kernel_padding(kernel_size, stride):
    k = np.atleast_1d(kernel_size)
    s = np.atleast_1d(stride)
    p = tuple(int(v) for v in (s - 1))
    return p if len(p) > 1 else p[0]


class ADN(nn.Sequential):
    def __init__(self, ordering="NDA", in_channels=None, act="RELU", norm=None, norm_dim=None, dropout=None,
                 dropout_dim=1):
        super().__init__()
        ops = {"A": None, "D": None, "N": None}
        if norm is not None:
            if norm_dim is None:
                raise ValueError("norm_dim needs to be specified.")
            ops["N"] = get_norm_layer(name=norm, spatial_dims=norm_dim, channels=in_channels)
        if act is not None:
            ops["A"] = get_act_layer(act)
        if dropout is not None:
            p = dropout if isinstance(dropout, (int, float)) else dropout[1].get("p", 0.5)
            ops["D"] = {1: nn.Dropout, 2: nn.Dropout2d, 3: nn.Dropout3d}[dropout_dim](p)
        for item in ordering.upper():
            if ops.get(item) is not None:
                self.add_module(item, ops[item])


class Convolution(nn.Sequential):
    def __init__(self, spatial_dims, in_channels, out_channels, strides=1, kernel_size=3, adn_ordering="NDA",
                 act="PRELU", norm="INSTANCE", dropout=None, dropout_dim=1, dilation=1, groups=1, bias=True,
                 conv_only=False, is_transposed=False, padding=None, output_padding=None):
        super().__init__()
        self.spatial_dims = spatial_dims
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.is_transposed = is_transposed
        if padding is None:
            padding = same_padding(kernel_size, dilation)
        conv_type = Conv["CONVTRANS" if is_transposed else "CONV", spatial_dims]
        if is_transposed:
            if output_padding is None:
                output_padding = stride_minus_kernel_padding(1, strides)
            conv = conv_type(in_channels, out_channels, kernel_size=kernel_size, stride=strides, padding=padding,
                             output_padding=output_padding, groups=groups, bias=bias, dilation=dilation)
        else:
            conv = conv_type(in_channels, out_channels, kernel_size=kernel_size, stride=strides, padding=padding,
                             dilation=dilation, groups=groups, bias=bias)
        self.add_module("conv", conv)
        if conv_only:
            return
        if act is None and norm is None and dropout is None:
            return
        self.add_module("adn", ADN(ordering=adn_ordering, in_channels=out_channels, act=act,
                                   norm=None if norm in (None, "INSTANCE") and "N" not in adn_ordering.upper() else norm,
                                   norm_dim=spatial_dims, dropout=dropout, dropout_dim=dropout_dim))
