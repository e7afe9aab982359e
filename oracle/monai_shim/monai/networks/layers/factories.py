import torch.nn as nn


class _Factory:
    def __init__(self, table):
        self._table = table

    def __getitem__(self, key):
        if isinstance(key, tuple):
            name, dim = key
            return self._table[str(name).upper()][dim]
        return self._table[str(key).upper()]

    def __getattr__(self, item):
        if item.startswith("_"):
            raise AttributeError(item)
        return item.upper()


Conv = _Factory({"CONV": {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d},
                 "CONVTRANS": {1: nn.ConvTranspose1d, 2: nn.ConvTranspose2d, 3: nn.ConvTranspose3d}})
Pool = _Factory({"AVG": {1: nn.AvgPool1d, 2: nn.AvgPool2d, 3: nn.AvgPool3d},
                 "MAX": {1: nn.MaxPool1d, 2: nn.MaxPool2d, 3: nn.MaxPool3d}})
Act = _Factory({"RELU": nn.ReLU, "LEAKYRELU": nn.LeakyReLU, "SILU": nn.SiLU, "SWISH": nn.SiLU, "GELU": nn.GELU,
                "TANH": nn.Tanh, "SIGMOID": nn.Sigmoid, "PRELU": nn.PReLU, "ELU": nn.ELU})


def get_act_layer(name):
    if name == "" or name is None:
        return nn.Identity()
    if isinstance(name, str):
        return Act[name]()
    n, kwargs = name
    return Act[n](**kwargs)


def get_pool_layer(name, spatial_dims=1):
    if isinstance(name, str):
        return Pool[name, spatial_dims]()
    n, kwargs = name
    return Pool[n, spatial_dims](**kwargs)


def get_norm_layer(name, spatial_dims=1, channels=1):
    """monai.networks.layers.utils.get_norm_layer for the norms the SPADE blocks use: "INSTANCE" -> nn.InstanceNorm{d}d
    (PyTorch defaults: no affine, no running stats), ("GROUP", {...}) -> nn.GroupNorm(num_channels=channels, ...),
    "BATCH" -> nn.BatchNorm{d}d."""
    if name == "" or name is None:
        return nn.Identity()
    if isinstance(name, str):
        n, kwargs = name, {}
    else:
        n, kwargs = name
    n = str(n).upper()
    if n == "INSTANCE":
        return {1: nn.InstanceNorm1d, 2: nn.InstanceNorm2d, 3: nn.InstanceNorm3d}[spatial_dims](channels, **kwargs)
    if n == "GROUP":
        return nn.GroupNorm(num_channels=channels, **kwargs)
    if n == "BATCH":
        return {1: nn.BatchNorm1d, 2: nn.BatchNorm2d, 3: nn.BatchNorm3d}[spatial_dims](channels, **kwargs)
    raise NotImplementedError(f"shim: norm {n}")
