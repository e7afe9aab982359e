"""monai MLPBlock with act="GEGLU": linear1 = Linear(hidden, 2*mlp_dim), fn(x) = a * gelu(gate), linear2."""
import torch.nn as nn
import torch.nn.functional as F

from ..layers.factories import get_act_layer


class GEGLU(nn.Module):
    def forward(self, x):
        x, gate = x.chunk(2, dim=-1)
        return x * F.gelu(gate)


class MLPBlock(nn.Module):
    def __init__(self, hidden_size, mlp_dim, dropout_rate=0.0, act="GELU", dropout_mode="vit"):
        super().__init__()
        mlp_dim = mlp_dim or hidden_size
        geglu = isinstance(act, str) and act.upper() == "GEGLU"
        self.linear1 = nn.Linear(hidden_size, mlp_dim * 2) if geglu else nn.Linear(hidden_size, mlp_dim)
        self.linear2 = nn.Linear(mlp_dim, hidden_size)
        self.fn = GEGLU() if geglu else get_act_layer(act)
        self.drop1 = nn.Dropout(dropout_rate)
        self.drop2 = nn.Dropout(dropout_rate)

    def forward(self, x):
        x = self.fn(self.linear1(x))
        x = self.drop1(x)
        x = self.linear2(x)
        return self.drop2(x)
