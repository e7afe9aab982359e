from .convolutions import ADN, Convolution  # noqa: F401
from .mlp import MLPBlock  # noqa: F401
