from .spade_norm import SPADE, SegPyramid  # noqa: F401
