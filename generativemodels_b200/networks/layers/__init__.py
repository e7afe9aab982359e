from .vector_quantizer import EMAQuantizer, VectorQuantizer  # noqa: F401
