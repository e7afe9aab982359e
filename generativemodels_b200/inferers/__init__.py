from .inferer import (ControlNetDiffusionInferer, ControlNetLatentDiffusionInferer, DiffusionInferer,  # noqa: F401
                      LatentDiffusionInferer, VQVAETransformerInferer)
