from .ddim import DDIMScheduler  # noqa: F401
from .ddpm import DDPMScheduler  # noqa: F401
from .pndm import PNDMScheduler  # noqa: F401
from .scheduler import NoiseSchedules, Scheduler  # noqa: F401
