from .factories import get_act_layer, get_pool_layer  # noqa: F401
