from .factories import Act, Conv, Pool, get_act_layer, get_pool_layer  # noqa: F401
