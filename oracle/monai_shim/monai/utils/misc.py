from . import ensure_tuple_rep  # noqa: F401
