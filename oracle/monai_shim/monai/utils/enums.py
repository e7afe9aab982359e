from . import StrEnum  # noqa: F401
