from .component_store import ComponentStore  # noqa: F401
from .misc import unsqueeze_left, unsqueeze_right  # noqa: F401
from .ordering import Ordering  # noqa: F401
