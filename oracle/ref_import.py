"""Import the UNMODIFIED reference (/root/reference/generative) as `generative`, on top of oracle/monai_shim when
real MONAI is absent.  Only usable in the build container (the GPU box has no /root/reference): used to validate
oracle/torch_oracle.py and by tests/golden/make_golden.py to generate the committed fixtures."""
import sys
from pathlib import Path

REF_ROOT = Path("/root/reference")
_SHIM = Path(__file__).resolve().parent / "monai_shim"


def available() -> bool:
    return (REF_ROOT / "generative" / "__init__.py").exists()


def import_reference():
    if not available():
        raise ImportError("/root/reference is not present (expected on the GPU box); use the committed golden vectors")
    try:
        import monai  # noqa: F401
    except Exception:
        if str(_SHIM) not in sys.path:
            sys.path.insert(0, str(_SHIM))
    if str(REF_ROOT) not in sys.path:
        sys.path.insert(0, str(REF_ROOT))
    import generative  # noqa: F401
    from generative import inferers, networks  # noqa: F401
    from generative.networks import nets, schedulers  # noqa: F401
    return generative
