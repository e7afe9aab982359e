from enum import Enum
import importlib


class StrEnum(str, Enum):
    def __str__(self):
        return self.value

    def __repr__(self):
        return self.value


def ensure_tuple_rep(tup, dim):
    if isinstance(tup, (list, tuple)):
        if len(tup) == dim:
            return tuple(tup)
        raise ValueError(f"Sequence must have length {dim}, got {len(tup)}.")
    return (tup,) * dim


def min_version(*_a, **_k):
    return True


def optional_import(module, version="", version_checker=min_version, name="", descriptor="", version_args=None,
                    allow_namespace_pkg=False, as_type="default"):
    try:
        mod = importlib.import_module(module)
        obj = getattr(mod, name) if name else mod
        return obj, True
    except Exception:
        if as_type == "base":
            class _Missing:  # dummy base class
                pass
            return _Missing, False

        class _Lazy:
            def __getattr__(self, item):
                raise ImportError(f"optional module {module} is not available")

            def __call__(self, *a, **k):
                raise ImportError(f"optional module {module} is not available")
        return _Lazy(), False


class LossReduction(StrEnum):
    NONE = "none"
    MEAN = "mean"
    SUM = "sum"


class MetricReduction(StrEnum):
    NONE = "none"
    MEAN = "mean"
    SUM = "sum"
    MEAN_BATCH = "mean_batch"
    SUM_BATCH = "sum_batch"
    MEAN_CHANNEL = "mean_channel"
    SUM_CHANNEL = "sum_channel"


def convert_data_type(data, output_type=None, device=None, dtype=None, wrap_sequence=False, safe=False):
    return data, type(data), getattr(data, "device", None)


def convert_to_dst_type(src, dst, dtype=None, wrap_sequence=False, device=None, safe=False):
    return src, type(src), getattr(src, "device", None)
