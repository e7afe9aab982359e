from . import convert_to_dst_type, convert_data_type  # noqa: F401
