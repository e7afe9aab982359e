class IgniteInfo:
    OPT_IMPORT_VERSION = "0.4.4"
