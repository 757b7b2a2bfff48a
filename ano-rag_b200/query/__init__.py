from .hybrid_rescore import HybridRescorer  # noqa: F401
