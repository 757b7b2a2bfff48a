from .hybrid_rescore import HybridRescorer  # noqa: F401
from .rescore_stages import EnhancedHybridSearchV2, KhopReranker  # noqa: F401
