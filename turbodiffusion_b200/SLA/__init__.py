# Mirror of turbodiffusion/SLA/__init__.py:16-24
from .core import SparseLinearAttention, SageSparseLinearAttention

__all__ = ["SparseLinearAttention", "SageSparseLinearAttention"]
