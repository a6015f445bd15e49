"""``from tensorlink_b200.ml import DistributedModel`` — same import shape as ``tensorlink.ml``
(/root/reference/tensorlink/ml/__init__.py:1)."""
from .module import CausalLMOutput, DistributedModel  # noqa: F401
