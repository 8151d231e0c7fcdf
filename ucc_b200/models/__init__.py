"""Model families used by the examples / DDP benchmark (random-init, synthetic data: there is no dataset here)."""
from .resnet import resnet50  # noqa: F401
from .mlp import MLP, TPTransformerBlock  # noqa: F401
