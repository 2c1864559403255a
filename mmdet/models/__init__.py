from .builder import LOSSES, build_loss  # noqa: F401
