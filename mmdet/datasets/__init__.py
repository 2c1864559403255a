from .builder import PIPELINES  # noqa: F401
