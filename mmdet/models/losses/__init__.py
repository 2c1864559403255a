from .oadg import ContrastiveLossPlus, supcontrast  # noqa: F401
