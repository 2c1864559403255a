from oadg_b200.registry import Compose  # noqa: F401
from .oa_mix import OAMix  # noqa: F401
