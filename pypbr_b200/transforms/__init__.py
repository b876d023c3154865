"""pypbr_b200.transforms — class-based and functional material transforms (pypbr/transforms/__init__.py:11-35)."""

from . import functional
from .transforms import *  # noqa: F401,F403
from .transforms import __all__ as _classes

__all__ = ["functional", *_classes]
