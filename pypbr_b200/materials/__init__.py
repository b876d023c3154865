"""pypbr_b200.materials — mirrors pypbr/materials/__init__.py."""

from .base import MaterialBase
from .diffuse import DiffuseSpecularMaterial
from .metallic import BasecolorMetallicMaterial

__all__ = ["MaterialBase", "BasecolorMetallicMaterial", "DiffuseSpecularMaterial"]
