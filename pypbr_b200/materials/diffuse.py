"""pypbr_b200.materials.diffuse — mirrors pypbr/materials/diffuse.py."""

from __future__ import annotations

from ..utils import linear_to_srgb, srgb_to_linear
from .base import MaterialBase


class DiffuseSpecularMaterial(MaterialBase):
    """
    PBR material in the diffuse-specular workflow.

    Attributes:
        albedo, normal, roughness, specular (torch.Tensor)
    """

    def __init__(
        self,
        albedo=None,
        albedo_is_srgb: bool = True,
        normal=None,
        roughness=None,
        specular=None,
        specular_is_srgb: bool = True,
        **kwargs,
    ):
        super().__init__(albedo=albedo, albedo_is_srgb=albedo_is_srgb, normal=normal, roughness=roughness, **kwargs)
        self.specular_is_srgb = specular_is_srgb
        if specular is not None:
            self.specular = specular

    @property
    def diffuse(self):
        return self.albedo

    @diffuse.setter
    def diffuse(self, value):
        self.albedo = value

    @property
    def linear_specular(self):
        """Specular map in linear space (pypbr/materials/diffuse.py:76-91)."""
        specular = self._maps.get("specular", None)
        if specular is None:
            return None
        return srgb_to_linear(specular) if self.specular_is_srgb else specular

    def to_basecolor_metallic_material(self, albedo_is_srgb: bool = False):
        """
        Convert to the basecolor-metallic workflow (pypbr/materials/diffuse.py:93-158) in one fused
        streaming kernel.  Reference behaviour kept: the RAW specular map is used (not linear_specular),
        the metallic map comes out with 3 channels, the result is not the inverse of the forward
        conversion.
        """
        from ._convert import convert
        from .metallic import BasecolorMetallicMaterial

        if self.albedo is None or self.specular is None:
            raise ValueError("Both albedo (diffuse) and specular maps are required for conversion.")
        basecolor, metallic = convert(self.albedo, self.specular, self.albedo_is_srgb, m2s=False)
        return BasecolorMetallicMaterial(
            albedo=basecolor,
            metallic=metallic,
            normal=self.normal,
            roughness=self.roughness,
            albedo_is_srgb=albedo_is_srgb,
            device=basecolor.device,
        )

    def to_linear(self):
        super().to_linear()
        specular = self._maps.get("specular", None)
        if specular is not None and self.specular_is_srgb:
            self._maps["specular"] = srgb_to_linear(specular)
            self.specular_is_srgb = False
        return self

    def to_srgb(self):
        super().to_srgb()
        specular = self._maps.get("specular", None)
        if specular is not None and not self.specular_is_srgb:
            self._maps["specular"] = linear_to_srgb(specular)
            self.specular_is_srgb = True
        return self
