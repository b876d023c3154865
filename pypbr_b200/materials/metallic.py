"""pypbr_b200.materials.metallic — mirrors pypbr/materials/metallic.py."""

from __future__ import annotations

from .base import MaterialBase


class BasecolorMetallicMaterial(MaterialBase):
    """
    PBR material in the basecolor-metallic workflow.

    Attributes:
        albedo, normal, roughness, metallic (torch.Tensor)
    """

    def __init__(self, albedo=None, albedo_is_srgb: bool = True, normal=None, roughness=None, metallic=None, **kwargs):
        super().__init__(albedo=albedo, albedo_is_srgb=albedo_is_srgb, normal=normal, roughness=roughness, **kwargs)
        if metallic is not None:
            self.metallic = metallic

    @property
    def basecolor(self):
        return self.albedo

    @basecolor.setter
    def basecolor(self, value):
        self.albedo = value

    def to_diffuse_specular_material(self, albedo_is_srgb: bool = False):
        """
        Convert to the diffuse-specular workflow (pypbr/materials/metallic.py:71-120) in one fused
        streaming kernel: sRGB decode, diffuse = a(1-m), specular = 0.04(1-m) + a*m.

        Reference behaviour kept: normal and roughness are passed on (the constructor re-runs normal
        ingestion, which aliases the tensor when it already has negative components); other maps
        (height, ...) are dropped; `specular_is_srgb` of the result stays at its default True.
        """
        from ._convert import convert
        from .diffuse import DiffuseSpecularMaterial

        if self.albedo is None or self.metallic is None:
            raise ValueError("Both albedo and metallic maps are required for conversion.")
        diffuse, specular = convert(self.albedo, self.metallic, self.albedo_is_srgb, m2s=True)
        return DiffuseSpecularMaterial(
            albedo=diffuse,
            specular=specular,
            normal=self.normal,
            roughness=self.roughness,
            albedo_is_srgb=albedo_is_srgb,
            device=diffuse.device,
        )
