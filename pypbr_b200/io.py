"""
pypbr_b200.io — loading a material from a folder of images and saving it back (pypbr/io.py:22-230): the data format on
either side of the shading path.

Same functions, arguments, naming conventions, workflow selection and warnings as the reference.  One addition:
`load_material_from_folder(..., device=...)`.  With a CUDA device every 8- / 16-bit image is uploaded as the bytes the
decoder produced (1 - 3 bytes per texel instead of 4 - 12) and expanded on the device by pbr_ingest_image - `/255`,
`/65535`, and for the normal map the `*2-1` + normalise - bit-identical to the host conversion
(pypbr_b200/materials/_ingest.py); without it the maps stay on the host exactly like the reference's.
"""

from __future__ import annotations

import os
import warnings
from typing import Dict, List, Optional, Type

import torch
from PIL import Image

from .materials import BasecolorMetallicMaterial, DiffuseSpecularMaterial, MaterialBase

__all__ = ["load_material_from_folder", "select_material_class", "save_material_to_folder"]

# map type -> accepted file stems, in the order they are looked for (io.py:43-52)
DEFAULT_MAP_NAMES: Dict[str, List[str]] = {
    "basecolor": ["albedo", "basecolor"],
    "diffuse": ["diffuse"],
    "normal": ["normal", "normalmap"],
    "height": ["height", "displacement", "bump"],
    "roughness": ["roughness"],
    "metallic": ["metallic", "metalness"],
    "specular": ["specular"],
}
EXTENSIONS = ("png", "jpg", "jpeg", "tiff", "bmp", "exr")           # io.py:54
_COLOUR_MAPS = ("basecolor", "diffuse", "normal", "specular")       # decoded as RGB (io.py:67-68)
_KEEP_MODES = ("I", "I;16", "I;16B", "I;16L", "I;16N", "F")         # height maps keep 16-bit / float data (io.py:69-75)


def _first_image(folder: str, stems: List[str]) -> Optional[str]:
    """The first existing `<stem>.<ext>`, stems outermost (io.py:58-82)."""
    for stem in stems:
        for ext in EXTENSIONS:
            path = os.path.join(folder, f"{stem}.{ext}")
            if os.path.isfile(path):
                return path
    return None


def _decode(map_type: str, path: str) -> Image.Image:
    image = Image.open(path)
    if map_type in _COLOUR_MAPS:
        return image.convert("RGB")
    if map_type == "height" and image.mode in _KEEP_MODES:
        return image
    return image if image.mode == "L" else image.convert("L")


def select_material_class(loaded_maps: Dict[str, Image.Image], preferred_workflow: Optional[str] = None) -> Type[MaterialBase]:
    """
    The material class for the maps that were found (io.py:131-186).  With both a metallic and a specular map the
    preferred workflow wins (metallic when none is given) and the other map is REMOVED from `loaded_maps`, with a warning.
    """
    has_metallic, has_specular = "metallic" in loaded_maps, "specular" in loaded_maps
    if has_metallic and has_specular:
        if preferred_workflow == "specular":
            warnings.warn("Both metallic and specular maps are present. Using specular workflow as preferred.")
            loaded_maps.pop("metallic", None)
            return DiffuseSpecularMaterial
        if preferred_workflow == "metallic":
            warnings.warn("Both metallic and specular maps are present. Using metallic workflow as preferred.")
        else:
            warnings.warn("Both metallic and specular maps are present. Specify preferred_workflow to choose. "
                          "Defaulting to metallic workflow.")
        loaded_maps.pop("specular", None)
        return BasecolorMetallicMaterial
    if has_metallic:
        return BasecolorMetallicMaterial
    if has_specular:
        return DiffuseSpecularMaterial
    if "basecolor" in loaded_maps:
        return BasecolorMetallicMaterial
    if "diffuse" in loaded_maps:
        return DiffuseSpecularMaterial
    warnings.warn("Neither metallic nor specular map found, and no albedo map found. Defaulting to BasecolorMetallicMaterial.")
    return BasecolorMetallicMaterial


def load_material_from_folder(
    folder_path: str,
    map_names: Optional[Dict[str, List[str]]] = None,
    preferred_workflow: Optional[str] = None,
    is_srgb: bool = True,
    device: Optional[torch.device] = None,
) -> MaterialBase:
    """
    Load material maps from a folder using naming conventions (io.py:22-128).

    Args:
        folder_path: folder with one image per map.
        map_names: map type -> accepted file stems (default: DEFAULT_MAP_NAMES).
        preferred_workflow: 'metallic' or 'specular' when the folder holds both kinds of maps.
        is_srgb: whether the albedo and specular maps are sRGB-encoded.
        device: where the maps are to live; a CUDA device uploads the images' own bytes and converts them on the device.

    Returns:
        BasecolorMetallicMaterial or DiffuseSpecularMaterial.
    """
    found: Dict[str, Image.Image] = {}
    for map_type, stems in (map_names or DEFAULT_MAP_NAMES).items():
        path = _first_image(folder_path, stems)
        if path is not None:
            found[map_type] = _decode(map_type, path)

    cls = select_material_class(found, preferred_workflow)
    if issubclass(cls, BasecolorMetallicMaterial):
        albedo_key, other_key = "basecolor", "diffuse"
        missing = "Basecolor map not found for metallic workflow. Looking for 'albedo' or 'basecolor' maps."
    else:
        albedo_key, other_key = "diffuse", "basecolor"
        missing = "Diffuse map not found for specular workflow. Looking for 'diffuse' map."
    albedo = found.get(albedo_key, None)
    if albedo is None:
        warnings.warn(missing)
    maps = {k: v for k, v in found.items() if k not in (albedo_key, other_key)}
    extra = {} if device is None else {"device": torch.device(device)}
    return cls(albedo=albedo, albedo_is_srgb=is_srgb, specular_is_srgb=is_srgb, **maps, **extra)


def save_material_to_folder(material: MaterialBase, folder_path: str, map_names: Optional[Dict[str, str]] = None,
                            format: str = "png"):
    """
    Save every map of `material` as `<name>.<format>` (io.py:189-230); `map_names` renames maps (map -> file stem).
    Maps come back from the device once, as the 8-bit (16-bit for 'I' modes) images `MaterialBase.to_pil()` produces.
    """
    os.makedirs(folder_path, exist_ok=True)
    names = map_names or {}
    for map_type, image in material.to_pil().items():
        if image is None:
            continue
        stem = map_type.lstrip("_")
        image.save(os.path.join(folder_path, f"{names.get(stem, stem)}.{format}"))
