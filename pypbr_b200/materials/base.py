"""
pypbr_b200.materials.base — the material container the shading kernels are fed from.

Host-side mirror of pypbr/materials/base.py (the attribute protocol of SURVEY.md §8a row a19): a dict
of named texture maps reached through attribute access, plus the index transforms and colour-space
switches that operate map by map.  Same names, arguments and error behaviour as the reference, with
two deliberate, documented differences:

  * CUDA float32 tensors are accepted as maps.  The reference gates on
    ``isinstance(value, torch.FloatTensor)`` (base.py:96-101), which is true for CPU tensors only, so
    a CUDA tensor silently becomes a plain attribute there and the BRDF crashes.
  * maps may carry a leading batch dimension, (B, C, H, W): a batch is B independent materials.

The container itself does no per-texel arithmetic except normal-map ingestion, which runs in the
kernels for CUDA tensors (pbr_normal_min / pbr_normal_ingest) and, for CPU tensors created while a
material is being assembled on the host, with the same torch ops as the reference.
"""

from __future__ import annotations

import copy
import math
import weakref
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F

from .. import _cabi
from ..utils import NormalConvention, compute_height_from_normal, compute_normal_from_height, linear_to_srgb, rotate_normals, srgb_to_linear

try:  # PIL / torchvision are only needed for image ingestion and resize/crop
    from PIL import Image
except Exception:  # pragma: no cover
    Image = None


def _is_map_value(value) -> bool:
    if value is None:
        return True
    if Image is not None and isinstance(value, Image.Image):
        return True
    if isinstance(value, np.ndarray):
        return True
    return isinstance(value, torch.Tensor) and value.dtype == torch.float32


def _channel_dim(t: torch.Tensor) -> int:
    return t.dim() - 3


class MaterialBase:
    """
    Base class for PBR materials: a registry of texture maps with attribute access.

    Attributes:
        albedo, normal, roughness (torch.Tensor): maps of shape (C, H, W) or (B, C, H, W).
    """

    _PLAIN = ("albedo_is_srgb", "_maps")

    def __init__(
        self,
        albedo=None,
        albedo_is_srgb: bool = True,
        normal=None,
        roughness=None,
        normal_convention: NormalConvention = NormalConvention.OPENGL,
        device: torch.device = torch.device("cpu"),
        **kwargs,
    ):
        self.device = device
        self.normal_convention = normal_convention
        self._maps = {}
        self.albedo_is_srgb = albedo_is_srgb
        for name, value in (("albedo", albedo), ("normal", normal), ("roughness", roughness)):
            if value is not None:
                setattr(self, name, value)
        for name, value in kwargs.items():
            setattr(self, name, value)

    # ------------------------------------------------------------------ attribute protocol
    def __setattr__(self, name, value):
        # base.py:86-103: images / arrays / float tensors / None are texture maps, the rest are attributes
        if name in self._PLAIN or not _is_map_value(value):
            object.__setattr__(self, name, value)
        else:
            self._maps[name] = self._process_map(name, value)

    def __getattr__(self, name):
        # only reached when normal lookup fails (base.py:105-120)
        maps = self.__dict__.get("_maps")
        if maps is not None and name in maps:
            return maps[name]
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'")

    # ------------------------------------------------------------------ ingestion
    def _to_tensor(self, image) -> Optional[torch.Tensor]:
        """base.py:122-168."""
        if image is None:
            return None
        if isinstance(image, torch.Tensor):
            return image.to(self.device)
        if isinstance(image, np.ndarray):
            return torch.from_numpy(image).float().to(self.device)
        if Image is not None and isinstance(image, Image.Image):
            if image.mode in ("I", "I;16", "I;16B", "I;16L", "I;16N"):
                arr = np.array(image, dtype=np.uint16).astype(np.float32)
                return (torch.from_numpy(arr).unsqueeze(0) / 65535.0).to(self.device)
            if image.mode == "F":
                return torch.from_numpy(np.array(image, dtype=np.float32)).unsqueeze(0).to(self.device)
            from torchvision.transforms import functional as TF

            if image.mode == "RGBA":
                image = image.convert("RGB")
            return TF.to_tensor(image).to(self.device)
        raise TypeError(
            f"Unsupported image type: {type(image)}. Supported types are PIL.Image.Image, np.ndarray, and torch.FloatTensor."
        )

    def _process_map(self, name, value):
        if value is None:
            return None
        if Image is not None and isinstance(value, Image.Image) and torch.device(self.device).type == "cuda":
            # 8/16-bit image straight to the device: /255 (/65535) and the normal remap fused in one kernel
            from ._ingest import ingest_pil

            fused = ingest_pil(value, self.device, name == "normal")
            if fused is not None:
                return fused
        tensor = self._to_tensor(value)
        return self._process_normal_map(tensor) if name == "normal" else tensor

    def _process_normal_map(self, normal_map: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        """
        base.py:191-221.  2 channels: reconstruct z.  3 channels: if ANY component is negative the map
        is taken as already in [-1, 1] and returned untouched (aliased); otherwise it is read as an
        RGB-encoded map, remapped with *2-1 and normalised.  The probe is a full min() reduction.
        """
        if normal_map is None:
            return None
        ch = normal_map.shape[_channel_dim(normal_map)]
        if ch == 2:
            return self._compute_normal_map_z_component(normal_map)
        if ch != 3:
            raise ValueError("Normal map must have 2 or 3 channels.")
        if normal_map.is_cuda:
            if _normal_min_cuda(normal_map) < 0:
                return normal_map
            return normal_ingest(normal_map, 3)
        if normal_map.min() < 0:
            return normal_map
        return F.normalize(normal_map * 2.0 - 1.0, dim=_channel_dim(normal_map))

    def _compute_normal_map_z_component(self, normal_xy: torch.Tensor) -> torch.Tensor:
        """base.py:223-242: xy in [0,1] -> [-1,1], z = sqrt(clamp(1 - x^2 - y^2, 1e-6)), normalise."""
        if normal_xy.is_cuda:
            return normal_ingest(normal_xy, 2)
        cd = _channel_dim(normal_xy)
        xy = normal_xy * 2 - 1
        x, y = xy.narrow(cd, 0, 1), xy.narrow(cd, 1, 1)
        z = torch.sqrt(torch.clamp(1.0 - (x**2 + y**2), min=1e-6))
        return F.normalize(torch.cat([x, y, z], dim=cd), dim=cd)

    # ------------------------------------------------------------------ device
    def to(self, device: torch.device):
        """Move every map to `device` (base.py:245-259). Returns self."""
        self.device = device
        for name, t in self._maps.items():
            if t is not None:
                self._maps[name] = t.to(device)
        return self

    # ------------------------------------------------------------------ properties
    @property
    def linear_albedo(self):
        """Albedo in linear space (base.py:262-277); None when there is no albedo map."""
        albedo = self._maps.get("albedo", None)
        if albedo is None:
            return None
        return srgb_to_linear(albedo) if self.albedo_is_srgb else albedo

    @property
    def normal_rgb(self):
        normal = self._maps.get("normal", None)
        return None if normal is None else (normal + 1.0) * 0.5

    @property
    def size(self) -> Optional[Tuple[int, int]]:
        """(height, width) of the first non-None map, else None (base.py:293-307)."""
        for t in self._maps.values():
            if t is not None:
                return (t.shape[-2], t.shape[-1])
        return None

    # ------------------------------------------------------------------ packing
    def as_dict(self) -> Dict[str, torch.Tensor]:
        return dict(self._maps)

    def as_tensor(self, names: Optional[List[Union[str, Tuple[str, int]]]] = None, normalize: Optional[bool] = False):
        """Stack (a subset of) the maps along the channel dimension (base.py:319-414)."""
        if not names:
            wanted = [(n, None) for n in self._maps]
        else:
            if not isinstance(names, list):
                raise TypeError("names must be a list of strings or tuples.")
            wanted = []
            for item in names:
                if isinstance(item, str):
                    wanted.append((item, None))
                elif isinstance(item, tuple):
                    if len(item) != 2:
                        raise ValueError("Each tuple in names must have exactly two elements: (map_name, channel_limit).")
                    nm, lim = item
                    if not isinstance(nm, str):
                        raise TypeError("The first element of each tuple must be a string (map name).")
                    if not isinstance(lim, int) or lim <= 0:
                        raise ValueError("The second element of each tuple must be a positive integer (channel limit).")
                    wanted.append((nm, lim))
                else:
                    raise TypeError("Each item in names must be either a string or a tuple of (str, int).")
        parts = []
        for nm, lim in wanted:
            if nm not in self._maps:
                raise KeyError(f"Map '{nm}' does not exist in the texture maps.")
            t = self._maps[nm]
            if not isinstance(t, torch.Tensor):
                raise TypeError(f"Map '{nm}' is not a torch.Tensor.")
            cd = _channel_dim(t)
            if lim is not None:
                if lim > t.size(cd):
                    raise ValueError(
                        f"Requested {lim} channels for map '{nm}', but only {t.size(cd)} channels are available."
                    )
                t = t.narrow(cd, 0, lim)
            if normalize and nm != "normal":
                t = (t - 0.5) / 0.5
            parts.append(t)
        if not parts:
            raise ValueError("No valid texture maps found to stack.")
        if any(p.shape[-2:] != parts[0].shape[-2:] for p in parts):
            raise ValueError("All texture maps must have the same spatial dimensions for concatenation.")
        return torch.cat(parts, dim=_channel_dim(parts[0]))

    @classmethod
    def from_tensor(
        cls,
        tensor: torch.Tensor,
        names: Optional[List[Union[str, Tuple[str, int]]]] = None,
        normal_convention: NormalConvention = NormalConvention.OPENGL,
        is_normalized: bool = False,
        device: torch.device = torch.device("cpu"),
    ) -> "MaterialBase":
        """Unpack a channel-stacked tensor into a material (base.py:416-487)."""
        inst = cls(normal_convention=normal_convention, device=device)
        if not names:
            names = [(n, inst._maps[n].size(0)) for n in inst._maps]
        layout = []
        for item in names:
            if isinstance(item, str):
                if item in inst._maps and isinstance(inst._maps[item], torch.Tensor):
                    layout.append((item, inst._maps[item].size(0)))
                else:
                    raise KeyError(f"Cannot infer channel count for map '{item}'. Provide a tuple instead.")
            elif isinstance(item, tuple):
                if len(item) != 2:
                    raise ValueError("Each tuple must be (map_name, channel_limit).")
                layout.append(item)
            else:
                raise TypeError("Configuration items must be a string or tuple (str, int).")
        cd = _channel_dim(tensor)
        total = sum(n for _, n in layout)
        if tensor.size(cd) != total:
            raise ValueError(
                f"Packed tensor has {tensor.size(cd)} channels, but configuration expects {total} channels."
            )
        at = 0
        for nm, n in layout:
            part = tensor.narrow(cd, at, n).clone()
            if is_normalized:
                part = part * 0.5 + 0.5
            if nm == "normal" and n == 2:
                part = inst._compute_normal_map_z_component(part)
            inst._maps[nm] = part
            at += n
        return inst

    # ------------------------------------------------------------------ index transforms
    # Pure data movement: they call the very torch / torchvision ops the reference calls, so results
    # are bit-identical to the reference on the same device by construction.
    def _each(self, fn):
        for name, t in self._maps.items():
            if t is not None:
                self._maps[name] = fn(name, t)
        return self

    def resize(self, size: Union[int, Tuple[int, int]], antialias: bool = True):
        from torchvision.transforms import functional as TF

        return self._each(lambda _n, t: TF.resize(t, size, antialias=antialias))

    def crop(self, top: int, left: int, height: int, width: int):
        from torchvision.transforms import functional as TF

        return self._each(lambda _n, t: TF.crop(t, top, left, height, width))

    def tile(self, num_tiles: int):
        """Repeat every map num_tiles x num_tiles (base.py:521-537: tensor.repeat)."""
        if self._all_cuda():
            return self._index_transform(lambda H, W: (H * num_tiles, W * num_tiles, 0, 1, 0, 1, True),
                                         adjoint=lambda H, W: (H, W, 0, 1, 0, 1, False, num_tiles, num_tiles))
        return self._each(lambda _n, t: t.repeat(*([1] * (t.dim() - 2)), num_tiles, num_tiles))

    def flip_horizontal(self):
        """Mirror along W; the normal's X component changes sign (base.py:605-621)."""
        if self._all_cuda():
            return self._index_transform(lambda H, W: (H, W, 0, 1, W - 1, -1, False), negate={"normal": 0b001},
                                         adjoint=lambda H, W: (H, W, 0, 1, W - 1, -1, False, 0, 0))

        def f(name, t):
            out = t.flip(-1)
            if name == "normal":
                out = out.clone()
                out.select(_channel_dim(out), 0).neg_()
            return out

        return self._each(f)

    def flip_vertical(self):
        """Mirror along H; the normal's Y component changes sign (base.py:623-639)."""
        if self._all_cuda():
            return self._index_transform(lambda H, W: (H, W, H - 1, -1, 0, 1, False), negate={"normal": 0b010},
                                         adjoint=lambda H, W: (H, W, H - 1, -1, 0, 1, False, 0, 0))

        def f(name, t):
            out = t.flip(-2)
            if name == "normal":
                out = out.clone()
                out.select(_channel_dim(out), 1).neg_()
            return out

        return self._each(f)

    def roll(self, shift: Tuple[int, int]):
        """torch.roll(map, shift, dims=(H, W)) on every map (base.py:641-655)."""
        if self._all_cuda():
            return self._index_transform(lambda H, W: (H, W, (-int(shift[0])) % H, 1, (-int(shift[1])) % W, 1, True),
                                         adjoint=lambda H, W: (H, W, int(shift[0]) % H, 1, int(shift[1]) % W, 1, True, 0, 0))
        return self._each(lambda _n, t: torch.roll(t, shift, dims=(-2, -1)))

    # -- CUDA path of the index transforms: every map of the material in ONE gather kernel (pbr_index_transform)
    def _all_cuda(self) -> bool:
        ts = [t for t in self._maps.values() if t is not None]
        return bool(ts) and all(t.is_cuda and t.dtype == torch.float32 for t in ts)

    def _index_transform(self, geometry, negate: Optional[Dict[str, int]] = None, adjoint=None):
        """
        geometry(H_in, W_in) -> (H_out, W_out, origin_y, step_y, origin_x, step_x, wrap).  Maps that share
        (batch, H, W) travel in one launch; results are bit-identical to the torch calls of the reference
        (pure data movement; the sign flip of a normal component is exact).
        adjoint(H_in, W_in) -> the same tuple + (reduce_y, reduce_x) describing the transposed gather (gradient w.r.t.
        the input from the gradient w.r.t. the output): maps that require grad go through an autograd.Function whose
        backward is that launch, so flip / roll / tile stay differentiable like the reference's tensor.flip / roll / repeat.
        """
        negate = negate or {}
        groups: Dict[tuple, list] = {}
        for name, t in self._maps.items():
            if t is None:
                continue
            if t.shape[-3] > 4:
                raise ValueError(f"map '{name}' has {t.shape[-3]} channels; at most 4 are supported per map")
            key = (t.shape[0] if t.dim() == 4 else 1, t.dim(), t.shape[-2], t.shape[-1], t.device)
            groups.setdefault(key, []).append(name)
        for (B, _dim, H, W, device), names in groups.items():
            fwd = (*geometry(H, W), 0, 0)
            for start in range(0, len(names), _cabi.PBR_MAX_INDEX_MAPS):
                part = names[start : start + _cabi.PBR_MAX_INDEX_MAPS]
                srcs = [self._maps[n] for n in part]
                negs = [int(negate.get(n, 0)) for n in part]
                if torch.is_grad_enabled() and any(t.requires_grad for t in srcs):
                    if adjoint is None:
                        raise RuntimeError("pypbr_b200: this index transform has no adjoint; detach the maps first")
                    outs = _IndexFn.apply((H, W), fwd, (*adjoint(H, W),), tuple(negs), *srcs)
                else:
                    outs = _index_launch([t.detach() for t in srcs], (H, W), fwd, negs)
                for n, out in zip(part, outs):
                    self._maps[n] = out
        return self

    def apply_transform(self, transform):
        return self._each(lambda _n, t: transform(t))

    def rotate(self, angle: float, expand: bool = False, padding_mode: str = "constant"):
        """
        Rotate all maps by `angle` degrees (pypbr/materials/base.py:539-603).  The resampling (pad, TF.rotate with
        nearest-neighbour lookup, centre crop) is the reference's own sequence of library calls on the maps' device,
        hence identical texel for texel; the normal VECTORS are then rotated by pbr_normal_op (rotate_normals,
        utils/functions.py:69-108).
        """
        from torchvision.transforms import functional as TF

        assert padding_mode in ["constant", "circular"], "Invalid padding mode. Must be 'constant' or 'circular'."
        angle_rad = math.radians(angle)
        for name, t in self._maps.items():
            if t is None:
                continue
            height, width = t.shape[-2:]
            if expand:
                new_width = math.ceil(abs(width * math.cos(angle_rad)) + abs(height * math.sin(angle_rad)))
                new_height = math.ceil(abs(width * math.sin(angle_rad)) + abs(height * math.cos(angle_rad)))
                height, width = new_height, new_width
            pad = math.ceil(math.sqrt(height**2 + width**2)) - height
            padded = F.pad(t, (pad, pad, pad, pad), padding_mode)
            rotated = TF.center_crop(TF.rotate(padded, angle, expand=True), (height, width)).contiguous()
            if name == "normal":
                rotated = rotate_normals(rotated, angle) if rotated.is_cuda else _rotate_normals_host(rotated, angle)
            self._maps[name] = rotated
        return self

    # ------------------------------------------------------------------ normal helpers
    def invert_normal(self):
        """Flip the Y component (utils/functions.py:107-120) and toggle the convention."""
        normal = self._maps.get("normal", None)
        if normal is not None:
            normal = normal.clone()
            normal.select(_channel_dim(normal), 1).neg_()
        self._maps["normal"] = normal
        self.normal_convention = (
            NormalConvention.DIRECTX if self.normal_convention == NormalConvention.OPENGL else NormalConvention.OPENGL
        )
        return self

    def adjust_normal_strength(self, strength_factor: float):
        """
        pypbr/materials/base.py:689-706: scale the normal's x / y by `strength_factor`, renormalise.  On CUDA maps the scale and
        the normalisation are one pbr_normal_op launch (ROTATE with cos = strength_factor, sin = 0: x*f - y*0 and x*0 + y*f are
        exactly RN(x*f), RN(y*f)).  The reference also scales the x / y of the ORIGINAL tensor in place (`normal[:2] *= f`, visible
        to every material that aliases it); that side effect is kept.
        """
        if self.normal is not None:
            normal = self.normal
            cd = _channel_dim(normal)
            if normal.is_cuda and normal.dtype == torch.float32:
                if torch.is_grad_enabled() and normal.requires_grad:
                    # differentiable form (pbr_normal_op ROTATE_BWD): a new map, the source left as it is - autograd needs it
                    from ..utils.functions import _RotateNormalsFn

                    self._maps["normal"] = _RotateNormalsFn.apply(normal, float(strength_factor), 0.0)
                    return self
                from ..utils.functions import _normal_op

                src = _cabi.rowmajor(normal.detach())
                out = torch.empty(src.shape, dtype=torch.float32, device=src.device)
                _normal_op(src, out, _cabi.NORMAL_OP_ROTATE, cos_a=float(strength_factor), sin_a=0.0)
                normal.narrow(cd, 0, 2).mul_(strength_factor)
                self._maps["normal"] = out
                return self
            normal.narrow(cd, 0, 2).mul_(strength_factor)
            self._maps["normal"] = F.normalize(normal, dim=cd)
        return self

    def compute_normal_from_height(self, scale: float = 1.0):
        """pypbr/materials/base.py:708-729: normal map from the height map (pbr_normal_op), in this material's convention."""
        self._maps["normal"] = compute_normal_from_height(self._maps.get("height", None), scale, convention=self.normal_convention)
        return self

    def compute_height_from_normal(self, scale: float = 1.0):
        """pypbr/materials/base.py:731-751: height map from the normal map (pbr_normal_op DIVERGENCE + cuFFT Poisson solve)."""
        self._maps["height"] = compute_height_from_normal(self._maps.get("normal", None), scale, convention=self.normal_convention)
        return self

    # ------------------------------------------------------------------ colour space
    def to_linear(self):
        albedo = self._maps.get("albedo", None)
        if albedo is not None and self.albedo_is_srgb:
            self._maps["albedo"] = srgb_to_linear(albedo)
            self.albedo_is_srgb = False
        return self

    def to_srgb(self):
        albedo = self._maps.get("albedo", None)
        if albedo is not None and not self.albedo_is_srgb:
            self._maps["albedo"] = linear_to_srgb(albedo)
            self.albedo_is_srgb = True
        return self

    # ------------------------------------------------------------------ export
    def to_numpy(self):
        return {n: (t.detach().cpu().numpy() if t is not None else None) for n, t in self._maps.items()}

    def to_pil(self, maps_mode: Dict[str, str] = None):
        """PIL images of every map (base.py:793-850); host-side convenience."""
        from torchvision.transforms import functional as TF

        maps_mode = maps_mode or {}
        out = {}
        for name, t in self._maps.items():
            if t is None:
                out[name] = None
                continue
            if name == "normal":
                if t.shape[0] == 2:
                    t = self._compute_normal_map_z_component(t)
                t = (t + 1.0) * 0.5
            t = t.detach().cpu()
            mode = maps_mode.get(name, "RGB")
            if mode in ("I", "I;16", "I;16B", "I;16L", "I;16N"):
                arr = t.numpy()
                if arr.ndim == 3 and arr.shape[0] == 1:
                    arr = arr[0]
                out[name] = Image.fromarray((arr * 65535).clip(0, 65535).astype(np.uint16), mode="I;16")
            else:
                out[name] = TF.to_pil_image(t).convert("RGB" if t.shape[0] == 3 else "L")
        return out

    def save_to_folder(self, folder_path: str):
        """Save the maps as images (base.py:869-878 -> pypbr/io.py:189-230)."""
        from ..io import save_material_to_folder

        save_material_to_folder(self, folder_path)

    # ------------------------------------------------------------------ misc
    def __repr__(self):
        parts = [f"{n}={(tuple(t.shape) if t is not None else None)}" for n, t in self._maps.items()]
        return f"{self.__class__.__name__}(" + ", ".join(parts) + ")"

    def clone(self):
        """Deep copy: every map tensor is cloned (base.py:880-912)."""
        new = self.__class__.__new__(self.__class__)
        for key, val in self.__dict__.items():
            if key == "_maps":
                object.__setattr__(new, "_maps", {n: (t.clone() if t is not None else None) for n, t in val.items()})
            else:
                object.__setattr__(new, key, copy.copy(val))
        return new


# ---------------------------------------------------------------------- CUDA normal ingestion
def _rotate_normals_host(normal: torch.Tensor, angle: float) -> torch.Tensor:
    """rotate_normals (utils/functions.py:69-108) for a material still being assembled on the host (CPU tensors)."""
    theta = math.radians(angle)
    cd = _channel_dim(normal)
    x, y, z = normal.unbind(cd)
    rot = torch.stack([x * math.cos(theta) - y * math.sin(theta), x * math.sin(theta) + y * math.cos(theta), z], dim=cd)
    return F.normalize(rot, dim=cd)


def _normal_desc(t: torch.Tensor, out: Optional[torch.Tensor], channels: int) -> "_cabi.PbrNormalDesc":
    src = _cabi.rowmajor(t)
    B = src.shape[0] if src.dim() == 4 else 1
    return _cabi.PbrNormalDesc(B, src.shape[-2], src.shape[-1], channels, _cabi.plane(src), _cabi.plane(out)), src


# min() of a normal map that has already been probed, keyed by tensor identity and guarded by its version counter, so
# that handing the same map to another material (every conversion does: metallic.py:112-118 passes normal=self.normal
# through the constructor, which re-runs base.py:210-217) costs neither a reduction nor a device->host read-back.
# In-place writers that bypass torch's version counter (the kernels' in-place entry points) bump it: _cabi.touch().
_PROBED: Dict[int, tuple] = {}


def _normal_min_cuda(t: torch.Tensor) -> float:
    """The `normal_map.min() < 0` probe of base.py:212 as one reduction kernel + one 4-byte readback (memoised)."""
    _cabi.require_cuda(t, "normal")
    hit = _PROBED.get(id(t))
    if hit is not None and hit[0]() is t and hit[1] == t._version:
        return hit[2]
    lib = _cabi.load()
    res = torch.full((1,), float("inf"), dtype=torch.float32, device=t.device)
    d, _keep = _normal_desc(t.detach(), None, t.shape[_channel_dim(t)])
    with torch.cuda.device(t.device):
        _cabi.check(lib.pbr_normal_min(_cabi.byref(d), res.data_ptr(), _cabi.stream_ptr(t.device)), "pbr_normal_min")
    value = float(res.item())
    key = id(t)
    _PROBED[key] = (weakref.ref(t, lambda _r, k=key: _PROBED.pop(k, None)), t._version, value)
    return value


def _normal_ingest_cuda(t: torch.Tensor, channels: int, out: Optional[torch.Tensor] = None,
                        cond_min: Optional[torch.Tensor] = None) -> torch.Tensor:
    """pbr_normal_ingest.  out: write there (out is t: in place, 3 channels only).  cond_min: DEVICE scalar; the launch
    leaves the map alone when it is negative (base.py:212 decided on the device, no read-back)."""
    _cabi.require_cuda(t, "normal")
    lib = _cabi.load()
    if out is None:
        shape = list(t.shape)
        shape[_channel_dim(t)] = 3
        out = torch.empty(shape, dtype=torch.float32, device=t.device)
    d, _keep = _normal_desc(t.detach(), out, channels)
    if cond_min is not None:
        d.cond_min = cond_min.data_ptr()
    with torch.cuda.device(t.device):
        _cabi.check(lib.pbr_normal_ingest(_cabi.byref(d), _cabi.stream_ptr(t.device)), "pbr_normal_ingest")
    if out is t:
        _cabi.touch(t)   # written in place through its raw pointer
    return out


class _NormalIngestFn(torch.autograd.Function):
    """normalize(n*2-1) / the 2-channel z reconstruction with their adjoint kernel (base.py:215-217, :235-242 are
    differentiable torch ops in the reference)."""

    @staticmethod
    def forward(ctx, t, channels: int):
        src = _cabi.rowmajor(t.detach())
        ctx.save_for_backward(src)
        ctx.channels = channels
        return _normal_ingest_cuda(src, channels)

    @staticmethod
    def backward(ctx, g):
        (src,) = ctx.saved_tensors
        lib = _cabi.load()
        g = _cabi.rowmajor(g)
        d_in = torch.empty(src.shape, dtype=torch.float32, device=src.device)
        d, _keep = _normal_desc(src, None, ctx.channels)
        gr = _cabi.PbrNormalGrads(_cabi.plane(g), _cabi.plane(d_in))
        with torch.cuda.device(src.device):
            _cabi.check(lib.pbr_normal_ingest_backward(_cabi.byref(d), _cabi.byref(gr), _cabi.stream_ptr(src.device)),
                        "pbr_normal_ingest_backward")
        return d_in, None


def normal_ingest(t: torch.Tensor, channels: int) -> torch.Tensor:
    if torch.is_grad_enabled() and t.requires_grad:
        return _NormalIngestFn.apply(t, channels)
    return _normal_ingest_cuda(t, channels)


# ---------------------------------------------------------------------- CUDA index transforms
def _index_launch(srcs, size, geom, negs):
    """One pbr_index_transform launch over maps that share (batch, H, W).  geom = (H_out, W_out, origin_y, step_y,
    origin_x, step_x, wrap, reduce_y, reduce_x)."""
    lib = _cabi.load()
    H, W = size
    H_out, W_out, oy, sy, ox, sx, wrap, ry, rx = geom
    first = srcs[0]
    d = _cabi.PbrIndexDesc()
    d.B = first.shape[0] if first.dim() == 4 else 1
    d.H_in, d.W_in, d.H_out, d.W_out = H, W, H_out, W_out
    d.origin_y, d.step_y, d.origin_x, d.step_x, d.wrap = oy, sy, ox, sx, int(wrap)
    d.reduce_y, d.reduce_x = ry, rx
    d.n_maps = len(srcs)
    keep, outs = [], []
    for i, t in enumerate(srcs):
        src = _cabi.rowmajor(t)
        out = torch.empty(*src.shape[:-2], H_out, W_out, dtype=torch.float32, device=src.device)
        d.maps[i] = _cabi.PbrIndexMap(_cabi.plane(src), _cabi.plane(out), src.shape[-3], negs[i])
        keep.append(src)
        outs.append(out)
    with torch.cuda.device(first.device):
        _cabi.check(lib.pbr_index_transform(_cabi.byref(d), _cabi.stream_ptr(first.device)), "pbr_index_transform")
    return outs


class _IndexFn(torch.autograd.Function):
    """flip / roll / tile of several maps in one gather; the backward is the transposed gather, one launch as well."""

    @staticmethod
    def forward(ctx, size, fwd, adj, negs, *srcs):
        ctx.size, ctx.adj, ctx.negs = size, adj, negs
        ctx.out_size = (fwd[0], fwd[1])
        ctx.metas = [(t.shape, t.device) for t in srcs]
        return tuple(_index_launch([t.detach() for t in srcs], size, fwd, list(negs)))

    @staticmethod
    def backward(ctx, *gs):
        gs = [g if g is not None else torch.zeros(*shape[:-2], *ctx.out_size, dtype=torch.float32, device=dev)
              for g, (shape, dev) in zip(gs, ctx.metas)]
        outs = _index_launch(gs, ctx.out_size, ctx.adj, list(ctx.negs))
        return (None, None, None, None, *outs)
