"""
Image ingestion on the device (pbr_ingest_image): an 8- or 16-bit image is uploaded as it is (1-6 bytes per texel
instead of 4-12) and converted to the float32 channel-planar map the kernels read - `/255` (TF.to_tensor) or
`/65535` (pypbr/materials/base.py:122-168), and for normal maps the `*2-1` + normalise of
MaterialBase._process_normal_map (base.py:215-217) or the z reconstruction of a 2-channel map (base.py:223-242),
in the same pass.  Values are bit-identical to the reference's torch ops (IEEE division, same normalisation code
as pbr_normal_ingest).
"""

from __future__ import annotations

import numpy as np
import torch

from .. import _cabi

PLAIN, NORMAL3, NORMAL2 = _cabi.INGEST_PLAIN, _cabi.INGEST_NORMAL3, _cabi.INGEST_NORMAL2


def ingest_uint(src: torch.Tensor, device, mode: int = PLAIN, channels: int = None) -> torch.Tensor:
    """
    src: (H, W), (H, W, C) or (B, H, W, C) tensor of dtype uint8, or uint16 / int16 (the 16 bits are read as unsigned),
         interleaved channels (the layout image decoders produce), on the host or already on `device`.
    mode: PLAIN -> (C', H, W) with C' = `channels` or C; NORMAL3 / NORMAL2 -> (3, H, W) unit normals.
    Returns a float32 CUDA tensor, (B, ...) if src was batched.
    """
    lib = _cabi.load()
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("pypbr_b200: ingest_uint runs on CUDA only (no CPU fallback)")
    if src.dtype == torch.uint8:
        bits = 8
    elif src.dtype in (torch.int16, torch.uint16):
        bits = 16
        src = src.view(torch.int16)
    else:
        raise TypeError(f"ingest_uint expects uint8 or (u)int16 data, got {src.dtype}")
    if src.dim() == 2:
        src = src.unsqueeze(-1)
    batched = src.dim() == 4
    if src.dim() not in (3, 4):
        raise ValueError(f"expected (H,W), (H,W,C) or (B,H,W,C), got shape {tuple(src.shape)}")
    src = src.contiguous().to(device, non_blocking=True)
    B = src.shape[0] if batched else 1
    H, W, C = src.shape[-3:]
    if C > 4:
        raise ValueError("at most 4 interleaved channels")
    need = {NORMAL3: 3, NORMAL2: 2}.get(mode, channels or C)
    if need > C:
        raise ValueError(f"mode needs {need} channels, the image has {C}")
    out_c = 3 if mode != PLAIN else need
    out = torch.empty((B, out_c, H, W) if batched else (out_c, H, W), dtype=torch.float32, device=device)
    d = _cabi.PbrIngestDesc()
    d.B, d.H, d.W, d.bits, d.src_channels, d.channels, d.mode = B, H, W, bits, C, need, mode
    d.src = src.data_ptr()
    esz = bits // 8
    d.src_row_stride = W * C * esz
    d.src_batch_stride = H * W * C * esz
    d.out = _cabi.plane(out)
    with torch.cuda.device(device):
        _cabi.check(lib.pbr_ingest_image(_cabi.byref(d), _cabi.stream_ptr(device)), "pbr_ingest_image")
    return out


def ingest_pil(image, device, is_normal: bool):
    """PIL image -> device map, or None when the mode is not one the kernel covers (caller falls back to torchvision)."""
    mode = image.mode
    if mode in ("I;16", "I;16B", "I;16L", "I;16N", "I"):
        if is_normal:
            return None
        arr = np.array(image, dtype=np.uint16)
        return ingest_uint(torch.from_numpy(arr.view(np.int16)), device, PLAIN)
    if mode == "RGBA":
        image = image.convert("RGB")
        mode = "RGB"
    if mode not in ("L", "RGB"):
        return None
    arr = np.asarray(image, dtype=np.uint8)
    if is_normal:
        if mode != "RGB":
            return None
        return ingest_uint(torch.from_numpy(np.ascontiguousarray(arr)), device, NORMAL3)
    return ingest_uint(torch.from_numpy(np.ascontiguousarray(arr)), device, PLAIN)
