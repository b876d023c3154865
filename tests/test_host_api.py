"""Host-side logic that needs no GPU: the C-ABI surface, the material container, argument errors, sharding."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT

import pypbr_b200
from pypbr_b200 import _cabi
from pypbr_b200.blending import BlendFactory, GradientBlend, MaskBlend, blend_materials
from pypbr_b200.materials import BasecolorMetallicMaterial, DiffuseSpecularMaterial, MaterialBase
from pypbr_b200.models import CookTorranceBRDF
from pypbr_b200.utils import NormalConvention


# ------------------------------------------------------------------ C ABI
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "pbrcuda.h")).read()
    declared = set(re.findall(r"^\s*(?:int|uint64_t|const char\*)\s+(pbr_\w+)\s*\(", header, re.M))
    assert declared == set(_cabi.EXPORTS)
    lib = ctypes.CDLL(_cabi.lib_path())
    for name in declared:
        assert hasattr(lib, name), name
    assert _cabi.load().pbr_abi_version() == _cabi.ABI_VERSION == 5


def test_struct_layouts_match_header_sizes():
    # sizes the C compiler produces for the same declarations (LP64): catches field-order drift
    assert ctypes.sizeof(_cabi.PbrPlane) == 32
    assert ctypes.sizeof(_cabi.PbrCtDesc) == 12 * 4 + 4 * 32 + 8 + 3 * 8 + 32 + 8 + 8
    assert ctypes.sizeof(_cabi.PbrCtGrads) == 32 + 8 + 4 * 32 + 3 * 8 + 32
    assert ctypes.sizeof(_cabi.PbrCtAdam) == 8 * 32 + 6 * 4 + 4 + 4  # trailing pad to 8
    assert ctypes.sizeof(_cabi.PbrBlendMap) == 3 * 32 + 8
    # ... and against the sizes the library itself reports (pbr_sizeof), for every descriptor
    lib = _cabi.load()
    for which, st in enumerate(_cabi.STRUCTS):
        assert lib.pbr_sizeof(which) == ctypes.sizeof(st), st.__name__
    assert lib.pbr_sizeof(99) == 0


def test_argument_errors_come_back_as_codes_without_touching_the_gpu():
    lib = _cabi.load()
    assert lib.pbr_ct_forward(None, None) == -1  # PBR_E_NULL
    d = _cabi.PbrCtDesc()
    d.B, d.H, d.W, d.L = 1, 0, 4, 1
    assert lib.pbr_ct_forward(ctypes.byref(d), None) == -2  # PBR_E_SHAPE
    d.H, d.L = 4, 65
    assert lib.pbr_ct_forward(ctypes.byref(d), None) == -4  # PBR_E_TOO_MANY
    d.L, d.workflow = 1, 7
    assert lib.pbr_ct_forward(ctypes.byref(d), None) == -3  # PBR_E_ENUM
    d.workflow = 0
    assert lib.pbr_ct_forward(ctypes.byref(d), None) == -1  # maps are NULL
    assert b"NULL" in lib.pbr_strerror(-1)
    b = _cabi.PbrBlendDesc()
    b.B, b.H, b.W, b.n_maps = 1, 4, 4, 13
    assert lib.pbr_blend(ctypes.byref(b), None) == -4
    # the one-launch fit step: descriptors are validated in the same order, and batch-broadcast maps (sb == 0) are
    # refused because they would be updated in place by every material
    assert lib.pbr_ct_fit_step(None, None, None, None, None) == -1
    fake = 0x1000   # never dereferenced: argument checks only
    d = _cabi.PbrCtDesc()
    d.B, d.H, d.W, d.L = 2, 4, 4, 1
    for pl in ("albedo", "roughness", "metspec"):
        setattr(d, pl, _cabi.PbrPlane(fake, 0, 16, 4))
    hv = (ctypes.c_float * 3)(0, 0, 1)
    d.view = d.lights = d.intensity = ctypes.cast(hv, ctypes.c_void_p)
    ls = _cabi.PbrCtLoss()
    assert lib.pbr_ct_fit_step(ctypes.byref(d), ctypes.byref(ls), None, None, None) == -1      # no target / Adam state
    ls.target, ls.loss_sum = _cabi.PbrPlane(fake, 48, 16, 4), fake
    a = _cabi.PbrCtAdam()
    assert lib.pbr_ct_fit_step(ctypes.byref(d), ctypes.byref(ls), ctypes.byref(a), None, None) == -2   # broadcast maps
    for pl in ("albedo", "roughness", "metspec"):
        setattr(d, pl, _cabi.PbrPlane(fake, 48, 16, 4))
    assert lib.pbr_ct_fit_step(ctypes.byref(d), ctypes.byref(ls), ctypes.byref(a), None, None) == -1   # moments are NULL
    # normal utilities
    n = _cabi.PbrNormalOpDesc()
    assert lib.pbr_normal_op(None, None) == -1
    n.B, n.H, n.W, n.op = 1, 4, 4, 9
    assert lib.pbr_normal_op(ctypes.byref(n), None) == -3
    n.op = _cabi.NORMAL_OP_FROM_HEIGHT
    assert lib.pbr_normal_op(ctypes.byref(n), None) == -1
    n.in_ = n.out = _cabi.PbrPlane(fake, 0, 16, 4)
    assert lib.pbr_normal_op(ctypes.byref(n), None) == -1    # a stencil cannot run in place
    for op in (_cabi.NORMAL_OP_FROM_HEIGHT_BWD, _cabi.NORMAL_OP_ROTATE_BWD, _cabi.NORMAL_OP_DIVERGENCE_BWD):
        n.op = op
        n.in_, n.out = _cabi.PbrPlane(fake, 0, 16, 4), _cabi.PbrPlane(fake + 4096, 0, 16, 4)
        n.aux = _cabi.PbrPlane(None, 0, 0, 0)
        assert lib.pbr_normal_op(ctypes.byref(n), None) == -1    # an adjoint needs the incoming gradient
        n.aux = n.out
        assert lib.pbr_normal_op(ctypes.byref(n), None) == -1    # ... and cannot write over it
    n.op = _cabi.NORMAL_OP_DIVERGENCE_BWD + 1
    assert lib.pbr_normal_op(ctypes.byref(n), None) == -3


# ------------------------------------------------------------------ material container
def _mat(H=8, W=12, cls=BasecolorMetallicMaterial, **kw):
    g = torch.Generator().manual_seed(0)
    n = torch.nn.functional.normalize(torch.randn(3, H, W, generator=g), dim=0)
    base = dict(albedo=torch.rand(3, H, W, generator=g), normal=n, roughness=torch.rand(1, H, W, generator=g))
    if cls is BasecolorMetallicMaterial:
        base["metallic"] = torch.rand(1, H, W, generator=g)
    else:
        base["specular"] = torch.rand(3, H, W, generator=g)
    base.update(kw)
    return cls(**base)


def test_attribute_protocol():
    m = _mat(height=torch.rand(1, 8, 12))
    assert set(m._maps) == {"albedo", "normal", "roughness", "metallic", "height"}
    assert m.albedo is m._maps["albedo"] and m.basecolor is m.albedo
    with pytest.raises(AttributeError):
        _ = m.specular
    assert not hasattr(m, "specular") and hasattr(m, "metallic")
    m.opacity = None  # None registers an empty map entry (base.py:96-101)
    assert "opacity" in m._maps and m.opacity is None
    m.note = "plain attribute"
    assert "note" not in m._maps and m.note == "plain attribute"
    m.f64 = torch.zeros(1, 2, 2, dtype=torch.float64)  # non-float32 tensors are plain attributes, as in the reference
    assert "f64" not in m._maps
    assert m.size == (8, 12)
    assert MaterialBase().size is None
    with pytest.raises(TypeError):
        MaterialBase(albedo=None)._to_tensor("nope")


def test_normal_ingestion_quirk_cpu():
    # all components >= 0 -> read as RGB-encoded and remapped; any negative -> aliased untouched
    rgb = torch.rand(3, 4, 4)
    m = MaterialBase(normal=rgb)
    assert torch.equal(m.normal, torch.nn.functional.normalize(rgb * 2 - 1, dim=0))
    signed = torch.nn.functional.normalize(torch.randn(3, 4, 4), dim=0)
    m = MaterialBase(normal=signed)
    assert m.normal is signed
    flat = torch.tensor([0.0, 0.0, 1.0]).view(3, 1, 1).expand(3, 2, 2).contiguous()
    m = MaterialBase(normal=flat)  # min() == 0 -> silently remapped (SURVEY.md §7 hard part 5)
    assert torch.allclose(m.normal[:, 0, 0], torch.tensor([-1.0, -1.0, 1.0]) / 3**0.5)
    with pytest.raises(ValueError):
        MaterialBase(normal=torch.rand(4, 2, 2))
    two = MaterialBase(normal=torch.rand(2, 4, 4)).normal
    assert two.shape == (3, 4, 4) and torch.allclose(two.norm(dim=0), torch.ones(4, 4), atol=1e-6)


def test_index_transforms_bit_exact_against_torch():
    m = _mat()
    a, n = m.albedo.clone(), m.normal.clone()
    f = m.clone().flip_horizontal()
    assert torch.equal(f.albedo, a.flip(-1))
    exp = n.flip(-1).clone(); exp[0] = -exp[0]
    assert torch.equal(f.normal, exp)
    f = m.clone().flip_vertical()
    exp = n.flip(-2).clone(); exp[1] = -exp[1]
    assert torch.equal(f.albedo, a.flip(-2)) and torch.equal(f.normal, exp)
    assert torch.equal(m.clone().roll((2, 3)).albedo, torch.roll(a, (2, 3), dims=(1, 2)))
    assert torch.equal(m.clone().tile(2).albedo, a.repeat(1, 2, 2))
    c = m.clone().crop(1, 2, 4, 6)
    assert torch.equal(c.albedo, a[:, 1:5, 2:8]) and c.size == (4, 6)
    from torchvision.transforms import functional as TF
    assert torch.equal(m.clone().resize((4, 6)).albedo, TF.resize(a, (4, 6), antialias=True))
    inv = m.clone().invert_normal()
    assert torch.equal(inv.normal[1], -n[1]) and inv.normal_convention == NormalConvention.DIRECTX


def test_clone_and_packing():
    m = _mat()
    c = m.clone()
    assert c.albedo is not m.albedo and torch.equal(c.albedo, m.albedo) and type(c) is type(m)
    t = m.as_tensor(names=["albedo", "normal", "roughness"])
    assert t.shape[0] == 7
    assert m.as_tensor([("albedo", 2), "roughness"]).shape[0] == 3
    with pytest.raises(KeyError):
        m.as_tensor(["nope"])
    with pytest.raises(ValueError):
        m.as_tensor([("albedo", 5)])
    r = MaterialBase.from_tensor(t, names=[("albedo", 3), ("normal", 3), ("roughness", 1)])
    assert torch.equal(r.albedo, m.albedo) and torch.equal(r.roughness, m.roughness)
    assert "BasecolorMetallicMaterial(" in repr(m)


# ------------------------------------------------------------------ error behaviour of the hot-path entry points
def test_brdf_constructor_and_cpu_refusal():
    with pytest.raises(ValueError):
        CookTorranceBRDF(light_type="spot")
    assert CookTorranceBRDF(light_type="POINT").light_type == "point"
    brdf = CookTorranceBRDF("directional")
    v = torch.tensor([0.0, 0.0, 1.0])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        brdf(_mat(), v, v, torch.ones(3))  # CPU material: loud failure, never a silent fallback
    with pytest.raises(AttributeError):
        brdf(MaterialBase(albedo=torch.rand(3, 4, 4), device=torch.device("cpu")), v, v, torch.ones(3))


def test_conversion_and_blend_errors_before_any_launch():
    # never-set map: AttributeError from __getattr__, as in the reference (metallic.py:84 reads self.metallic)
    with pytest.raises(AttributeError):
        BasecolorMetallicMaterial(albedo=torch.rand(3, 4, 4)).to_diffuse_specular_material()
    m = BasecolorMetallicMaterial(albedo=torch.rand(3, 4, 4))
    m.metallic = None  # explicit None entry -> the ValueError of metallic.py:84-87
    with pytest.raises(ValueError):
        m.to_diffuse_specular_material()
    d = DiffuseSpecularMaterial(albedo=torch.rand(3, 4, 4))
    d.specular = None
    with pytest.raises(ValueError):
        d.to_basecolor_metallic_material()
    with pytest.raises(RuntimeError, match="CUDA only"):
        _mat().to_diffuse_specular_material()
    m1, m2 = _mat(), _mat()
    with pytest.raises(ValueError):
        blend_materials(m1, m2, "mask")
    with pytest.raises(ValueError):
        blend_materials(m1, m2, "nope")
    with pytest.raises(ValueError):
        blend_materials(m1, m2, "mask", mask=torch.rand(2, 8, 12))
    with pytest.raises(ValueError):
        blend_materials(m1, m2, "height")
    with pytest.raises(ValueError):
        blend_materials(m1, m2, "properties", property_name="height")
    with pytest.raises(ValueError):
        blend_materials(m1, m2, "gradient", direction="diagonal")
    with pytest.raises(ValueError):
        MaskBlend(torch.rand(3, 4, 4))
    with pytest.raises(ValueError):
        GradientBlend("diagonal")
    with pytest.raises(ValueError):
        BlendFactory.get_blend_method("nope")
    assert isinstance(BlendFactory.get_blend_method("Gradient", direction="vertical"), GradientBlend)
    with pytest.raises(RuntimeError, match="CUDA only"):
        blend_materials(m1, m2, "mask", mask=torch.rand(1, 8, 12))


# ------------------------------------------------------------------ sharding / the one collective
def test_shard_range_partitions_exactly():
    from pypbr_b200.fit import shard_range

    for total in (1, 7, 64, 4096):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from pypbr_b200.fit import allreduce_loss_and_shared, shard_range

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    L = 2
    full = torch.arange(10 * (1 + 3 * L), dtype=torch.float32).view(10, 1 + 3 * L)  # per-material [loss, dI...]
    lo, hi = shard_range(10, rank, world)
    buf = full[lo:hi].sum(0)
    allreduce_loss_and_shared(buf)
    q.put((rank, buf.tolist()))
    dist.destroy_process_group()


def test_loss_allreduce_world2_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    expect = torch.arange(70, dtype=torch.float32).view(10, 7).sum(0).tolist()
    assert res[0] == expect and res[1] == expect


def _gloo_async_worker(rank, world, port, q):
    """The side-stream loss all-reduce of the one-launch fit step, on CPU tensors (gloo): PendingLoss.wait() / .item() deliver
    the reduced buffer, and the two-slot ring never hands out a buffer whose all-reduce is still pending."""
    import torch.distributed as dist
    from pypbr_b200.fit import PendingLoss, allreduce_loss_async

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    out = []
    pend = [None, None]
    for step in range(5):
        i = step % 2
        if pend[i] is not None:
            pend[i].wait()
        buf = torch.full((7,), float(rank + 1 + step))
        pend[i] = allreduce_loss_async(buf, scale=0.5)
        assert isinstance(pend[i], PendingLoss)
        out.append(pend[i])
    vals = [p.item() for p in out]
    bufs = [p.wait().tolist() for p in out]
    q.put((rank, vals, bufs))
    dist.destroy_process_group()


def test_async_loss_allreduce_world2_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_async_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r: (v, b) for r, v, b in (q.get(timeout=120) for _ in procs)}
    for p in procs:
        p.join(timeout=60)
    for step in range(5):
        total = (1 + step) + (2 + step)
        for r in (0, 1):
            assert res[r][0][step] == total * 0.5 and res[r][1][step] == [float(total)] * 7


def test_pending_loss_without_a_process_group_is_the_local_buffer():
    from pypbr_b200.fit import allreduce_loss_async

    buf = torch.tensor([3.0, 1.0])
    p = allreduce_loss_async(buf, scale=2.0)
    assert p.wait() is buf and p.item() == 6.0
