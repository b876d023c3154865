"""
pypbr_b200.transforms against the reference's own test strategy (tests/test_transforms.py:34-155 of the reference: output
sizes, flips equal tensor.flip, normal signs, colour round trip) - on CPU maps (the host branches of the material methods)
and, marked gpu, on CUDA maps where every map of a material moves in one pbr_index_transform launch - plus what the
reference's tests do not pin: the source material is never modified, the result never aliases it, seeded pipelines draw
the same parameters as the reference, and (when the reference is present on this machine) the results are identical.
"""
import os
import random
import shutil
import sys
import tempfile

import pytest
import torch

from pypbr_b200.materials import BasecolorMetallicMaterial
from pypbr_b200.transforms import (AdjustNormalStrength, CenterCrop, Compose, Crop, FlipHorizontal, FlipVertical, InvertNormal,
                                   RandomCrop, RandomHorizontalFlip, RandomResize, RandomRotate, RandomVerticalFlip, Resize, Roll,
                                   Rotate, Tile, ToLinear, ToSrgb)
from pypbr_b200.transforms import functional as TF_

DEV = torch.device("cuda:0") if torch.cuda.is_available() else None


def _maps(H=32, W=32, seed=0, batch=None):
    g = torch.Generator().manual_seed(seed)
    shp = (lambda c: (batch, c, H, W)) if batch else (lambda c: (c, H, W))
    up = torch.tensor([0.0, 0.0, 1.0]).view((1, 3, 1, 1) if batch else (3, 1, 1))
    normal = torch.nn.functional.normalize(torch.randn(shp(3), generator=g) * 0.4 + up, dim=-3)
    return dict(albedo=torch.rand(shp(3), generator=g), normal=normal, roughness=torch.rand(shp(1), generator=g),
                metallic=torch.rand(shp(1), generator=g))


def _material(device=None, **kw):
    maps = _maps(**kw)
    m = BasecolorMetallicMaterial(device=device or torch.device("cpu"))
    for k, v in maps.items():
        m._maps[k] = v.to(device) if device is not None else v
    return m, maps


def _snapshot(m):
    return {k: (t.clone(), t.data_ptr()) for k, t in m._maps.items() if t is not None}


def _assert_untouched_and_unaliased(src, before, new):
    for k, (val, ptr) in before.items():
        assert src._maps[k].data_ptr() == ptr and torch.equal(src._maps[k], val), f"{k}: the source material was modified"
        t = new._maps.get(k)
        if t is not None:
            assert t.untyped_storage().data_ptr() != src._maps[k].untyped_storage().data_ptr(), f"{k}: result aliases the source"


# every entry: (transform instance, expected size given (H, W) = (32, 32))
SIZE_CASES = [
    (Resize((64, 48)), (64, 48)), (Crop(top=5, left=5, height=20, width=18), (20, 18)), (CenterCrop(height=16, width=12), (16, 12)),
    (RandomCrop(height=16, width=16), (16, 16)), (Tile(num_tiles=2), (64, 64)), (Rotate(angle=90, expand=False), (32, 32)),
    (RandomRotate(min_angle=0, max_angle=360), (32, 32)), (Roll(shift=(2, 3)), (32, 32)), (FlipHorizontal(), (32, 32)),
    (FlipVertical(), (32, 32)), (InvertNormal(), (32, 32)), (AdjustNormalStrength(1.5), (32, 32)),
    (Compose([Resize((64, 64)), Crop(10, 10, 40, 40)]), (40, 40)), (Resize((32, 32)), (32, 32)),
]


def _check_sizes_and_copies(device):
    for tr, size in SIZE_CASES:
        m, _ = _material(device)
        before = _snapshot(m)
        out = tr(m)
        assert out.size == size, (tr, out.size)
        assert out is not m and type(out) is type(m)
        _assert_untouched_and_unaliased(m, before, out)
    m, _ = _material(device)
    out = RandomResize(40, 80)(m)
    assert all(40 <= s <= 80 for s in out.size)


def _check_values(device):
    m, maps = _material(device, H=16, W=24)
    dev = lambda t: t.to(device) if device is not None else t
    a, n, r = dev(maps["albedo"]), dev(maps["normal"]), dev(maps["roughness"])
    fh = FlipHorizontal()(m)
    assert torch.equal(fh.albedo, a.flip(-1)) and torch.equal(fh.roughness, r.flip(-1))
    assert torch.equal(fh.normal[0], -n.flip(-1)[0]) and torch.equal(fh.normal[1:], n.flip(-1)[1:])
    fv = FlipVertical()(m)
    assert torch.equal(fv.albedo, a.flip(-2)) and torch.equal(fv.normal[1], -n.flip(-2)[1]) and torch.equal(fv.normal[0], n.flip(-2)[0])
    assert torch.equal(RandomHorizontalFlip()(m, p=1.0).albedo, a.flip(-1)) and torch.equal(RandomVerticalFlip()(m, p=1.0).albedo, a.flip(-2))
    assert torch.equal(RandomHorizontalFlip()(m, p=0.0).albedo, a) and torch.equal(RandomVerticalFlip()(m, p=0.0).normal, n)
    assert torch.equal(Roll((2, -3))(m).albedo, torch.roll(a, (2, -3), dims=(-2, -1)))
    assert torch.equal(Tile(3)(m).normal, n.repeat(1, 3, 3))
    assert torch.equal(Crop(1, 2, 8, 12)(m).albedo, a[:, 1:9, 2:14]) and Crop(1, 2, 8, 12)(m).albedo.is_contiguous()
    assert torch.equal(CenterCrop(8, 12)(m).roughness, r[:, 4:12, 6:18])
    inv = InvertNormal()(m)
    assert torch.equal(inv.normal[1], -n[1]) and torch.equal(inv.normal[0], n[0]) and inv.normal_convention != m.normal_convention
    strong = AdjustNormalStrength(2.0)(m)
    want = torch.nn.functional.normalize(n * dev(torch.tensor([2.0, 2.0, 1.0]).view(3, 1, 1)), dim=0)
    assert torch.allclose(strong.normal, want, rtol=0, atol=2e-7)
    assert torch.equal(m.normal, n)   # the reference's in-place side effect lands on the transform's private copy


def _check_seeded_draws(device):
    """The draws come from random.random() in the reference's order (functional.py:88-90, :153-154, :224)."""
    m, maps = _material(device)
    random.seed(7)
    u = [random.random() for _ in range(5)]
    random.seed(7)
    rr = RandomResize(40, 80)(m)
    assert rr.size == (int(40 + 40 * u[0]), int(40 + 40 * u[1]))
    rc = RandomCrop(10, 12)(m)
    top, left = int((32 - 10) * u[2]), int((32 - 12) * u[3])
    src = maps["albedo"].to(device) if device is not None else maps["albedo"]
    assert torch.equal(rc.albedo, src[:, top:top + 10, left:left + 12])
    angle = 0.0 + 360.0 * u[4]                       # the fifth draw of the stream
    assert torch.equal(RandomRotate()(m).albedo, Rotate(angle)(m).albedo)


def test_sizes_and_copies_on_host_maps():
    _check_sizes_and_copies(None)


def test_values_on_host_maps():
    _check_values(None)


def test_seeded_draws_on_host_maps():
    _check_seeded_draws(None)


def test_functional_names_and_signatures_match_the_reference_interface():
    import inspect

    want = {"resize": ["material", "size", "antialias"], "random_resize": ["material", "min_size", "max_size", "antialias"],
            "crop": ["material", "top", "left", "height", "width"], "center_crop": ["material", "crop_size"],
            "random_crop": ["material", "crop_size"], "tile": ["material", "num_tiles"],
            "rotate": ["material", "angle", "expand", "padding_mode"],
            "random_rotate": ["material", "min_angle", "max_angle", "expand", "padding_mode"],
            "flip_horizontal": ["material"], "flip_vertical": ["material"], "random_horizontal_flip": ["material", "p"],
            "random_vertical_flip": ["material", "p"], "roll": ["material", "shift"], "invert_normal_map": ["material"],
            "adjust_normal_strength": ["material", "strength_factor"], "to_linear": ["material"], "to_srgb": ["material"]}
    for name, params in want.items():
        assert list(inspect.signature(getattr(TF_, name)).parameters) == params, name
    with pytest.raises(AssertionError):
        Rotate(10.0, padding_mode="reflect")
    with pytest.raises(AssertionError):
        RandomRotate(padding_mode="edge")


def _reference():
    """The unmodified reference package, when this machine has it (the build container; never the GPU box)."""
    root = "/root/reference/pypbr"
    if not os.path.isdir(root):
        return None
    if "pypbr" not in sys.modules:
        tmp = tempfile.mkdtemp(prefix="pypbr_ref_")
        shutil.copytree(root, os.path.join(tmp, "pypbr"))
        with open(os.path.join(tmp, "pypbr", "_version.py"), "w") as f:   # git-ignored file pypbr/__init__.py imports
            f.write('__version__ = version = "0+reference"\n')
        sys.path.insert(0, tmp)
    import pypbr   # noqa: F401

    return sys.modules["pypbr"]


def test_same_results_as_the_reference_on_a_seeded_pipeline():
    ref = _reference()
    if ref is None:
        pytest.skip("reference sources not on this machine")
    import pypbr.transforms as RT
    from pypbr.materials import BasecolorMetallicMaterial as RefMaterial

    maps = _maps(H=24, W=40, seed=5)
    ours = BasecolorMetallicMaterial(device=torch.device("cpu"))
    theirs = RefMaterial()
    for k, v in maps.items():
        ours._maps[k] = v.clone()
        theirs._maps[k] = v.clone()

    def pipeline(T):
        return T.Compose([T.RandomCrop(16, 32), T.FlipHorizontal(), T.Roll((3, -5)), T.Tile(2), T.RandomResize(20, 30),
                          T.RandomRotate(0.0, 90.0), T.FlipVertical(), T.InvertNormal(), T.CenterCrop(12, 12)])

    import pypbr_b200.transforms as OT

    random.seed(11)
    a = pipeline(OT)(ours)
    random.seed(11)
    b = pipeline(RT)(theirs)
    assert a.size == b.size
    for k in maps:
        if k == "normal":
            # rotate_normals is a 2 x 2 matmul in the reference (utils/functions.py:98, BLAS: the FMA order is its own) -
            # the rotated vectors agree to an ulp, the resampling (which texel goes where) exactly
            assert torch.allclose(a._maps[k], b._maps[k], rtol=0, atol=3e-7), k
        else:
            assert torch.equal(a._maps[k], b._maps[k]), k
    assert a.normal_convention.name == b.normal_convention.name


# ------------------------------------------------------------------------------------------------------------ CUDA maps
@pytest.mark.gpu
def test_sizes_and_copies_on_cuda_maps():
    _check_sizes_and_copies(DEV)


@pytest.mark.gpu
def test_values_on_cuda_maps():
    _check_values(DEV)


@pytest.mark.gpu
def test_seeded_draws_on_cuda_maps():
    _check_seeded_draws(DEV)


@pytest.mark.gpu
def test_colour_space_transforms_round_trip_on_cuda_maps():
    """tests/test_transforms.py:143-148 of the reference, at the kernels' tolerance instead of 1e-2."""
    m, maps = _material(DEV, H=16, W=16)
    before = _snapshot(m)
    lin = ToLinear()(m)
    assert not lin.albedo_is_srgb and m.albedo_is_srgb
    back = ToSrgb()(lin)
    assert back.albedo_is_srgb and torch.allclose(back.albedo, m.albedo, rtol=1e-5, atol=1e-6)
    _assert_untouched_and_unaliased(m, before, lin)
    from oracle import pbr_oracle as O

    assert torch.allclose(lin.albedo.cpu(), O.srgb_to_linear(maps["albedo"]), rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_batched_material_pipeline_on_cuda_maps_matches_host_maps():
    """A (B, C, H, W) material through an augmentation pipeline: the CUDA path (gather kernels) and the host path agree bit
    for bit on the pure index transforms."""
    pipe = Compose([Crop(2, 4, 24, 16), FlipHorizontal(), Roll((5, -3)), Tile(2), FlipVertical(), CenterCrop(20, 20), InvertNormal()])
    md, _ = _material(DEV, batch=3, seed=9)
    mh, _ = _material(None, batch=3, seed=9)
    od, oh = pipe(md), pipe(mh)
    assert od.size == oh.size == (20, 20)
    for k in oh._maps:
        assert torch.equal(od._maps[k].cpu(), oh._maps[k]), k
