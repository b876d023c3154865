"""
Parity of the CUDA path against the oracle / golden fixtures, through the public API (which reaches the
kernels through the C ABI).  Tolerances are BASELINE.json's: forward rel 1e-5 (+1e-6 abs), gradients
rel 1e-4 (+1e-4 * mean|g| abs); index transforms, given-mask blends and linear conversions bit-exact.
"""
import numpy as np
import pytest
import torch

from conftest import case_inputs, fwd_ok, golden_ct_cases, grad_ok, load_golden

pytestmark = pytest.mark.gpu

DEV = torch.device("cuda:0") if torch.cuda.is_available() else None


@pytest.fixture(params=["auto", "generic"])
def ct_path(request):
    """Run a test once with the automatic kernel choice (streamed TMA-fed kernels where they apply) and once
    pinned to the generic kernels (PbrCtDesc.force_generic)."""
    from pypbr_b200.models import cooktorrance as ct

    ct.FORCE_GENERIC = request.param == "generic"
    yield request.param
    ct.FORCE_GENERIC = False


def _material(maps, p, requires_grad=False, device=None):
    from pypbr_b200.materials import BasecolorMetallicMaterial, DiffuseSpecularMaterial

    device = device or DEV
    cls = BasecolorMetallicMaterial if "metallic" in maps else DiffuseSpecularMaterial
    m = cls(albedo_is_srgb=p.get("albedo_is_srgb", True), device=device)
    if cls is DiffuseSpecularMaterial:
        m.specular_is_srgb = p.get("specular_is_srgb", True)
    leaves = {}
    for k, v in maps.items():
        t = (torch.from_numpy(v) if isinstance(v, np.ndarray) else v).to(device).clone().requires_grad_(requires_grad)
        leaves[k] = t
        m._maps[k] = t
    if "normal" not in maps:
        m._maps["normal"] = None
    return m, leaves


def _brdf(p, per_light):
    from pypbr_b200.models import CookTorranceBRDF

    return CookTorranceBRDF(light_type=p["light_type"], multi_light="per_light" if per_light else "accumulate")


@pytest.mark.parametrize("name", golden_ct_cases())
def test_cooktorrance_golden_forward_backward(name, ct_path):
    z = load_golden(name)
    maps, view, lights, inten, p, multi, per_light = case_inputs(z)
    mat, leaves = _material(maps, p, requires_grad=True)
    lt = torch.from_numpy(z["lights"])
    it = torch.from_numpy(z["intensity"])
    out = _brdf(p, per_light)(mat, torch.from_numpy(view), lt, it, p["light_size"], p["return_srgb"])
    assert tuple(out.shape) == z["out32"].shape
    ratio, ok = fwd_ok(out.detach().cpu().numpy(), z["out32"])
    assert ok, f"forward max err/tol {ratio}"
    e_ref = np.abs(z["out32"] - z["out64"]).max()
    e_us = np.abs(out.detach().cpu().numpy() - z["out64"]).max()
    assert e_us <= 2 * e_ref + 1e-6
    out.backward(torch.from_numpy(z["grad_out"]).to(DEV))
    for k in maps:
        ratio, ok = grad_ok(leaves[k].grad.cpu().numpy(), z["g32_" + k])
        assert ok, f"d_{k} max err/tol {ratio}"


def test_config1_tiles_fixture():
    """BASELINE.json configs[0]: the tiles fixture at 256x256 with the example's lighting."""
    z = load_golden("config1_tiles_256")
    maps, view, lights, inten, p, *_ = case_inputs(z)
    mat, _ = _material(maps, p)
    out = _brdf(p, False)(mat, torch.from_numpy(view), torch.from_numpy(z["lights"]), torch.from_numpy(z["intensity"]), 1.0)
    o = out.cpu().numpy()
    assert fwd_ok(o, z["out32"])[1]
    assert abs(float(o.mean()) - 0.4920087) < 2e-6


def test_override_device_moves_cpu_material():
    """The reference's `override_device` route: maps stay on the CPU, every call uploads them."""
    from pypbr_b200.models import CookTorranceBRDF

    z = load_golden("ct_metal_point_37x53")
    maps, view, lights, inten, p, *_ = case_inputs(z)
    mat, _ = _material(maps, p, device=torch.device("cpu"))
    out = CookTorranceBRDF("point", override_device=DEV)(mat, torch.from_numpy(view), torch.from_numpy(z["lights"]),
                                                         torch.from_numpy(z["intensity"]), 1.0)
    assert out.is_cuda and fwd_ok(out.cpu().numpy(), z["out32"])[1]


def _random_case(seed, B, H, W, L, workflow="metallic", rough_lo=0.2, normal=True):
    g = torch.Generator().manual_seed(seed)
    shp = lambda c: (B, c, H, W) if B else (c, H, W)
    cd = 1 if B else 0
    maps = {"albedo": torch.rand(shp(3), generator=g), "roughness": torch.rand(shp(1), generator=g) * (1 - rough_lo) + rough_lo}
    if normal:
        n = torch.randn(shp(3), generator=g)
        sc = torch.tensor([0.3, 0.3, 0.0]).view((1, 3, 1, 1) if B else (3, 1, 1))
        up = torch.tensor([0.0, 0.0, 1.0]).view((1, 3, 1, 1) if B else (3, 1, 1))
        maps["normal"] = torch.nn.functional.normalize(n * sc + up, dim=cd)
    if workflow == "metallic":
        maps["metallic"] = torch.rand(shp(1), generator=g)
    else:
        maps["specular"] = torch.rand(shp(3), generator=g)
    ang = torch.arange(L, dtype=torch.float32) * (6.2831853 / max(L, 1))
    lights = torch.stack([0.4 * torch.cos(ang), 0.4 * torch.sin(ang), torch.ones(L)], dim=1) if L > 1 else torch.tensor([0.1, 0.1, 1.0])
    inten = torch.ones(L, 3) / L if L > 1 else torch.ones(3)
    return maps, lights, inten, g


@pytest.mark.parametrize("B,H,W,L,acc,wf", [
    (None, 64, 64, 1, True, "metallic"),
    (3, 37, 53, 1, True, "metallic"),     # ragged W: scalar path; B>1 with the hoisted light geometry
    (5, 32, 64, 1, True, "specular"),     # B not a multiple of the materials-per-thread chunk
    (2, 40, 48, 5, True, "metallic"),     # two-pass accumulate backward
    (2, 24, 36, 3, False, "specular"),    # per-light outputs
    (None, 1, 1, 1, True, "metallic"),    # degenerate 1x1 (linspace of one step)
    (2, 3, 1030, 1, True, "metallic"),    # wider than one CTA row
    (19, 9, 256, 1, True, "metallic"),    # streamed path: 4 rows per tile, H % 4 != 0, B > one CTA's walk
    (2, 2, 1536, 1, True, "specular"),    # streamed path: second strip half empty
    (3, 5, 72, 1, False, "metallic"),     # streamed path: per_light flag with a single light, narrow image
    (19, 9, 40, 6, False, "metallic"),    # geometry cache (l, h cached): B > one CTA's walk, per-light outputs
    (18, 8, 33, 3, True, "specular"),     # geometry cache (all 8 fields): two-pass accumulate backward, ragged W
    (17, 4, 24, 12, True, "metallic"),    # geometry cache at 12 lights
    (3, 6, 24, 20, True, "metallic"),     # more lights than the cache takes: per-texel geometry, one material per CTA
    (2, 5, 20, 18, False, "specular"),    # ... per-light outputs
    (None, 16, 64, 6, True, "metallic"),  # ONE material under several lights: per-texel geometry, plain-case (kFmAccum) flavour
    (4, 8, 64, 3, True, "specular"),      # all-8-field cache, specular workflow, plain-case flavour
    (5, 6, 32, 7, True, "metallic"),      # 4 < L <= 8: forward on the 6-field cache, backward on the big all-field cache
])
def test_against_oracle_seeded(B, H, W, L, acc, wf, ct_path):
    """Fresh seeded inputs (not in the fixtures) against the oracle run on the host."""
    from oracle import pbr_oracle as O

    maps, lights, inten, g = _random_case(1234 + H + W, B, H, W, L, wf)
    view = torch.tensor([0.05, -0.1, 1.0])
    p = dict(light_type="point", light_size=1.0, return_srgb=True, albedo_is_srgb=True)
    leaves_ref = {k: v.clone().requires_grad_(True) for k, v in maps.items()}
    ref = O.render(leaves_ref, view, lights, inten, 1.0, "point", accumulate=acc)
    go = torch.rand(ref.shape, generator=g)
    ref.backward(go)
    mat, leaves = _material(maps, p, requires_grad=True)
    out = _brdf(p, not acc)(mat, view, lights, inten, 1.0)
    assert out.shape == ref.shape
    ratio, ok = fwd_ok(out.detach().cpu().numpy(), ref.detach().numpy())
    assert ok, f"forward {ratio}"
    out.backward(go.to(DEV))
    for k in maps:
        ratio, ok = grad_ok(leaves[k].grad.cpu().numpy(), leaves_ref[k].grad.numpy())
        assert ok, f"d_{k} {ratio}"


def test_strided_views_and_unaligned_pointers():
    """Crops are views (TF.crop slices): strides and a 4-byte-aligned base pointer take the scalar path."""
    maps, lights, inten, g = _random_case(77, 2, 40, 72, 1)
    p = dict(light_type="point")
    view = torch.tensor([0.0, 0.0, 1.0])
    full, _ = _material(maps, p)
    crop = {k: t[..., 3:35, 5:66] for k, t in full._maps.items()}       # unaligned start, odd width
    mat_view, _ = _material({k: v for k, v in crop.items()}, p)
    for k in crop:
        mat_view._maps[k] = crop[k]                                      # keep them as views
    mat_copy, _ = _material({k: v.contiguous() for k, v in crop.items()}, p)
    brdf = _brdf(p, False)
    a = brdf(mat_view, view, lights, inten, 1.0)
    b = brdf(mat_copy, view, lights, inten, 1.0)
    assert torch.equal(a, b)
    al = {k: t[..., 4:36, 8:72] for k, t in full._maps.items()}         # aligned crop: vector path on views
    mv, _ = _material(al, p)
    for k in al:
        mv._maps[k] = al[k]
    mc, _ = _material({k: v.contiguous() for k, v in al.items()}, p)
    assert torch.equal(brdf(mv, view, lights, inten, 1.0), brdf(mc, view, lights, inten, 1.0))


def test_full_size_properties():
    """
    At a BASELINE-sized image (1024x1024, B=4) the oracle is too slow, so size-independent properties:
    a batch equals its per-material calls bit-for-bit; the hoisted single-light path equals the generic
    multi-light path with a second, zero-intensity light to a few ulp (same per-texel code, but the compiler
    may contract multiply-adds of the tolerant zone differently in the two kernels); a horizontal flip of the maps (with the
    normal's x sign) under a mirrored light mirrors the image; unclamped output is linear in intensity.
    """
    from pypbr_b200.models import CookTorranceBRDF

    maps, lights, inten, g = _random_case(9, 4, 1024, 1024, 1)
    p = dict(light_type="point")
    view = torch.tensor([0.0, 0.0, 1.0])
    mat, _ = _material(maps, p)
    brdf = CookTorranceBRDF("point")
    full = brdf(mat, view, lights, inten, 1.0)
    assert full.shape == (4, 3, 1024, 1024) and bool(torch.isfinite(full).all())
    assert float(full.min()) >= 0.0 and float(full.max()) <= 1.0
    for b in (0, 3):
        one, _ = _material({k: v[b] for k, v in maps.items()}, p)
        assert torch.equal(brdf(one, view, lights, inten, 1.0), full[b])
    two = torch.stack([lights, torch.tensor([0.3, -0.2, 0.8])])
    two_i = torch.stack([inten, torch.zeros(3)])
    multi = brdf(mat, view, two, two_i, 1.0)
    assert bool(((multi - full).abs() <= 2e-6 * full.abs() + 2e-6 * full.abs().mean()).all())
    # mirror symmetry
    flipped = mat.clone().flip_horizontal()
    lf = lights * torch.tensor([-1.0, 1.0, 1.0])
    mirrored = brdf(flipped, view, lf, inten, 1.0)
    assert torch.allclose(mirrored, full.flip(-1), rtol=1e-5, atol=1e-6)
    # linearity in the light intensity below the clamp (linear output)
    lo = brdf(mat, view, lights, inten * 0.05, 1.0, False)
    lo2 = brdf(mat, view, lights, inten * 0.1, 1.0, False)
    sel = lo2 < 0.99
    assert torch.allclose(lo2[sel], 2 * lo[sel], rtol=2e-6, atol=1e-7)


def test_intensity_gradient_and_device_resident_light_parameters():
    from oracle import pbr_oracle as O

    maps, lights, inten, g = _random_case(5, 2, 24, 40, 3)
    view = torch.tensor([0.0, 0.1, 1.0])
    p = dict(light_type="point")
    ref_i = inten.clone().requires_grad_(True)
    ref = O.render({k: v for k, v in maps.items()}, view, lights, ref_i, 1.0, "point", accumulate=False)
    go = torch.rand(ref.shape, generator=g)
    ref.backward(go)
    mat, _ = _material(maps, p)
    dev_i = inten.to(DEV).requires_grad_(True)
    out = _brdf(p, True)(mat, view.to(DEV), lights.to(DEV), dev_i, 1.0)   # all parameters already on the device
    assert fwd_ok(out.detach().cpu().numpy(), ref.detach().numpy())[1]
    out.backward(go.to(DEV))
    ratio, ok = grad_ok(dev_i.grad.cpu().numpy(), ref_i.grad.numpy())
    assert ok, f"d_intensity {ratio}"


def test_fused_loss_step_matches_oracle_and_two_kernel_path():
    from oracle import pbr_oracle as O
    from pypbr_b200.fit import RenderingLoss, fused_loss_step

    maps, lights, inten, g = _random_case(21, 3, 32, 48, 4)
    view = torch.tensor([0.0, 0.0, 1.0])
    p = dict(light_type="point")
    leaves_ref = {k: v.clone().requires_grad_(True) for k, v in maps.items()}
    ref = O.render(leaves_ref, view, lights, inten, 1.0, "point", accumulate=False)
    target = torch.rand(ref.shape, generator=g)
    loss_ref = ((ref - target) ** 2).mean()
    loss_ref.backward()
    mat, leaves = _material(maps, p)
    buf, grads = fused_loss_step(mat, target.to(DEV), view, lights, inten, "point", 1.0, want_intensity_grad=True)
    loss = float(buf[0]) / target.numel()
    assert abs(loss - float(loss_ref.detach())) <= 2e-6 * float(loss_ref.detach())
    for k in maps:
        ratio, ok = grad_ok(grads[k].cpu().numpy(), leaves_ref[k].grad.numpy())
        assert ok, f"{k}: {ratio}"
    # nn.Module form (tutorial API), single light
    mat1, leaves1 = _material({k: v[0] for k, v in maps.items()}, p, requires_grad=True)
    gt, _ = _material({k: v[1] for k, v in maps.items()}, p)
    crit = RenderingLoss(light_type="point")
    l = crit(mat1, gt)
    l.backward()
    r1 = {k: v[0].clone().requires_grad_(True) for k, v in maps.items()}
    a = O.render(r1, crit.view_dir, crit.light_dir, crit.light_intensity, None, "point")
    b = O.render({k: v[1] for k, v in maps.items()}, crit.view_dir, crit.light_dir, crit.light_intensity, None, "point")
    lr = torch.nn.MSELoss()(a, b)
    lr.backward()
    assert abs(float(l) - float(lr.detach())) <= 2e-6 * float(lr.detach()) + 1e-9
    for k in maps:
        assert grad_ok(leaves1[k].grad.cpu().numpy(), r1[k].grad.numpy())[1], k


def test_workflow_conversions():
    from pypbr_b200.materials import BasecolorMetallicMaterial, DiffuseSpecularMaterial

    z = load_golden("convert_31x45")
    for srgb in (True, False):
        m = BasecolorMetallicMaterial(albedo=torch.from_numpy(z["in_m_albedo"]).to(DEV), normal=torch.from_numpy(z["in_m_normal"]).to(DEV),
                                      roughness=torch.from_numpy(z["in_m_roughness"]).to(DEV), metallic=torch.from_numpy(z["in_m_metallic"]).to(DEV),
                                      albedo_is_srgb=srgb, device=DEV, height=torch.rand(1, 31, 45, device=DEV))
        d = m.to_diffuse_specular_material()
        assert isinstance(d, DiffuseSpecularMaterial) and d.albedo_is_srgb is False and d.specular_is_srgb is True
        assert d.normal is m.normal and d.roughness is m.roughness and "height" not in d._maps
        rd, rs = z[f"m2s_diffuse_srgb{int(srgb)}"], z[f"m2s_specular_srgb{int(srgb)}"]
        if srgb:
            assert fwd_ok(d.albedo.cpu().numpy(), rd)[1] and fwd_ok(d.specular.cpu().numpy(), rs)[1]
        else:
            assert np.array_equal(d.albedo.cpu().numpy(), rd) and np.array_equal(d.specular.cpu().numpy(), rs)
        s = DiffuseSpecularMaterial(albedo=torch.from_numpy(z["in_s_albedo"]).to(DEV), normal=torch.from_numpy(z["in_s_normal"]).to(DEV),
                                    roughness=torch.from_numpy(z["in_s_roughness"]).to(DEV), specular=torch.from_numpy(z["in_s_specular"]).to(DEV),
                                    albedo_is_srgb=srgb, device=DEV)
        bm = s.to_basecolor_metallic_material()
        assert isinstance(bm, BasecolorMetallicMaterial) and bm.metallic.shape[0] == 3
        rb, rm = z[f"s2m_basecolor_srgb{int(srgb)}"], z[f"s2m_metallic_srgb{int(srgb)}"]
        b, mm = bm.albedo.cpu().numpy(), bm.metallic.cpu().numpy()
        if srgb:
            # EVERY texel (conftest.s2m_ok): inside rel 1e-5 of the reference, or - where the division by (diffuse - 0.04)
            # amplifies an ulp of ATen's own pow beyond that - not further from the reference's fp64 run than twice the
            # reference itself is.  The kernel decodes with a correctly rounded pow where |diffuse - 0.04| < 4e-3.
            from conftest import s2m_ok
            for got, key in ((b, "basecolor"), (mm, "metallic")):
                frac, ok = s2m_ok(got, z[f"s2m_{key}_srgb1"], z[f"s2m_{key}64_srgb1"])
                assert ok and frac > 0.995, (key, frac)
        else:
            assert np.array_equal(b, rb) and np.array_equal(mm, rm)
        # the converted material (3-channel metallic) still renders: metallic workflow with per-channel metallic
        from pypbr_b200.models import CookTorranceBRDF
        img = CookTorranceBRDF("point")(bm, torch.tensor([0.0, 0.0, 1.0]), torch.tensor([0.1, 0.1, 1.0]), torch.ones(3))
        assert img.shape == (3, 31, 45) and bool(torch.isfinite(img).all())


def _blend_inputs(z):
    from pypbr_b200.materials import BasecolorMetallicMaterial

    mats = []
    for tag in ("in1_", "in2_"):
        kw = {k[len(tag):]: torch.from_numpy(v).to(DEV) for k, v in z.items() if k.startswith(tag)}
        mats.append(BasecolorMetallicMaterial(device=DEV, **kw))
    return mats


def test_blends_match_reference_fixture():
    from pypbr_b200.blending import blend_materials, HeightBlend

    z = load_golden("blend_29x43")
    m1, m2 = _blend_inputs(z)
    mask = torch.from_numpy(z["mask"]).to(DEV)
    runs = {
        "mask": dict(method="mask", mask=mask),
        "mask2d": dict(method="mask", mask=mask[0]),
        "height": dict(method="height", blend_width=0.1),
        "height_w03": dict(method="height", blend_width=0.3),
        "prop_metallic": dict(method="properties", property_name="metallic", blend_width=0.1),
        "prop_roughness": dict(method="properties", property_name="roughness", blend_width=0.05),
        "grad_h": dict(method="gradient", direction="horizontal"),
        "grad_v": dict(method="gradient", direction="vertical"),
    }
    for tag, kw in runs.items():
        blended, used = blend_materials(m1, m2, **kw)   # tuple return, like the reference
        assert type(blended) is type(m1) and blended.albedo_is_srgb == m1.albedo_is_srgb
        exact = tag in ("mask", "mask2d", "grad_h", "grad_v")
        um = used.cpu().numpy()
        assert um.shape == z[tag + "_mask"].shape
        if exact:
            assert np.array_equal(um, z[tag + "_mask"]), tag
        else:
            assert np.allclose(um, z[tag + "_mask"], rtol=2e-6, atol=1e-7), tag
        for name in ("albedo", "normal", "roughness", "metallic", "height", "opacity"):
            got = blended._maps[name].cpu().numpy()
            ref = z[f"{tag}_{name}"]
            if exact:
                assert np.array_equal(got, ref), (tag, name)
            else:
                assert np.allclose(got, ref, rtol=1e-5, atol=2e-6), (tag, name)
        assert blended._maps["opacity"] is m2._maps["opacity"]          # one-sided map: passed by reference
    b2, _ = HeightBlend(0.1, shift=0.2)(m1, m2)
    from oracle import pbr_oracle as O
    hm = O.sigmoid_mask(torch.from_numpy(z["in1_height"]), torch.from_numpy(z["in2_height"]), 0.1, 0.2)
    exp = hm * torch.from_numpy(z["in1_albedo"]) + (1 - hm) * torch.from_numpy(z["in2_albedo"])
    assert np.allclose(b2.albedo.cpu().numpy(), exp.numpy(), rtol=1e-5, atol=2e-6)


def test_blend_normal_requantisation_quirk():
    """A blended normal with no negative component is re-read as RGB by setattr (base.py:210-217)."""
    from pypbr_b200.blending import blend_with_mask
    from pypbr_b200.materials import MaterialBase
    from oracle import pbr_oracle as O

    n1 = torch.nn.functional.normalize(torch.rand(3, 8, 12) + 0.1, dim=0)
    n2 = torch.nn.functional.normalize(torch.rand(3, 8, 12) + 0.1, dim=0)
    mask = torch.rand(1, 8, 12)
    a = MaterialBase(device=DEV); a._maps["normal"] = n1.to(DEV)
    b = MaterialBase(device=DEV); b._maps["normal"] = n2.to(DEV)
    out, _ = blend_with_mask(a, b, mask.to(DEV))
    exp = O.process_normal_map(O.blend_normals(n1, n2, mask))
    assert np.array_equal(out.normal.cpu().numpy(), exp.numpy())


def test_colour_space_and_normal_ingestion_kernels():
    from oracle import pbr_oracle as O
    from pypbr_b200.materials import MaterialBase
    from pypbr_b200.utils import linear_to_srgb, srgb_to_linear

    x = torch.linspace(-0.2, 1.2, 40000).view(1, 1, 200, 200)
    lin = srgb_to_linear(x.to(DEV)).cpu()
    assert fwd_ok(lin.numpy(), O.srgb_to_linear(x).numpy())[1]
    enc = linear_to_srgb(x.to(DEV)).cpu()
    assert fwd_ok(enc.numpy(), O.linear_to_srgb(x).numpy())[1]
    m = MaterialBase(albedo=torch.rand(3, 16, 20, device=DEV), device=DEV)
    ref = O.srgb_to_linear(m.albedo.cpu())
    assert fwd_ok(m.linear_albedo.cpu().numpy(), ref.numpy())[1]
    m.to_linear()
    assert m.albedo_is_srgb is False and fwd_ok(m.albedo.cpu().numpy(), ref.numpy())[1]
    rgb = torch.rand(3, 19, 27)
    assert np.array_equal(MaterialBase(normal=rgb.to(DEV), device=DEV).normal.cpu().numpy(), O.process_normal_map(rgb).numpy())
    signed = torch.nn.functional.normalize(torch.randn(3, 19, 27), dim=0).to(DEV)
    assert MaterialBase(normal=signed, device=DEV).normal is signed
    two = torch.rand(2, 19, 27)
    assert np.allclose(MaterialBase(normal=two.to(DEV), device=DEV).normal.cpu().numpy(), O.process_normal_map(two).numpy(), rtol=3e-7, atol=1e-8)


def test_native_library_is_the_code_path():
    """Launch counter moves with every call: the kernels, not a fallback, produce the results."""
    from pypbr_b200 import _cabi
    from pypbr_b200.models import CookTorranceBRDF

    maps, lights, inten, g = _random_case(1, None, 16, 16, 1)
    mat, _ = _material(maps, dict(light_type="point"))
    before = _cabi.launch_count()
    CookTorranceBRDF("point")(mat, torch.tensor([0.0, 0.0, 1.0]), lights, inten)
    assert _cabi.launch_count() == before + 1


@pytest.mark.parametrize("light_type", ["point", "directional"])
@pytest.mark.parametrize("wf,normal", [("metallic", True), ("metallic", False), ("specular", True)])
def test_streamed_kernels_equal_generic_kernels(light_type, wf, normal):
    """The streamed (TMA-fed) and the generic kernels run the same per-texel code; the compiler may contract
    multiply-adds of the tolerant zone differently in the two instantiations, so they must agree to a few ulp
    (far inside the parity tolerance); the fused loss / intensity reductions also differ in summation order."""
    from pypbr_b200.fit import fused_loss_step
    from pypbr_b200.models import CookTorranceBRDF
    from pypbr_b200.models import cooktorrance as ct

    maps, lights, inten, g = _random_case(31, 21, 12, 520, 1, wf, normal=normal)
    if light_type == "directional":
        lights = torch.tensor([0.2, -0.3, 0.9])
    view = torch.tensor([0.1, 0.0, 1.0])
    p = dict(light_type=light_type)
    res = {}
    for path in ("auto", "generic"):
        ct.FORCE_GENERIC = path == "generic"
        try:
            mat, leaves = _material(maps, p, requires_grad=True)
            it = inten.clone().to(DEV).requires_grad_(True)
            out = CookTorranceBRDF(light_type)(mat, view, lights, it, 1.0)
            go = torch.rand(out.shape, generator=torch.Generator().manual_seed(3)).to(DEV)
            out.backward(go)
            tgt = torch.rand(out.shape, generator=torch.Generator().manual_seed(4)).to(DEV)
            buf, grads = fused_loss_step(mat, tgt, view, lights, inten, light_type, 1.0, want_intensity_grad=True)
            res[path] = (out.detach().clone(), {k: v.grad.clone() for k, v in leaves.items()}, it.grad.clone(),
                         buf.clone(), {k: v.clone() for k, v in grads.items() if v is not None})
        finally:
            ct.FORCE_GENERIC = False
    a, b = res["auto"], res["generic"]

    def close(x, y, rtol):
        return bool(((x - y).abs() <= rtol * y.abs() + rtol * y.abs().mean()).all())

    assert close(a[0], b[0], 2e-6)
    for k in a[1]:
        assert close(a[1][k], b[1][k], 1e-5), k
    assert torch.allclose(a[2], b[2], rtol=1e-4, atol=1e-6)
    assert torch.allclose(a[3], b[3], rtol=1e-4, atol=1e-6)
    for k in a[4]:
        assert close(a[4][k], b[4][k], 1e-5), k


# ---------------------------------------------------------------------------------------------- SURVEY.md §8f rows
def test_index_transforms_are_bit_exact():
    """flip / roll / tile of every map in one gather kernel == the reference's torch calls (torch.equal), incl. the
    sign flip of the normal's x / y, ragged sizes, batched maps and maps of different sizes in one material."""
    from oracle import pbr_oracle as O
    from pypbr_b200.materials import BasecolorMetallicMaterial

    g = torch.Generator().manual_seed(11)
    for (H, W) in ((33, 47), (64, 128), (5, 3)):
        maps = dict(albedo=torch.rand(3, H, W, generator=g), normal=torch.randn(3, H, W, generator=g),
                    roughness=torch.rand(1, H, W, generator=g), metallic=torch.rand(1, H, W, generator=g),
                    height=torch.rand(1, H // 2 + 1, W // 2 + 2, generator=g))

        def fresh():
            m = BasecolorMetallicMaterial(device=DEV)
            for k, v in maps.items():
                m._maps[k] = v.to(DEV)
            return m

        for name, ours, ref in (
            ("flip_h", lambda m: m.flip_horizontal(), O.flip_horizontal(maps)),
            ("flip_v", lambda m: m.flip_vertical(), O.flip_vertical(maps)),
            ("roll", lambda m: m.roll((7, -5)), O.roll(maps, (7, -5))),
            ("roll_big", lambda m: m.roll((-2 * H - 1, 3 * W + 2)), O.roll(maps, (-2 * H - 1, 3 * W + 2))),
            ("tile", lambda m: m.tile(3), O.tile(maps, 3)),
        ):
            l0 = _cabi_launches()
            got = ours(fresh())
            assert _cabi_launches() > l0, "index transform did not go through the native library"
            for k in maps:
                assert torch.equal(got._maps[k].cpu(), ref[k]), (name, k, H, W)
    # batched maps: a batch is B independent materials
    B, H, W = 3, 18, 20
    bm = dict(albedo=torch.rand(B, 3, H, W, generator=g), normal=torch.randn(B, 3, H, W, generator=g))
    m = BasecolorMetallicMaterial(device=DEV)
    for k, v in bm.items():
        m._maps[k] = v.to(DEV)
    m.flip_horizontal()
    for b in range(B):
        ref = O.flip_horizontal({k: v[b] for k, v in bm.items()})
        for k in bm:
            assert torch.equal(m._maps[k][b].cpu(), ref[k])


def _cabi_launches():
    from pypbr_b200 import _cabi

    return _cabi.launch_count()


def test_image_ingestion_is_bit_exact():
    """uint8 / uint16 images straight to the device: /255, /65535 and the normal remap fused == the reference's
    TF.to_tensor + _process_normal_map (torch.equal); also through the material constructor with PIL images."""
    from PIL import Image

    from oracle import pbr_oracle as O
    from pypbr_b200.materials import BasecolorMetallicMaterial
    from pypbr_b200.materials._ingest import NORMAL2, NORMAL3, PLAIN, ingest_uint

    rng = np.random.default_rng(5)
    for (H, W) in ((40, 64), (17, 23)):
        rgb = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        rgb[0, :8, 0] = np.arange(8) * 36   # make sure 0 and 252.. are present
        gray = rng.integers(0, 256, (H, W), dtype=np.uint8)
        h16 = rng.integers(0, 65536, (H, W), dtype=np.uint16)
        assert torch.equal(ingest_uint(torch.from_numpy(rgb), DEV, PLAIN).cpu(), O.ingest_image(rgb))
        assert torch.equal(ingest_uint(torch.from_numpy(gray), DEV, PLAIN).cpu(), O.ingest_image(gray))
        assert torch.equal(ingest_uint(torch.from_numpy(h16.view(np.int16)), DEV, PLAIN).cpu(), O.ingest_image(h16))
        assert torch.equal(ingest_uint(torch.from_numpy(rgb), DEV, NORMAL3).cpu(), O.ingest_image(rgb, is_normal=True))
        two = O.process_normal_map(O.ingest_image(rgb)[:2])
        # 2-channel maps: z = sqrt(clamp(1 - x^2 - y^2)) then normalise - within an ulp of the CPU's sqrt / vector norm
        assert np.allclose(ingest_uint(torch.from_numpy(rgb), DEV, NORMAL2).cpu().numpy(), two.numpy(), rtol=3e-7, atol=1e-8)
        # every byte value
        ramp = np.arange(256, dtype=np.uint8).reshape(1, 256)
        assert torch.equal(ingest_uint(torch.from_numpy(ramp), DEV, PLAIN).cpu(), O.ingest_image(ramp))
        # batched (B, H, W, C)
        bat = rng.integers(0, 256, (3, H, W, 3), dtype=np.uint8)
        got = ingest_uint(torch.from_numpy(bat), DEV, PLAIN).cpu()
        for b in range(3):
            assert torch.equal(got[b], O.ingest_image(bat[b]))
        # through the constructor, PIL in -> CUDA maps out
        l0 = _cabi_launches()
        mat = BasecolorMetallicMaterial(albedo=Image.fromarray(rgb), normal=Image.fromarray(rgb), roughness=Image.fromarray(gray),
                                        metallic=Image.fromarray(gray), device=DEV)
        assert _cabi_launches() >= l0 + 4
        assert mat.albedo.is_cuda and torch.equal(mat.albedo.cpu(), O.ingest_image(rgb))
        assert torch.equal(mat.normal.cpu(), O.ingest_image(rgb, is_normal=True))
        assert torch.equal(mat.roughness.cpu(), O.ingest_image(gray))


def test_fused_adam_matches_torch_adam_and_fit_converges():
    """pbr_adam_step == torch.optim.Adam + projection over several steps (rel 1e-5; the moment updates are the same
    formulas, the compiler may contract multiply-adds), then a short inverse-rendering fit must reduce the loss."""
    from oracle import pbr_oracle as O
    from pypbr_b200.fit import FusedAdam, fit_step
    from pypbr_b200.materials import BasecolorMetallicMaterial
    from pypbr_b200.models import CookTorranceBRDF

    g = torch.Generator().manual_seed(21)
    B, H, W = 2, 19, 36
    params = dict(albedo=torch.rand(B, 3, H, W, generator=g), roughness=torch.rand(B, 1, H, W, generator=g),
                  metallic=torch.rand(B, 1, H, W, generator=g),
                  normal=torch.nn.functional.normalize(torch.randn(B, 3, H, W, generator=g) * 0.3 + torch.tensor([0, 0, 1.0]).view(1, 3, 1, 1), dim=1))
    steps = [{k: torch.randn(v.shape, generator=g) * (10.0 ** (i - 2)) for k, v in params.items()} for i in range(5)]
    ref = O.adam_fit_steps(params, steps, lr=0.05)
    dev_params = {k: v.clone().to(DEV) for k, v in params.items()}
    opt = FusedAdam(dev_params, lr=0.05)
    for gr in steps:
        opt.step({k: v.to(DEV) for k, v in gr.items()})
    for k in params:
        a, b = dev_params[k].cpu(), ref[k]
        assert bool(((a - b).abs() <= 1e-5 * b.abs() + 1e-6).all()), (k, float((a - b).abs().max()))
    assert float(dev_params["albedo"].min()) >= 0.0 and float(dev_params["albedo"].max()) <= 1.0
    nrm = dev_params["normal"].norm(dim=1)
    assert torch.allclose(nrm, torch.ones_like(nrm), atol=1e-6)
    # grad_scale is applied to the gradient
    p1 = {k: v.clone().to(DEV) for k, v in params.items()}
    p2 = {k: v.clone().to(DEV) for k, v in params.items()}
    FusedAdam(p1, lr=0.01).step({k: (v * 4).to(DEV) for k, v in steps[2].items()})
    FusedAdam(p2, lr=0.01).step({k: v.to(DEV) for k, v in steps[2].items()}, grad_scale=4.0)
    for k in p1:
        assert torch.allclose(p1[k], p2[k], rtol=1e-6, atol=1e-7)

    # a short fit: render a ground-truth batch under 3 lights, start from a perturbed copy, the loss must fall
    maps, lights, inten, g2 = _random_case(41, 2, 32, 48, 3)
    view = torch.tensor([0.0, 0.0, 1.0])
    gt, _ = _material(maps, dict(light_type="point"))
    with torch.no_grad():
        target = CookTorranceBRDF("point", multi_light="per_light")(gt, view, lights, inten, 1.0)
    start = {k: v.clone() for k, v in maps.items()}
    start["albedo"] = (start["albedo"] * 0.6 + 0.2)
    start["roughness"] = (start["roughness"] * 0.5 + 0.3)
    pred = BasecolorMetallicMaterial(albedo_is_srgb=True, device=DEV)
    leaves = {}
    for k, v in start.items():
        leaves[k] = v.contiguous().to(DEV)
        pred._maps[k] = leaves[k]
    opt = FusedAdam(leaves, lr=0.02)
    losses = []
    scratch = {}
    for _ in range(40):
        buf = fit_step(pred, opt, target, view, lights, inten, "point", 1.0, scratch=scratch)
        losses.append(float(buf[0]) / target.numel())
    assert losses[-1] < 0.25 * losses[0], losses[::8]


@pytest.mark.parametrize("L,wf,normal,project", [(1, "metallic", True, True), (3, "metallic", True, True),
                                                 (6, "specular", True, True), (3, "metallic", False, False)])
def test_one_launch_fit_step_equals_loss_kernel_plus_adam_kernel(L, wf, normal, project):
    """pbr_ct_fit_step (render + MSE + backward + Adam + projection in one launch, gradients in registers) must walk
    the same trajectory as pbr_ct_loss_fwd_bwd followed by pbr_adam_step: the same update on the same fp32 gradient
    values; the epilogue evaluates it with the MUFU reciprocal / rsqrt (relative error ~2e-7 per step) and the two
    instantiations contract multiply-adds differently, hence 1e-5 relative to |x| (+ 4e-6 mean|x|) per step (parity proper is against the oracle: test_c5_shape_*)."""
    from pypbr_b200.fit import FusedAdam, fit_step
    from pypbr_b200.materials import BasecolorMetallicMaterial, DiffuseSpecularMaterial
    from pypbr_b200.models import CookTorranceBRDF

    maps, lights, inten, _ = _random_case(77 + L, 3, 27, 44, L, workflow=wf)   # ragged width: scalar tail path
    if not normal:
        maps.pop("normal")
    view = torch.tensor([0.1, -0.05, 1.0])
    gt, _ = _material(maps, dict(light_type="point"))
    with torch.no_grad():
        target = CookTorranceBRDF("point", multi_light="per_light")(gt, view, lights, inten, 1.0)
    cls = BasecolorMetallicMaterial if wf == "metallic" else DiffuseSpecularMaterial
    from pypbr_b200.models import cooktorrance as ct

    def make():
        pred = cls(albedo_is_srgb=True, device=DEV)
        leaves = {}
        for k, v in maps.items():
            t = torch.from_numpy(v) if isinstance(v, np.ndarray) else v
            leaves[k] = (t * 0.7 + 0.1 if k != "normal" else t).contiguous().to(DEV).clone()
            pred._maps[k] = leaves[k]
        if not normal:
            pred._maps["normal"] = None
        return pred, leaves, FusedAdam(leaves, lr=0.03, project=None if project else {k: None for k in leaves})

    (pa, la, oa), (pb, lb, ob) = make(), make()
    losses = []
    # the one-launch step always runs the generic kernels: pin the two-kernel step to them as well, so both see the
    # same gradient code (streamed == generic is test_streamed_kernels_equal_generic_kernels' business)
    ct.FORCE_GENERIC = True
    try:
        for _ in range(4):
            # same state before every step (an Adam step is lr-sized whatever the gradient, so two trajectories drift
            # apart on noise-level gradients; what is compared here is ONE step from identical state)
            for k in la:
                lb[k].copy_(la[k]); ob.state[k][0].copy_(oa.state[k][0]); ob.state[k][1].copy_(oa.state[k][1])
            ob.step_count = oa.step_count
            buf_a = fit_step(pa, oa, target, view, lights, inten, "point", 1.0, fused=False)
            buf_b = fit_step(pb, ob, target, view, lights, inten, "point", 1.0, fused=True)
            losses.append(float(buf_a[0]))
            assert abs(float(buf_a[0]) - float(buf_b[0])) <= 1e-6 * abs(float(buf_a[0]))
            for k in la:
                for a, b, what in ((la[k], lb[k], "param"), (oa.state[k][0], ob.state[k][0], "exp_avg"),
                                   (oa.state[k][1], ob.state[k][1], "exp_avg_sq")):
                    err = (a - b).abs()
                    # (5 < L <= 8: the loss kernel caches attenuation and (1-h.v)^5 per light, the fit kernel - whose Adam staging
                    # needs the shared memory - recomputes them from the plane position: a few ulp of the gradient)
                    # (a gradient is a sum over the lights with cancellation: an ulp of one light's attenuation shows up
                    # relative to the mean magnitude, not to the element)
                    tol = 1e-5 * a.abs() + 2e-5 * a.abs().mean() + (2e-7 if what == "param" else 0.0)
                    ok = err <= tol
                    if what == "param":
                        # an Adam step is lr-sized whatever the gradient: where the gradient is at noise level (|m| ~ eps) a few
                        # ulp of it move the step by a visible fraction of lr, so the parameter is compared where it is defined
                        m1 = oa.state[k][0].abs()
                        ok = ok | (m1 <= 1e-4 * m1.mean())
                    assert bool(ok.all()), (k, what, float((err / tol)[~ok].max()))
    finally:
        ct.FORCE_GENERIC = False
    assert losses[-1] < losses[0]
    if project:
        assert float(lb["albedo"].min()) >= 0.0 and float(lb["albedo"].max()) <= 1.0
        if normal:
            nrm = lb["normal"].norm(dim=1)
            assert torch.allclose(nrm, torch.ones_like(nrm), atol=1e-6)


def test_fit_step_rejects_what_the_one_launch_path_cannot_update_in_place():
    from pypbr_b200.fit import FusedAdam, fused_fit_step
    from pypbr_b200.models import CookTorranceBRDF

    maps, lights, inten, _ = _random_case(5, 2, 16, 16, 2)
    view = torch.tensor([0.0, 0.0, 1.0])
    mat, leaves = _material(maps, dict(light_type="point"))
    with torch.no_grad():
        target = CookTorranceBRDF("point", multi_light="per_light")(mat, view, lights, inten, 1.0)
    with pytest.raises(ValueError):   # the optimiser misses maps the kernel updates
        fused_fit_step(mat, FusedAdam({"albedo": leaves["albedo"]}), target, view, lights, inten)
    with pytest.raises(ValueError):   # mixed projection
        fused_fit_step(mat, FusedAdam(leaves, project={"albedo": None}), target, view, lights, inten)
    other = {k: v.clone() for k, v in leaves.items()}
    with pytest.raises(ValueError):   # not the material's own tensors
        fused_fit_step(mat, FusedAdam(other), target, view, lights, inten)


def _shared_tol(want):
    want = np.asarray(want, np.float64)
    return 1e-4 * np.abs(want) + 1e-4 * np.abs(want).mean()


@pytest.mark.parametrize("tag", ["point_metal", "dir_spec"])
def test_light_and_view_gradients_match_reference_fixture(tag):
    """view_dir / light position | direction / intensity gradients through autograd of the public API against the
    REFERENCE's own autograd results (tests/golden/geomgrad_shared_params.npz, fp64 run as the arbiter: these are sums
    over the whole image)."""
    import json
    import os

    from conftest import GOLDEN

    z = np.load(os.path.join(GOLDEN, "geomgrad_shared_params.npz"))
    p = json.loads(str(z[tag + "_params"]))
    names = ("albedo", "normal", "roughness", "metallic" if p["workflow"] == "metallic" else "specular")
    mat, _ = _material({k: z[f"{tag}_in_{k}"] for k in names}, p)
    for where in ("cpu", "cuda"):   # CPU leaves (the reference's usual call) and device-resident parameters
        view, light, inten = (torch.tensor(p[k], device=where, requires_grad=True) for k in ("view", "light", "intensity"))
        out = _brdf(p, False)(mat, view, light, inten, p["light_size"], True)
        out.backward(torch.from_numpy(z[tag + "_grad_out"]).to(DEV))
        for t, key in ((view, "view"), (light, "light"), (inten, "intensity")):
            assert t.grad is not None and t.grad.device.type == where and t.grad.shape == t.shape
            want = z[f"{tag}_g64_{key}"]
            assert np.all(np.abs(t.grad.cpu().numpy() - want) <= _shared_tol(want)), (key, t.grad, want)


@pytest.mark.parametrize("light_type,B,L,per_light,wf", [("point", 3, 4, False, "metallic"), ("point", 2, 3, True, "specular"),
                                                         ("directional", 2, 3, False, "metallic"), ("point", None, 1, False, "metallic")])
def test_light_and_view_gradients_against_oracle(light_type, B, L, per_light, wf):
    """Batched / multi-light: the gradients of the shared parameters are sums over materials (and lights, for the
    view) of the reference's single-call gradients; map gradients must not change when they are requested."""
    from oracle import pbr_oracle as O

    maps, lights, inten, g = _random_case(300 + L, B, 26, 38, L, workflow=wf)
    if L == 1:
        lights, inten = lights.view(3), inten.view(3)
    view = torch.tensor([0.12, -0.08, 0.95])
    p = dict(light_type=light_type)
    size = 1.0 if light_type == "point" else None
    v64, l64, i64 = (t.double().clone().requires_grad_(True) for t in (view, lights, inten))
    ref = O.render({k: t.double() for k, t in maps.items()}, v64, l64, i64, size, light_type, accumulate=not per_light)
    go = torch.rand(ref.shape, generator=g)
    ref.backward(go.double())

    mat, leaves = _material(maps, p, requires_grad=True)
    vd, ld, idv = (t.clone().requires_grad_(True) for t in (view, lights, inten))
    out = _brdf(p, per_light)(mat, vd, ld, idv, size, True)
    out.backward(go.to(DEV))
    for t, want, key in ((vd, v64.grad, "view"), (ld, l64.grad, "lights"), (idv, i64.grad, "intensity")):
        want = want.numpy()
        assert np.all(np.abs(t.grad.numpy() - want) <= _shared_tol(want)), (key, t.grad, want)
    # same map gradients as the kernels that do not compute the shared-parameter gradients
    mat2, leaves2 = _material(maps, p, requires_grad=True)
    _brdf(p, per_light)(mat2, view, lights, inten, size, True).backward(go.to(DEV))
    for k in leaves:
        a, b = leaves[k].grad, leaves2[k].grad
        assert bool(((a - b).abs() <= 2e-6 * b.abs() + 2e-6 * b.abs().mean()).all()), k


def test_fused_loss_step_delivers_shared_parameter_gradients():
    """[loss, d_intensity, d_lights, d_view]: the 1 + 6L + 3 float buffer a sharded fit all-reduces."""
    from oracle import pbr_oracle as O
    from pypbr_b200.fit import fused_loss_step

    L = 3
    maps, lights, inten, g = _random_case(55, 2, 20, 32, L)
    view = torch.tensor([0.05, 0.1, 1.0])
    v64, l64, i64 = (t.double().clone().requires_grad_(True) for t in (view, lights, inten))
    ref = O.render({k: t.double() for k, t in maps.items()}, v64, l64, i64, 1.0, "point", accumulate=False)
    target = torch.rand(ref.shape, generator=g)
    loss = ((ref - target.double()) ** 2).mean()
    loss.backward()
    mat, _ = _material(maps, dict(light_type="point"))
    buf, _grads = fused_loss_step(mat, target.to(DEV), view, lights, inten, "point", 1.0, want_intensity_grad=True,
                                  want_geometry_grad=True)
    buf = buf.cpu().numpy().astype(np.float64)
    assert buf.shape == (1 + 6 * L + 3,)
    assert abs(buf[0] / target.numel() - float(loss)) <= 1e-5 * float(loss)
    for got, want, key in ((buf[1:1 + 3 * L], i64.grad, "intensity"), (buf[1 + 3 * L:1 + 6 * L], l64.grad, "lights"),
                           (buf[1 + 6 * L:], v64.grad, "view")):
        want = want.numpy().reshape(-1)
        assert np.all(np.abs(got - want) <= _shared_tol(want)), (key, got, want)


def test_normal_from_height_and_rotate_match_reference_ops():
    """compute_normal_from_height (utils/functions.py:123-177) and rotate / rotate_normals
    (base.py:539-603, utils/functions.py:69-108) against the reference's sequence of torch ops, restated here."""
    import math
    import torch.nn.functional as F
    from torchvision.transforms import functional as TF

    from pypbr_b200.materials import BasecolorMetallicMaterial
    from pypbr_b200.utils import NormalConvention, compute_normal_from_height, rotate_normals

    g = torch.Generator().manual_seed(8)
    for shape in ((1, 37, 53), (1, 64, 64), (1, 1, 5), (1, 6, 1)):
        h = torch.rand(shape, generator=g)
        for conv in (NormalConvention.OPENGL, NormalConvention.DIRECTX):
            for scale in (1.0, 7.5):
                gx = (F.pad(h, (1, 0, 0, 0))[:, :, :-1] - F.pad(h, (0, 1, 0, 0))[:, :, 1:]) * scale
                gy = (F.pad(h, (0, 0, 1, 0))[:, :-1, :] - F.pad(h, (0, 0, 0, 1))[:, 1:, :]) * scale
                want = F.normalize(torch.cat([-gx, -gy if conv == NormalConvention.OPENGL else gy, torch.ones_like(h)], dim=0), dim=0)
                got = compute_normal_from_height(h.to(DEV), scale, conv).cpu()
                assert got.shape == want.shape
                assert bool(((got - want).abs() <= 1.2e-7 * want.abs() + 1e-9).all()), (shape, float((got - want).abs().max()))
    hb = torch.rand(3, 1, 20, 28, generator=g)   # batched = per-material calls
    nb = compute_normal_from_height(hb.to(DEV), 2.0)
    for b in range(3):
        assert torch.equal(nb[b], compute_normal_from_height(hb[b].to(DEV), 2.0))
    assert torch.equal(compute_normal_from_height(hb[0, 0].to(DEV)), compute_normal_from_height(hb[0].to(DEV)))
    with pytest.raises(ValueError):
        compute_normal_from_height(None)

    n = F.normalize(torch.randn(3, 33, 47, generator=g) * torch.tensor([0.4, 0.4, 0.1]).view(3, 1, 1) + torch.tensor([0.0, 0.0, 1.0]).view(3, 1, 1), dim=0)
    for angle in (30.0, -75.0, 180.0):
        th = math.radians(angle)
        R = torch.tensor([[math.cos(th), -math.sin(th)], [math.sin(th), math.cos(th)]])
        xy = torch.stack([n[0].reshape(-1), n[1].reshape(-1)], dim=1) @ R.T
        want = F.normalize(torch.stack([xy[:, 0], xy[:, 1], n[2].reshape(-1)], dim=1), dim=1).T.reshape(3, 33, 47)
        t = n.clone().to(DEV)
        got = rotate_normals(t, angle)
        assert got.data_ptr() == t.data_ptr()   # in place, like the reference
        assert bool(((got.cpu() - want).abs() <= 1e-6 * want.abs() + 2e-7).all()), float((got.cpu() - want).abs().max())

    # material.rotate: resampling through the same library calls as the reference + the vector rotation
    maps, _, _, _ = _random_case(12, None, 40, 56, 1)
    mat, _ = _material(maps, dict(light_type="point"))
    mat.rotate(33.0, expand=False, padding_mode="circular")
    ref = {}
    for k, v in maps.items():
        pad = math.ceil(math.sqrt(40**2 + 56**2)) - 40
        r = TF.center_crop(TF.rotate(F.pad(v.to(DEV), (pad, pad, pad, pad), "circular"), 33.0, expand=True), (40, 56)).contiguous()
        ref[k] = r
    for k in ("albedo", "roughness", "metallic"):
        assert torch.equal(mat._maps[k], ref[k]), k
    th = math.radians(33.0)
    rn = ref["normal"].cpu()
    want = F.normalize(torch.stack([rn[0] * math.cos(th) - rn[1] * math.sin(th), rn[0] * math.sin(th) + rn[1] * math.cos(th), rn[2]]), dim=0)
    assert bool(((mat._maps["normal"].cpu() - want).abs() <= 1e-6 * want.abs() + 2e-7).all())
    m2 = BasecolorMetallicMaterial(albedo_is_srgb=True, device=DEV)
    m2._maps["height"] = torch.rand(1, 16, 16, generator=g).to(DEV)
    m2.compute_normal_from_height(3.0)
    assert m2._maps["normal"].shape == (3, 16, 16)


def test_height_normal_round_trip_matches_reference_fixture():
    """tests/golden/height_normal_48x72.npz: compute_normal_from_height and compute_height_from_normal of the REFERENCE
    (utils/functions.py:123-323), both conventions.  The divergence kernel is also checked bit for bit against the
    reference's op sequence restated with torch (replicate-padded forward differences)."""
    import os

    import torch.nn.functional as F

    from conftest import GOLDEN
    from pypbr_b200 import _cabi
    from pypbr_b200.materials import BasecolorMetallicMaterial
    from pypbr_b200.utils import NormalConvention, compute_height_from_normal, compute_normal_from_height
    from pypbr_b200.utils.functions import _normal_op

    z = np.load(os.path.join(GOLDEN, "height_normal_48x72.npz"))
    height = torch.from_numpy(z["height"]).to(DEV)
    for conv, tag in ((NormalConvention.OPENGL, "gl"), (NormalConvention.DIRECTX, "dx")):
        n = compute_normal_from_height(height, 4.0, conv)
        want_n = torch.from_numpy(z[f"normal_{tag}"])
        assert bool(((n.cpu() - want_n).abs() <= 1.2e-7 * want_n.abs() + 1e-9).all())
        nref = want_n.to(DEV)
        # divergence: bit-exact against the same chain of individually rounded ops
        nz = want_n[2] + 1e-8
        gx = (-want_n[0] / nz) * 1.0
        gy = ((-want_n[1] if conv == NormalConvention.OPENGL else want_n[1]) / nz) * 1.0
        gxp = F.pad(gx[None, None], (0, 1, 0, 0), mode="replicate")[0, 0]
        gyp = F.pad(gy[None, None], (0, 0, 0, 1), mode="replicate")[0, 0]
        want_div = (gxp[:, 1:] - gxp[:, :-1]) + (gyp[1:, :] - gyp[:-1, :])
        div = torch.empty(1, *want_div.shape, device=DEV)
        _normal_op(nref, div, _cabi.NORMAL_OP_DIVERGENCE, scale=1.0, flip_y=int(conv == NormalConvention.DIRECTX))
        assert torch.equal(div[0].cpu(), want_div)
        # the whole reconstruction (cuFFT vs the reference's CPU FFT): heights are normalised to [0, 1]
        h = compute_height_from_normal(nref, 1.0, conv)
        want_h = torch.from_numpy(z[f"height_back_{tag}"])
        assert h.shape == want_h.shape == (1, 48, 72)
        assert float((h.cpu() - want_h).abs().max()) <= 2e-5
    m = BasecolorMetallicMaterial(albedo_is_srgb=True, device=DEV)
    m._maps["normal"] = torch.from_numpy(z["normal_gl"]).to(DEV)
    m.compute_height_from_normal()
    assert float((m._maps["height"].cpu() - torch.from_numpy(z["height_back_gl"])).abs().max()) <= 2e-5
    with pytest.raises(ValueError):
        compute_height_from_normal(None)
    with pytest.raises(ValueError):
        compute_height_from_normal(torch.zeros(2, 4, 4, device=DEV))


@pytest.mark.parametrize("B,L,wf,light_type", [(3, 5, "metallic", "point"), (18, 3, "specular", "point"), (2, 4, "metallic", "directional")])
def test_accumulate_backward_one_pass_from_saved_output_and_two_pass(B, L, wf, light_type, ct_path):
    """Accumulate mode, L > 1: autograd hands the forward output to the backward (PbrCtGrads.fwd_out), which then skips the
    recomputation of the summed image; without it (plain C-ABI callers) the kernel runs two passes.  Both against the oracle."""
    from oracle import pbr_oracle as O
    from pypbr_b200.models import cooktorrance as ct

    maps, lights, inten, g = _random_case(900 + L, B, 22, 36, L, workflow=wf)
    inten = inten * 40.0   # some texels saturate: the clamp of the sum gates
    view = torch.tensor([0.0, 0.1, 1.0])
    p = dict(light_type=light_type)
    size = 1.0 if light_type == "point" else None
    leaves = {k: t.clone().requires_grad_(True) for k, t in maps.items()}
    ref = O.render(leaves, view, lights, inten, size, light_type, accumulate=True)
    go = torch.rand(ref.shape, generator=g)
    ref.backward(go)
    assert float((ref.detach() >= 0.9999999).float().mean()) > 0.01   # the case does exercise the saturated gate (encode(1) = 0.99999994)
    got = {}
    for two_pass in (False, True):
        ct.NO_SAVED_OUT = two_pass
        try:
            mat, lv = _material(maps, p, requires_grad=True)
            out = _brdf(p, False)(mat, view, lights, inten, size, True)
            out.backward(go.to(DEV))
        finally:
            ct.NO_SAVED_OUT = False
        ratio, ok = fwd_ok(out.detach().cpu().numpy(), ref.detach().numpy())
        assert ok, ratio
        for k in maps:
            ratio, ok = grad_ok(lv[k].grad.cpu().numpy(), leaves[k].grad.numpy())
            assert ok, (k, two_pass, ratio)
        got[two_pass] = {k: lv[k].grad for k in maps}
    for k in maps:
        a, b = got[False][k], got[True][k]
        bad = (a - b).abs() > 4e-6 * (b.abs() + b.abs().mean())
        assert float(bad.float().mean()) < 1e-3, k


# ---------------------------------------------------------------------------------------------- parity AT the benchmarked shapes
def _oracle_case(maps_b, view, lights, inten, acc, go=None):
    """Oracle (CPU, the reference's op sequence) on a (sub)batch: output and, with `go`, gradients."""
    from oracle import pbr_oracle as O

    leaves = {k: v.clone().requires_grad_(go is not None) for k, v in maps_b.items()}
    out = O.render(leaves, view, lights, inten, 1.0, "point", accumulate=acc)
    if go is not None:
        out.backward(go)
    return out.detach(), {k: v.grad for k, v in leaves.items()} if go is not None else None


def _bench_like_maps(B, H, W, seed, rough_lo=0.2, workflow="metallic", normal=True):
    """bench.synth_maps (SURVEY.md 8d distribution), generated on the host so the oracle sees the very same bits."""
    maps, _l, _i, g = _random_case(seed, B, H, W, 1, workflow, rough_lo=rough_lo, normal=normal)
    return maps, g


@pytest.mark.parametrize("case", ["metallic_point", "specular_directional", "no_normal", "linear_albedo"])
def test_c2_shape_auto_path_against_oracle(case):
    """BASELINE.json configs[1]'s shape, 1024 x 1024, one light, with B = 17 so the 16-material walk of a CTA ends and a
    second chunk starts: the AUTOMATIC kernel choice (the streamed TMA-fed kernels, per-warp pipelines, multi-CTA rows)
    against the oracle on two materials of the batch - the first, and the one in the second chunk."""
    from oracle import pbr_oracle as O

    B, H, W = 17, 1024, 1024
    wf = "specular" if case == "specular_directional" else "metallic"
    maps, g = _bench_like_maps(B, H, W, 4242, workflow=wf, normal=case != "no_normal")
    light_type = "directional" if case == "specular_directional" else "point"
    lights = torch.tensor([0.2, -0.3, 0.9]) if light_type == "directional" else torch.tensor([0.1, 0.1, 1.0])
    inten, view = torch.tensor([1.0, 0.9, 0.8]), torch.tensor([0.0, 0.05, 1.0])
    p = dict(light_type=light_type, albedo_is_srgb=case != "linear_albedo")
    go = torch.rand(B, 3, H, W, generator=g)
    mat, leaves = _material(maps, p, requires_grad=True)
    from pypbr_b200 import _cabi
    l0 = _cabi.launch_count()
    out = _brdf(p, False)(mat, view, lights, inten, 1.0 if light_type == "point" else None)
    out.backward(go.to(DEV))
    assert _cabi.launch_count() == l0 + 2
    for b in (0, 16):
        leaves_ref = {k: v[b].clone().requires_grad_(True) for k, v in maps.items()}
        ref = O.render(leaves_ref, view, lights, inten, 1.0 if light_type == "point" else None, light_type,
                       albedo_is_srgb=p["albedo_is_srgb"])
        ref.backward(go[b])
        ratio, ok = fwd_ok(out[b].detach().cpu().numpy(), ref.detach().numpy())
        assert ok, (case, b, ratio)
        for k in maps:
            ratio, ok = grad_ok(leaves[k].grad[b].cpu().numpy(), leaves_ref[k].grad.numpy())
            assert ok, (case, b, k, ratio)


def test_c2_shape_roughness_stress_distribution():
    """Roughness U[0,1] at 1024 x 1024 (SURVEY.md 8c "stress distribution"): the GGX denominator amplifies an ulp of N.H by
    2/dn, the reference's own fp32 run is then far from its fp64 run on a tail of texels.  Report the fraction inside
    the plain tolerance and require |y - y64| <= max(2 |y32 - y64|, tol) everywhere."""
    from oracle import pbr_oracle as O

    B, H, W = 2, 1024, 1024
    maps, g = _bench_like_maps(B, H, W, 777, rough_lo=0.0)
    lights, inten, view = torch.tensor([0.1, 0.1, 1.0]), torch.ones(3), torch.tensor([0.0, 0.0, 1.0])
    mat, _ = _material(maps, dict(light_type="point"))
    out = _brdf(dict(light_type="point"), False)(mat, view, lights, inten, 1.0).cpu().numpy().astype(np.float64)
    m0 = {k: v[1] for k, v in maps.items()}
    y32 = O.render(m0, view, lights, inten, 1.0, "point").numpy().astype(np.float64)
    y64 = O.render({k: v.double() for k, v in m0.items()}, view.double(), lights.double(), inten.double(), 1.0, "point").numpy()
    tol = 1e-5 * np.abs(y32) + 1e-6
    plain = np.abs(out[1] - y32) <= tol
    arb = np.abs(out[1] - y64) <= np.maximum(2 * np.abs(y32 - y64), tol)
    print(f"stress distribution: fraction inside rel 1e-5 of the fp32 reference = {plain.mean():.6f}")
    assert plain.mean() > 0.999
    assert bool((plain | arb).all()), float((~(plain | arb)).mean())


def test_c3_shape_sixteen_lights_accumulate_against_oracle():
    """BASELINE.json configs[2]'s shape: 2048 x 2048, 16 point lights, accumulate mode, forward + backward (the cached
    light-geometry kernels, one-pass backward from the saved output) - one material of a batch of 2 against the oracle
    (16 reference calls + autograd, ~30 s of host time)."""
    from oracle import pbr_oracle as O

    B, H, W, L = 2, 2048, 2048, 16
    maps, g = _bench_like_maps(B, H, W, 31337)
    _m, lights, inten, _g = _random_case(1, None, 4, 4, L)
    view = torch.tensor([0.0, 0.0, 1.0])
    go = torch.rand(B, 3, H, W, generator=g)
    p = dict(light_type="point")
    mat, leaves = _material(maps, p, requires_grad=True)
    out = _brdf(p, False)(mat, view, lights, inten, 1.0)
    out.backward(go.to(DEV))
    b = 1
    # The oracle, light by light, so that only ONE reference call's autograd graph is alive at a time (16 of them at
    # 2048 x 2048 would hold ~20 GB): S = sum_l call_l without grad, dLoss/dS through encode(clamp(S)) by autograd on S
    # alone, then every call_l is back-propagated with that same dLoss/dS - the chain rule of out = encode(clamp(sum_l call_l)).
    mb = {k: v[b] for k, v in maps.items()}
    with torch.no_grad():
        S = sum(O.shade_linear(mb, view, lights[l], inten[l], 1.0, "point") for l in range(L))
    S.requires_grad_(True)
    ref = O.linear_to_srgb(torch.clamp(S, 0.0, 1.0))
    ref.backward(go[b])
    leaves_ref = {k: v.clone().requires_grad_(True) for k, v in mb.items()}
    for l in range(L):
        O.shade_linear(leaves_ref, view, lights[l], inten[l], 1.0, "point").backward(S.grad)
    ratio, ok = fwd_ok(out[b].detach().cpu().numpy(), ref.detach().numpy())
    assert ok, ratio
    for k in maps:
        ratio, ok = grad_ok(leaves[k].grad[b].cpu().numpy(), leaves_ref[k].grad.numpy())
        assert ok, (k, ratio)


def test_c5_shape_eight_lights_fused_loss_and_one_launch_fit_step_against_oracle():
    """BASELINE.json configs[4]'s per-GPU shape: 512 x 512, 8 lights, per-light targets.  Two materials through
    fused_loss_step (pbr_ct_loss_fwd_bwd) against the oracle's MSE + autograd, then the same step through pbr_ct_fit_step
    against torch.optim.Adam + projection applied to the oracle's gradients."""
    from oracle import pbr_oracle as O
    from pypbr_b200.fit import FusedAdam, fit_step, fused_loss_step

    B, H, W, L = 2, 512, 512, 8
    maps, g = _bench_like_maps(B, H, W, 555)
    _m, lights, inten, _g = _random_case(1, None, 4, 4, L)
    inten = inten * L     # per-light renders, full intensity
    view = torch.tensor([0.0, 0.0, 1.0])
    target = torch.rand(B, L, 3, H, W, generator=g)
    leaves_ref = {k: v.clone().requires_grad_(True) for k, v in maps.items()}
    ref = O.render(leaves_ref, view, lights, inten, 1.0, "point", accumulate=False)
    loss_ref = ((ref - target) ** 2).mean()
    loss_ref.backward()
    p = dict(light_type="point")
    mat, _ = _material(maps, p)
    buf, grads = fused_loss_step(mat, target.to(DEV), view, lights, inten, "point", 1.0)
    assert abs(float(buf[0]) / target.numel() - float(loss_ref.detach())) <= 2e-6 * float(loss_ref.detach())
    for k in maps:
        ratio, ok = grad_ok(grads[k].cpu().numpy(), leaves_ref[k].grad.numpy())
        assert ok, (k, ratio)
    # one launch: the same gradients feed Adam + projection in the epilogue
    want = O.adam_fit_steps(maps, [{k: leaves_ref[k].grad for k in maps}], lr=0.01)
    mat2, lv2 = _material(maps, p)
    opt = FusedAdam({k: mat2._maps[k] for k in maps}, lr=0.01)
    pend = fit_step(mat2, opt, target.to(DEV), view, lights, inten, "point", 1.0, fused=True, async_loss=True)
    assert abs(pend.item() - float(loss_ref.detach())) <= 2e-6 * float(loss_ref.detach())
    for k in maps:
        a, b = mat2._maps[k].cpu(), want[k]
        # an Adam step is lr-sized whatever the gradient: where |g| is at noise level its sign is not defined to 1e-4
        big = leaves_ref[k].grad.abs() > 1e-3 * leaves_ref[k].grad.abs().mean()
        assert bool(((a - b).abs()[big] <= 2e-5 * b.abs()[big] + 2e-6).all()), (k, float((a - b).abs()[big].max()))


def test_saturated_accumulate_gate_is_inclusive_at_one():
    """torch.clamp passes the gradient AT its bounds.  A texel whose summed light is exactly 1.0 (one saturating light plus
    a zero-intensity light) and texels at exactly 0: the one-pass backward (which reads the gate off the saved output,
    where 1.0 and > 1.0 encode alike) recomputes the sum for exactly those texels and must agree with the oracle - the
    zero-intensity light's intensity gradient is where it shows."""
    from oracle import pbr_oracle as O
    from pypbr_b200.models import cooktorrance as ct

    maps, _l, _i, g = _random_case(99, 2, 16, 24, 1)
    lights = torch.tensor([[0.1, 0.1, 1.0], [-0.2, 0.3, 0.9]])
    inten0 = torch.tensor([[400.0, 400.0, 400.0], [0.0, 0.0, 0.0]])
    view = torch.tensor([0.0, 0.0, 1.0])
    ri = inten0.clone().requires_grad_(True)
    leaves_ref = {k: v.clone().requires_grad_(True) for k, v in maps.items()}
    ref = O.render(leaves_ref, view, lights, ri, 1.0, "point", accumulate=True)
    go = torch.rand(ref.shape, generator=g)
    ref.backward(go)
    assert float((ref.detach() == ref.detach().max()).float().mean()) > 0.2   # many texels sit exactly at the top
    for two_pass in (False, True):
        ct.NO_SAVED_OUT = two_pass
        try:
            mat, lv = _material(maps, dict(light_type="point"), requires_grad=True)
            di = inten0.clone().to(DEV).requires_grad_(True)
            out = _brdf(dict(light_type="point"), False)(mat, view.to(DEV), lights.to(DEV), di, 1.0)
            out.backward(go.to(DEV))
        finally:
            ct.NO_SAVED_OUT = False
        assert fwd_ok(out.detach().cpu().numpy(), ref.detach().numpy())[1]
        want = ri.grad.numpy()
        assert np.all(np.abs(di.grad.cpu().numpy() - want) <= 1e-4 * np.abs(want) + 1e-4 * np.abs(want).mean()), (two_pass, di.grad, want)
        for k in maps:
            assert grad_ok(lv[k].grad.cpu().numpy(), leaves_ref[k].grad.numpy())[1], (k, two_pass)


# ---------------------------------------------------------------------------------------------- SURVEY 8f rank 4 remainder
def test_adjust_normal_strength_and_packing_on_cuda_maps():
    """adjust_normal_strength (base.py:689-706), as_tensor / from_tensor (base.py:319-487) on CUDA maps against the
    reference's op sequence restated on the host."""
    import torch.nn.functional as F

    from pypbr_b200.materials import BasecolorMetallicMaterial

    maps, _l, _i, g = _random_case(64, None, 33, 52, 1)
    mat, _ = _material(maps, dict(light_type="point"))
    shared = mat._maps["normal"]
    l0 = _cabi_launches()
    mat.adjust_normal_strength(2.5)
    assert _cabi_launches() == l0 + 1
    n = maps["normal"].clone()
    n[:2] *= 2.5
    want = F.normalize(n, dim=0)
    got = mat._maps["normal"].cpu()
    assert bool(((got - want).abs() <= 3e-7 * want.abs() + 1e-8).all()), float((got - want).abs().max())
    assert torch.equal(shared.cpu(), n)   # the reference scales the ORIGINAL tensor's x / y in place; kept
    # packing: channel-stacked tensor and back
    packed = mat.as_tensor(names=["albedo", ("normal", 2), "roughness", "metallic"], normalize=True)
    assert packed.is_cuda and packed.shape == (3 + 2 + 1 + 1, 33, 52)
    exp = torch.cat([(maps["albedo"] - 0.5) / 0.5, want[:2], (maps["roughness"] - 0.5) / 0.5, (maps["metallic"] - 0.5) / 0.5], dim=0)
    assert bool(((packed.cpu() - exp).abs() <= 3e-7 * exp.abs() + 1e-7).all())
    full = mat.as_tensor()
    assert full.shape[0] == 8
    order = [(k, t.shape[0]) for k, t in mat._maps.items()]   # as_tensor() without names stacks in the registry's order
    back = BasecolorMetallicMaterial.from_tensor(full, names=order, device=DEV)
    for k in ("albedo", "roughness", "metallic"):
        assert torch.equal(back._maps[k], mat._maps[k]) and back._maps[k].data_ptr() != mat._maps[k].data_ptr()
    rgb = torch.rand(2 + 3, 20, 28, generator=g)
    m2 = BasecolorMetallicMaterial.from_tensor(rgb.to(DEV), names=[("normal", 2), ("albedo", 3)], device=DEV)
    from oracle import pbr_oracle as O
    assert np.allclose(m2._maps["normal"].cpu().numpy(), O.process_normal_map(rgb[:2]).numpy(), rtol=3e-7, atol=1e-8)
    with pytest.raises(KeyError):
        mat.as_tensor(names=["nope"])
    with pytest.raises(ValueError):
        BasecolorMetallicMaterial.from_tensor(rgb.to(DEV), names=[("albedo", 3)], device=DEV)


def test_blend_on_height_resizes_a_mismatched_height_map():
    """functional.py:183-185: material2's height map is TF.resize'd (bilinear, antialias) to material1's size before the
    sigmoid.  Same library call on the device; the mask is compared with the host evaluation of the same sequence."""
    from torchvision.transforms import functional as TF

    from oracle import pbr_oracle as O
    from pypbr_b200.blending import blend_on_height
    from pypbr_b200.materials import BasecolorMetallicMaterial

    g = torch.Generator().manual_seed(12)
    H, W = 24, 40
    a1, a2 = torch.rand(3, H, W, generator=g), torch.rand(3, H, W, generator=g)
    h1, h2 = torch.rand(1, H, W, generator=g), torch.rand(1, H // 2, W // 2, generator=g)
    m1 = BasecolorMetallicMaterial(albedo=a1.to(DEV), device=DEV, height=h1.to(DEV))
    m2 = BasecolorMetallicMaterial(albedo=a2.to(DEV), device=DEV)
    m2._maps["height"] = h2.to(DEV)
    # the reference blends every shared map with the mask; a height map of another size cannot be blended map-wise there
    # either (the broadcast fails), so the shared maps here are albedo only and height enters through the mask
    m1b = BasecolorMetallicMaterial(albedo=a1.to(DEV), device=DEV)
    h2r = TF.resize(h2, [H, W], antialias=True)
    want_mask = O.sigmoid_mask(h1, h2r, 0.1, 0.0)
    from pypbr_b200.blending.functional import _sigmoid_blend
    blended, mask = _sigmoid_blend(m1b, m2.__class__(albedo=a2.to(DEV), device=DEV), h1.to(DEV), h2.to(DEV), 0.1, 0.0, True)
    assert mask.shape == (1, H, W)
    assert np.allclose(mask.cpu().numpy(), want_mask.numpy(), rtol=2e-5, atol=2e-6)   # CUDA vs CPU antialias resize + sigmoid
    exp = want_mask * a1 + (1 - want_mask) * a2
    assert np.allclose(blended.albedo.cpu().numpy(), exp.numpy(), rtol=2e-5, atol=2e-6)
    with pytest.raises(ValueError):
        blend_on_height(m1b, m2)   # material1 has no height map


def test_device_resident_parameters_need_no_host_staging(monkeypatch):
    """material.to('cuda') (a device without an index) with CUDA view / light / intensity tensors: the parameters are read by the
    kernel from device memory; nothing is copied to the host (the comparison `t.device == device` was never true)."""
    from pypbr_b200 import _cabi
    from pypbr_b200.models import CookTorranceBRDF

    maps, lights, inten, g = _random_case(3, 2, 16, 24, 2)
    mat, _ = _material(maps, dict(light_type="point"), device=torch.device("cpu"))
    mat.to("cuda")
    assert str(mat.device) == "cuda"

    def boom(_values):
        raise AssertionError("light parameters were staged through the host")

    monkeypatch.setattr(_cabi, "host_floats", boom)
    out = CookTorranceBRDF("point")(mat, torch.tensor([0.0, 0.0, 1.0], device="cuda"), lights.to("cuda"), inten.to("cuda"), 1.0)
    from oracle import pbr_oracle as O
    ref = O.render(maps, torch.tensor([0.0, 0.0, 1.0]), lights, inten, 1.0, "point")
    assert fwd_ok(out.cpu().numpy(), ref.numpy())[1]


def test_rendering_loss_module_shapes_and_no_grad():
    """RenderingLoss (06_advanced.rst:73-107): gradients land on the LEAVES' own shapes ((H,W) roughness, a batch-1 map
    broadcast over the batch), no gradient work under torch.no_grad(), light / view gradients when those require grad."""
    from oracle import pbr_oracle as O
    from pypbr_b200.fit import RenderingLoss
    from pypbr_b200.materials import BasecolorMetallicMaterial

    maps, _l, _i, g = _random_case(8, None, 20, 28, 1)
    gt = {k: v.clone() for k, v in _random_case(9, None, 20, 28, 1)[0].items()}
    pred = BasecolorMetallicMaterial(albedo_is_srgb=True, device=DEV)
    lv = {k: v.clone().to(DEV) for k, v in maps.items()}
    lv["roughness"] = lv["roughness"][0]           # (H, W): the reference accepts it through broadcasting
    for k in lv:
        lv[k].requires_grad_(True)
        pred._maps[k] = lv[k]
    gtm, _ = _material(gt, dict(light_type="point"))
    inten = torch.tensor([1.0, 0.9, 0.8], requires_grad=True)
    crit = RenderingLoss(light_type="point", light_intensity=inten)
    loss = crit(pred, gtm)
    loss.backward()
    assert lv["roughness"].grad.shape == (20, 28) and inten.grad is not None and inten.grad.shape == (3,)
    r = {k: v.clone().requires_grad_(True) for k, v in maps.items()}
    ri = inten.detach().clone().requires_grad_(True)
    a = O.render(r, crit.view_dir, crit.light_dir, ri, None, "point")
    b = O.render(gt, crit.view_dir, crit.light_dir, inten.detach(), None, "point")
    lr = torch.nn.MSELoss()(a, b)
    lr.backward()
    assert abs(float(loss) - float(lr.detach())) <= 2e-6 * float(lr.detach()) + 1e-9
    for k in maps:
        want = r[k].grad.numpy().reshape(lv[k].shape)
        assert grad_ok(lv[k].grad.cpu().numpy(), want)[1], k
    wi = ri.grad.numpy()
    assert np.all(np.abs(inten.grad.numpy() - wi) <= 1e-4 * np.abs(wi) + 1e-4 * np.abs(wi).mean())
    before = torch.cuda.memory_allocated()
    with torch.no_grad():
        l2 = crit(pred, gtm)
    assert not l2.requires_grad and abs(float(l2) - float(loss)) <= 1e-6 * float(loss)
    assert torch.cuda.memory_allocated() - before < 64 * 1024   # no gradient buffers (4 maps x 20 x 28 would be far below; guards growth)


# ---------------------------------------------------------------------------------------------- kernel-flavour dispatch
def _dispatch_cases():
    """Deterministic sweep over what decides the kernel flavour ('Kernel selection'): tile-filling widths
    (W % 512 == 0 -> kFast), lights (1 -> streamed; 2..4, 5..8, 9..16 -> the three geometry caches; one material ->
    per-texel geometry), colour flags, a missing normal map, per-light output, which leaves want gradients, row-strided
    views of wider tensors (16-byte aligned: still eligible for the fast flavours) and a batch-broadcast roughness."""
    cases = []
    i = 0
    for W in (512, 1024):
        for L in (1, 3, 6, 9):
            for B in ((None, 2, 17) if L == 1 else (1, 3)):
                i += 1
                cases.append(dict(
                    W=W, H=2 + i % 2, L=L, B=B,
                    wf="specular" if i % 3 == 0 else "metallic",
                    normal=i % 5 != 0,
                    albedo_is_srgb=i % 7 != 0,
                    return_srgb=i % 4 != 1,
                    per_light=(L > 1 and i % 6 == 2),
                    light_type="directional" if i % 8 == 3 else "point",
                    grads="all" if i % 3 != 1 else "some",
                    strided=i % 2 == 0,
                    broadcast_rough=(B not in (None, 1) and i % 4 == 2),
                ))
    return cases


@pytest.mark.parametrize("c", _dispatch_cases(), ids=lambda c: "W{W}_L{L}_B{B}_{wf}_{light_type}".format(**c))
def test_kernel_flavour_dispatch_against_oracle(c):
    """Every automatic kernel choice - kFast or general streamed kernels, plain-case or general multi-light kernels - computes
    the reference's function: forward and the requested gradients against the oracle on tile-filling widths."""
    from oracle import pbr_oracle as O

    B, H, W, L = c["B"], c["H"], c["W"], c["L"]
    maps, lights, inten, g = _random_case(4200 + W + 13 * L + (B or 0), B, H, W, L, c["wf"], normal=c["normal"])
    if c["broadcast_rough"]:
        # one roughness map shared by the whole batch: the reference semantics are the same values for every material
        maps["roughness"] = maps["roughness"][:1].expand_as(maps["roughness"]).clone()
    if c["light_type"] == "directional":
        lights = torch.nn.functional.normalize(lights + torch.tensor([0.0, 0.0, 1.0]), dim=-1)
    view = torch.tensor([0.03, -0.08, 1.0])
    size = 1.0 if c["light_type"] == "point" else None
    flags = dict(albedo_is_srgb=c["albedo_is_srgb"], specular_is_srgb=True, return_srgb=c["return_srgb"])
    want = set(maps) if c["grads"] == "all" else {"albedo", "roughness"}

    leaves_ref = {k: v.clone().requires_grad_(k in want) for k, v in maps.items()}
    ref = O.render(leaves_ref, view, lights, inten, size, c["light_type"], accumulate=not c["per_light"], **flags)
    go = torch.rand(ref.shape, generator=g)
    ref.backward(go)

    p = dict(light_type=c["light_type"], albedo_is_srgb=c["albedo_is_srgb"])
    mat, _ = _material(maps, p)
    leaves = {}
    for k, v in maps.items():
        t = v.to(DEV)
        if c["broadcast_rough"] and k == "roughness":
            t = t[:1].contiguous()           # (1, 1, H, W): batch stride 0 once the host expands it
        elif c["strided"]:
            wide = torch.zeros(*t.shape[:-1], W + 8, device=DEV)
            wide[..., 4:W + 4] = t
            t = wide[..., 4:W + 4]           # row stride W + 8, first texel 16-byte aligned
        t = t.detach().requires_grad_(k in want)
        leaves[k] = t
        mat._maps[k] = t
    out = _brdf(p, c["per_light"])(mat, view, lights, inten, size, c["return_srgb"])
    assert out.shape == ref.shape
    ratio, ok = fwd_ok(out.detach().cpu().numpy(), ref.detach().numpy())
    assert ok, f"forward {ratio}"
    out.backward(go.to(DEV))
    for k in want & set(maps):
        got, exp = leaves[k].grad, leaves_ref[k].grad
        if c["broadcast_rough"] and k == "roughness":
            exp = exp.sum(0, keepdim=True)   # the shared map collects every material's gradient
        assert got is not None, k
        ratio, ok = grad_ok(got.cpu().numpy(), exp.numpy())
        assert ok, f"d_{k} {ratio}"
    for k in set(maps) - want:
        assert leaves[k].grad is None
