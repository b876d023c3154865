"""
The device math headers (pbr_math.cuh / pbr_shade.cuh) compiled for the host and replayed against the
golden vectors: proves, without a GPU, that the expressions the kernels execute reproduce the
reference within the stated tolerances (only the MUFU pow / rsqrt / rcp seeds differ on the device).
"""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, s2m_ok, case_inputs, fwd_ok, golden_ct_cases, grad_ok, load_golden

F = ctypes.POINTER(ctypes.c_float)
D = ctypes.POINTER(ctypes.c_double)


def fp(a):
    return None if a is None else a.ctypes.data_as(F)


def _args(z):
    maps, view, lights, inten, p, multi, per_light = case_inputs(z)
    a = maps["albedo"]
    wf = 0 if "metallic" in maps else 1
    ms = maps["metallic"] if wf == 0 else maps["specular"]
    B = a.shape[0] if a.ndim == 4 else 1
    H, W = a.shape[-2:]
    L = lights.shape[0]
    lt = 1 if p["light_type"] == "point" else 0
    ls = p["light_size"] or 0.0
    args = (B, H, W, L, wf, lt, int(p["albedo_is_srgb"]), int(p.get("specular_is_srgb", True)), int(p["return_srgb"]),
            int(per_light), ctypes.c_float(ls), fp(a), fp(maps.get("normal")), fp(maps["roughness"]), fp(ms),
            fp(view), fp(lights), fp(inten))
    return args, maps, ms, wf, L, (a, view, lights, inten)


@pytest.mark.parametrize("generic", [0, 1, 2, 3, 5, 7, 15])  # bit 0: generic light mode, bit 1: two texels per call (f2 lanes), bit 2: geometry cache for L > 1, bit 3: all 8 fields cached
@pytest.mark.parametrize("name", golden_ct_cases())
def test_forward_matches_golden(hostsim, name, generic):
    z = load_golden(name)
    args, maps, ms, wf, L, keep = _args(z)
    out = np.zeros(z["out32"].shape, np.float32)
    assert hostsim.hs_ct_forward(*args, fp(out), generic) == 0
    ratio, ok = fwd_ok(out, z["out32"])
    assert ok, f"max err/tol {ratio}"
    # and never further from the fp64 arbiter than twice the reference itself (+ the tolerance)
    e_ref = np.abs(z["out32"] - z["out64"]).max()
    e_us = np.abs(out - z["out64"]).max()
    assert e_us <= 2 * e_ref + 1e-6


@pytest.mark.parametrize("generic", [0, 1, 2, 3, 5, 7, 15, 17, 23])  # bit 0: generic light mode, bit 1: two texels per call (f2 lanes), bit 2: geometry cache for L > 1, bit 3: all 8 fields cached, bit 4: one-pass accumulate backward from the saved forward output
@pytest.mark.parametrize("name", golden_ct_cases())
def test_backward_matches_golden(hostsim, name, generic):
    z = load_golden(name)
    args, maps, ms, wf, L, keep = _args(z)
    go = np.ascontiguousarray(z["grad_out"])
    da = np.zeros_like(maps["albedo"]); dn = np.zeros_like(maps["albedo"])
    dr = np.zeros_like(maps["roughness"]); dm = np.zeros_like(ms)
    di = np.zeros((L, 3), np.float64)
    loss = ctypes.c_double(0)
    rc = hostsim.hs_ct_backward(*args, fp(go), None, ctypes.c_float(0), ctypes.byref(loss), fp(da), fp(dn), fp(dr),
                                fp(dm), di.ctypes.data_as(D), generic)
    assert rc == 0
    names = [("albedo", da), ("roughness", dr), ("metallic" if wf == 0 else "specular", dm)]
    if "normal" in maps:
        names.append(("normal", dn))
    for k, g in names:
        ratio, ok = grad_ok(g, z["g32_" + k])
        assert ok, f"d_{k}: max err/tol {ratio}"


def test_intensity_gradient_and_loss_against_oracle(hostsim):
    """d/d intensity and the fused-loss path against autograd through the oracle."""
    from oracle import pbr_oracle as O

    z = load_golden("ct_metal_point_B2_L3_per_24x36")
    args, maps, ms, wf, L, keep = _args(z)
    p = z["params"]
    leaves = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in maps.items()}
    inten = torch.from_numpy(keep[3]).clone().requires_grad_(True)
    out = O.render(leaves, torch.from_numpy(keep[1]), torch.from_numpy(keep[2]), inten, p["light_size"], p["light_type"],
                   accumulate=False)
    gen = torch.Generator().manual_seed(5)
    target = torch.rand(out.shape, generator=gen)
    loss = ((out - target) ** 2).mean()
    loss.backward()
    tgt = np.ascontiguousarray(target.numpy())
    da = np.zeros_like(maps["albedo"]); dn = np.zeros_like(maps["albedo"])
    dr = np.zeros_like(maps["roughness"]); dm = np.zeros_like(ms)
    di = np.zeros((L, 3), np.float64)
    lsum = ctypes.c_double(0)
    scale = 1.0 / target.numel()
    rc = hostsim.hs_ct_backward(*args, None, fp(tgt), ctypes.c_float(scale), ctypes.byref(lsum), fp(da), fp(dn), fp(dr),
                                fp(dm), di.ctypes.data_as(D), 0)
    assert rc == 0
    assert abs(lsum.value * scale - float(loss.detach())) <= 1e-5 * float(loss.detach())
    for k, g in (("albedo", da), ("normal", dn), ("roughness", dr), ("metallic", dm)):
        ratio, ok = grad_ok(g, leaves[k].grad.numpy())
        assert ok, f"d_{k}: {ratio}"
    ratio, ok = grad_ok(di.astype(np.float32), inten.grad.numpy())
    assert ok, f"d_intensity: {ratio}"


def test_metallic_three_channels(hostsim):
    """A 3-channel metallic map (what to_basecolor_metallic_material produces) shades per channel."""
    from oracle import pbr_oracle as O

    gen = torch.Generator().manual_seed(11)
    H, W = 12, 20
    maps = {"albedo": torch.rand(3, H, W, generator=gen), "roughness": torch.rand(1, H, W, generator=gen) * 0.8 + 0.2,
            "metallic": torch.rand(3, H, W, generator=gen),
            "normal": torch.nn.functional.normalize(torch.randn(3, H, W, generator=gen) * 0.3 + torch.tensor([0, 0, 1.0]).view(3, 1, 1), dim=0)}
    leaves = {k: v.clone().requires_grad_(True) for k, v in maps.items()}
    view = torch.tensor([0.0, 0.0, 1.0]); light = torch.tensor([0.1, 0.1, 1.0]); inten = torch.tensor([1.0, 1.0, 1.0])
    out = O.render(leaves, view, light, inten, 1.0, "point")
    go = torch.rand(out.shape, generator=gen)
    out.backward(go)
    npm = {k: np.ascontiguousarray(v.numpy()) for k, v in maps.items()}
    res = np.zeros((3, H, W), np.float32)
    v_, l_, i_ = (np.ascontiguousarray(t.numpy()) for t in (view, light, inten))
    args = (1, H, W, 1, 2, 1, 1, 1, 1, 0, ctypes.c_float(1.0), fp(npm["albedo"]), fp(npm["normal"]), fp(npm["roughness"]),
            fp(npm["metallic"]), fp(v_), fp(l_), fp(i_))
    assert hostsim.hs_ct_forward(*args, fp(res), 0) == 0
    assert fwd_ok(res, out.detach().numpy())[1]
    da = np.zeros((3, H, W), np.float32); dn = np.zeros_like(da); dr = np.zeros((1, H, W), np.float32); dm = np.zeros_like(da)
    gon = np.ascontiguousarray(go.numpy())
    lsum = ctypes.c_double(0)
    assert hostsim.hs_ct_backward(*args, fp(gon), None, ctypes.c_float(0), ctypes.byref(lsum), fp(da), fp(dn), fp(dr),
                                  fp(dm), None, 0) == 0
    for k, g in (("albedo", da), ("normal", dn), ("roughness", dr), ("metallic", dm)):
        assert grad_ok(g, leaves[k].grad.numpy())[1], k


def test_conversions_match_golden(hostsim):
    z = load_golden("convert_31x45")
    n = 31 * 45
    for srgb in (1, 0):
        a = np.ascontiguousarray(z["in_m_albedo"]); m = np.ascontiguousarray(z["in_m_metallic"])
        d = np.zeros_like(a); s = np.zeros_like(a)
        hostsim.hs_convert_m2s(ctypes.c_int64(n), srgb, fp(a), fp(m), fp(d), fp(s))
        rd, rs = z[f"m2s_diffuse_srgb{srgb}"], z[f"m2s_specular_srgb{srgb}"]
        if srgb:
            assert fwd_ok(d, rd)[1] and fwd_ok(s, rs)[1]
        else:  # no pow on the path: bit exact
            assert np.array_equal(d, rd) and np.array_equal(s, rs)
        a = np.ascontiguousarray(z["in_s_albedo"]); sp = np.ascontiguousarray(z["in_s_specular"])
        b = np.zeros_like(a); mm = np.zeros_like(a)
        hostsim.hs_convert_s2m(ctypes.c_int64(3 * n), srgb, fp(a), fp(sp), fp(b), fp(mm))
        rb, rm = z[f"s2m_basecolor_srgb{srgb}"], z[f"s2m_metallic_srgb{srgb}"]
        if srgb:
            # every texel (conftest.s2m_ok): rel 1e-5, or within twice the reference's own distance from its fp64 run
            for got, key in ((b, "basecolor"), (mm, "metallic")):
                frac, ok = s2m_ok(got, z[f"s2m_{key}_srgb1"], z[f"s2m_{key}64_srgb1"])
                assert ok and frac > 0.995, (key, frac)
        else:
            assert np.array_equal(b, rb) and np.array_equal(mm, rm)


def O_srgb(a):
    from oracle import pbr_oracle as O

    return O.srgb_to_linear(torch.from_numpy(a)).numpy()


def test_blend_bit_exact(hostsim):
    z = load_golden("blend_29x43")
    n = 29 * 43
    mask = np.ascontiguousarray(z["mask"])
    for name in ("albedo", "roughness", "metallic", "height", "normal"):
        a = np.ascontiguousarray(z["in1_" + name]); b = np.ascontiguousarray(z["in2_" + name])
        out = np.zeros_like(a)
        hostsim.hs_blend(ctypes.c_int64(n), a.shape[0], int(name == "normal"), fp(mask), fp(a), fp(b), fp(out))
        assert np.array_equal(out, z["mask_" + name]), name
    for tag, w in (("height", 0.1), ("height_w03", 0.3)):
        h1 = np.ascontiguousarray(z["in1_height"]); h2 = np.ascontiguousarray(z["in2_height"])
        m = np.zeros_like(h1)
        hostsim.hs_sigmoid_mask(ctypes.c_int64(n), fp(h1), fp(h2), ctypes.c_float(0.0), 1, ctypes.c_float(w), fp(m))
        assert np.allclose(m, z[tag + "_mask"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("n", [1, 2, 3, 7, 16, 37, 64, 255, 1024, 4096])
@pytest.mark.parametrize("s", [1.0, 5.0, 0.7])
def test_linspace_bit_exact(hostsim, n, s):
    out = np.zeros(n, np.float32)
    hostsim.hs_linspace(ctypes.c_float(-s / 2), ctypes.c_float(s / 2), n, fp(out))
    assert np.array_equal(out, torch.linspace(-s / 2, s / 2, n).numpy())
    hostsim.hs_linspace(ctypes.c_float(0.0), ctypes.c_float(1.0), n, fp(out))
    assert np.array_equal(out, torch.linspace(0, 1, n).numpy())


def test_srgb_round_trip_and_reference(hostsim):
    from oracle import pbr_oracle as O

    x = np.linspace(-0.2, 1.2, 20001).astype(np.float32)
    lin = np.zeros_like(x); back = np.zeros_like(x)
    hostsim.hs_srgb(ctypes.c_int64(x.size), 1, fp(x), fp(lin))
    hostsim.hs_srgb(ctypes.c_int64(x.size), 0, fp(lin), fp(back))
    assert np.allclose(lin, O.srgb_to_linear(torch.from_numpy(x)).numpy(), rtol=1e-6, atol=1e-8)
    assert np.allclose(back, np.clip(x, 0, 1), atol=1e-6)


def test_normal_ingest_bit_exact(hostsim):
    from oracle import pbr_oracle as O

    gen = torch.Generator().manual_seed(3)
    for ch in (3, 2):
        x = torch.rand(ch, 17, 23, generator=gen)
        ref = O.process_normal_map(x).numpy()
        xin = np.ascontiguousarray(x.numpy()); out = np.zeros((3, 17, 23), np.float32)
        hostsim.hs_ingest_normal(ctypes.c_int64(17 * 23), ch, fp(xin), fp(out))
        if ch == 3:
            assert np.array_equal(out, ref)
        else:
            # aten's vectorised CPU sqrt (the z reconstruction) is not correctly rounded: 1 ulp on some inputs
            assert np.allclose(out, ref, rtol=2.5e-7, atol=1e-8)


def test_shared_reciprocal_division_is_ieee(hostsim):
    rng = np.random.default_rng(0)
    n = 2_000_000
    a = (rng.standard_normal(n) * np.exp(rng.uniform(-6, 6, n))).astype(np.float32)
    b = np.exp(rng.uniform(-8, 8, n)).astype(np.float32)
    assert hostsim.hs_div_check(ctypes.c_int64(n), fp(a), fp(b)) == 0


@pytest.mark.parametrize("light_type,L,per_light,packed", [("point", 1, False, 0), ("point", 3, True, 2), ("point", 4, False, 2),
                                                           ("directional", 1, False, 2), ("directional", 3, False, 0)])
def test_light_and_view_gradients_against_oracle(hostsim, light_type, L, per_light, packed):
    """d/d(light position | direction) and d/d(view direction), which the reference gets from plain autograd
    (cooktorrance.py:95,125-140), against autograd through the oracle in fp64 (sums over the image: the fp32 oracle's
    own summation noise is larger than the kernel's error)."""
    from oracle import pbr_oracle as O

    gen = torch.Generator().manual_seed(31 + L)
    B, H, W = 2, 21, 30
    maps = {"albedo": torch.rand(B, 3, H, W, generator=gen), "roughness": torch.rand(B, 1, H, W, generator=gen) * 0.8 + 0.2,
            "metallic": torch.rand(B, 1, H, W, generator=gen)}
    n = torch.randn(B, 3, H, W, generator=gen) * torch.tensor([0.3, 0.3, 0.0]).view(1, 3, 1, 1) + torch.tensor([0.0, 0.0, 1.0]).view(1, 3, 1, 1)
    maps["normal"] = torch.nn.functional.normalize(n, dim=1)
    ang = torch.arange(L, dtype=torch.float32) * (6.2831853 / L) + 0.3
    lights = torch.stack([0.4 * torch.cos(ang), 0.4 * torch.sin(ang), torch.ones(L) * 0.8], dim=1)
    inten = torch.rand(L, 3, generator=gen) * (1.0 if per_light else 1.5 / L) + 0.2
    view = torch.tensor([0.15, -0.1, 0.9])
    size = 1.0 if light_type == "point" else None

    v64, l64, i64 = (t.double().clone().requires_grad_(True) for t in (view, lights, inten))
    out = O.render({k: t.double() for k, t in maps.items()}, v64, l64, i64, size, light_type, accumulate=not per_light)
    go = torch.rand(out.shape, generator=gen)
    out.backward(go.double())

    a = {k: np.ascontiguousarray(t.numpy()) for k, t in maps.items()}
    vn, ln, inn, gon = (np.ascontiguousarray(t.numpy()) for t in (view, lights, inten, go))
    args = (B, H, W, L, 0, 1 if light_type == "point" else 0, 1, 1, 1, int(per_light), ctypes.c_float(size or 0.0),
            fp(a["albedo"]), fp(a["normal"]), fp(a["roughness"]), fp(a["metallic"]), fp(vn), fp(ln), fp(inn))
    da = np.zeros_like(a["albedo"]); dn = np.zeros_like(a["albedo"]); dr = np.zeros_like(a["roughness"]); dm = np.zeros_like(a["metallic"])
    di = np.zeros((L, 3), np.float64); dl = np.zeros((L, 3), np.float64); dv = np.zeros(3, np.float64)
    loss = ctypes.c_double(0)
    rc = hostsim.hs_ct_backward_geom(*args, fp(gon), None, ctypes.c_float(0), ctypes.byref(loss), fp(da), fp(dn), fp(dr), fp(dm),
                                     di.ctypes.data_as(D), dl.ctypes.data_as(D), dv.ctypes.data_as(D), packed)
    assert rc == 0
    for name, got, want in (("d_lights", dl, l64.grad.numpy()), ("d_view", dv, v64.grad.numpy()), ("d_intensity", di, i64.grad.numpy())):
        tol = 1e-4 * np.abs(want) + 1e-4 * np.abs(want).mean()
        assert np.all(np.abs(got - want) <= tol), (name, got, want)


@pytest.mark.parametrize("tag", ["point_metal", "dir_spec"])
def test_light_and_view_gradients_match_reference_fixture(hostsim, tag):
    """geomgrad_shared_params.npz holds the REFERENCE's own autograd gradients of view_dir, the light position /
    direction and the intensity (make_golden.py geomgrad)."""
    import json

    z = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "geomgrad_shared_params.npz"))
    p = json.loads(str(z[tag + "_params"]))
    wf = 0 if p["workflow"] == "metallic" else 1
    a = {k: np.ascontiguousarray(z[f"{tag}_in_{k}"]) for k in ("albedo", "normal", "roughness", "metallic" if wf == 0 else "specular")}
    ms = a["metallic" if wf == 0 else "specular"]
    H, W = a["albedo"].shape[-2:]
    vn, ln, inn = (np.asarray(p[k], np.float32) for k in ("view", "light", "intensity"))
    go = np.ascontiguousarray(z[tag + "_grad_out"])
    args = (1, H, W, 1, wf, 1 if p["light_type"] == "point" else 0, 1, 1, 1, 0, ctypes.c_float(p["light_size"] or 0.0),
            fp(a["albedo"]), fp(a["normal"]), fp(a["roughness"]), fp(ms), fp(vn), fp(ln), fp(inn))
    da = np.zeros_like(a["albedo"]); dn = np.zeros_like(a["albedo"]); dr = np.zeros_like(a["roughness"]); dm = np.zeros_like(ms)
    di = np.zeros((1, 3), np.float64); dl = np.zeros((1, 3), np.float64); dv = np.zeros(3, np.float64)
    loss = ctypes.c_double(0)
    for packed in (0, 2):
        dl[:] = 0; dv[:] = 0; di[:] = 0
        assert hostsim.hs_ct_backward_geom(*args, fp(go), None, ctypes.c_float(0), ctypes.byref(loss), fp(da), fp(dn), fp(dr),
                                           fp(dm), di.ctypes.data_as(D), dl.ctypes.data_as(D), dv.ctypes.data_as(D), packed) == 0
        for name, got, key in (("d_light", dl[0], "light"), ("d_view", dv, "view"), ("d_intensity", di[0], "intensity")):
            want = z[f"{tag}_g64_{key}"]
            tol = 1e-4 * np.abs(want) + 1e-4 * np.abs(want).mean()
            assert np.all(np.abs(got - want) <= tol), (name, got, want)
            # the reference's own fp32 run is further from fp64 than that or in the same ballpark
            assert np.all(np.abs(got - want) <= 4 * np.abs(z[f"{tag}_g32_{key}"] - want) + tol)


@pytest.mark.parametrize("flags", [1, 3, 17, 19, 23])
def test_saturated_accumulate_backward_one_and_two_pass(hostsim, flags):
    """Strong lights: clamp(sum over lights) gates.  Two-pass backward (bit 4 clear) and the one-pass flavour that reads the
    gate and the encode slope off the saved forward output (bit 4 set; encode(1) is 0.99999994 in fp32, in the reference too)."""
    from oracle import pbr_oracle as O

    gen = torch.Generator().manual_seed(77)
    B, H, W, L = 2, 18, 26, 4
    maps = {"albedo": torch.rand(B, 3, H, W, generator=gen), "roughness": torch.rand(B, 1, H, W, generator=gen) * 0.8 + 0.2,
            "metallic": torch.rand(B, 1, H, W, generator=gen)}
    n = torch.randn(B, 3, H, W, generator=gen) * torch.tensor([0.3, 0.3, 0.0]).view(1, 3, 1, 1) + torch.tensor([0.0, 0.0, 1.0]).view(1, 3, 1, 1)
    maps["normal"] = torch.nn.functional.normalize(n, dim=1)
    ang = torch.arange(L, dtype=torch.float32) * (6.2831853 / L)
    lights = torch.stack([0.4 * torch.cos(ang), 0.4 * torch.sin(ang), torch.ones(L)], dim=1)
    inten = torch.ones(L, 3) * 8.0
    view = torch.tensor([0.0, 0.0, 1.0])
    leaves = {k: t.clone().requires_grad_(True) for k, t in maps.items()}
    out = O.render(leaves, view, lights, inten, 1.0, "point", accumulate=True)
    assert float((out.detach() >= 0.9999999).float().mean()) > 0.01
    go = torch.rand(out.shape, generator=gen)
    out.backward(go)
    a = {k: np.ascontiguousarray(t.numpy()) for k, t in maps.items()}
    vn, ln, inn, gon = (np.ascontiguousarray(t.numpy()) for t in (view, lights, inten, go))
    args = (B, H, W, L, 0, 1, 1, 1, 1, 0, ctypes.c_float(1.0), fp(a["albedo"]), fp(a["normal"]), fp(a["roughness"]),
            fp(a["metallic"]), fp(vn), fp(ln), fp(inn))
    da = np.zeros_like(a["albedo"]); dn = np.zeros_like(a["albedo"]); dr = np.zeros_like(a["roughness"]); dm = np.zeros_like(a["metallic"])
    di = np.zeros((L, 3), np.float64)
    loss = ctypes.c_double(0)
    assert hostsim.hs_ct_backward(*args, fp(gon), None, ctypes.c_float(0), ctypes.byref(loss), fp(da), fp(dn), fp(dr), fp(dm),
                                  di.ctypes.data_as(D), flags) == 0
    for k, g in (("albedo", da), ("normal", dn), ("roughness", dr), ("metallic", dm)):
        ratio, ok = grad_ok(g, leaves[k].grad.numpy())
        assert ok, f"d_{k}: {ratio}"


def test_adjoints_of_conversions_blends_and_ingestion_match_reference_autograd(hostsim):
    """The per-texel adjoint code of pbr_grad_kernels.cuh (convert / blend / normal-ingest backward), compiled for the host,
    against the REFERENCE's own autograd results (tests/golden/autograd_convert_blend.npz, written by make_golden.py)."""
    z = np.load(os.path.join(GOLDEN, "autograd_convert_blend.npz"))
    c = lambda k: np.ascontiguousarray(z[k])
    H, W = z["conv_g0"].shape[-2:]
    n = H * W
    g0, g1 = c("conv_g0"), c("conv_g1")
    for srgb in (1, 0):
        a, m = c("m2s_in_albedo"), c("m2s_in_metallic")
        da, dm = np.zeros_like(a), np.zeros_like(m)
        hostsim.hs_convert_m2s_bwd(ctypes.c_int64(n), srgb, 1, fp(a), fp(m), fp(g0), fp(g1), fp(da), fp(dm))
        assert grad_ok(da, z[f"m2s_srgb{srgb}_g32_albedo"])[1] and grad_ok(dm, z[f"m2s_srgb{srgb}_g32_metallic"])[1]
        a, sp = c("s2m_in_albedo"), c("s2m_in_specular")
        da, ds = np.zeros_like(a), np.zeros_like(sp)
        hostsim.hs_convert_s2m_bwd(ctypes.c_int64(3 * n), srgb, fp(a), fp(sp), fp(g0), fp(g1), fp(da), fp(ds))
        for got, key in ((da, "albedo"), (ds, "specular")):
            # the same ill-conditioned division as the forward: judged against the fp64 run where fp32 autograd itself is off
            r32, r64 = z[f"s2m_srgb{srgb}_g32_{key}"], z[f"s2m_srgb{srgb}_g64_{key}"]
            tol = 1e-4 * np.abs(r32) + 1e-4 * np.abs(r32).mean()
            ok = (np.abs(got - r32) <= tol) | (np.abs(got - r64) <= 2 * np.abs(r32.astype(np.float64) - r64) + tol)
            assert ok.mean() > 0.999, (key, srgb, ok.mean())
    names = ("albedo", "normal", "roughness", "metallic", "height")
    mask = c("blend_mask")
    dmask = np.zeros_like(mask)
    for k in names:
        a, b, g = c(f"blend_in1_{k}"), c(f"blend_in2_{k}"), c(f"blend_gout_{k}")
        da, db = np.zeros_like(a), np.zeros_like(b)
        hostsim.hs_blend_bwd(ctypes.c_int64(n), a.shape[0], int(k == "normal"), fp(mask), fp(a), fp(b), fp(g), fp(da), fp(db), fp(dmask))
        assert grad_ok(da, z[f"blend_mask_g32_in1_{k}"])[1] and grad_ok(db, z[f"blend_mask_g32_in2_{k}"])[1], k
    assert grad_ok(dmask, z["blend_mask_g32_mask"])[1]
    gn = c("ingest_gout")
    for ch in (3, 2):
        src = c(f"ingest_in{ch}")
        d = np.zeros_like(src)
        hostsim.hs_ingest_normal_bwd(ctypes.c_int64(n), ch, fp(src), fp(gn), fp(d))
        assert grad_ok(d, z[f"ingest{ch}_g32"])[1], ch
