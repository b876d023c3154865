"""The oracle (CPU restatement) replayed against the fixtures generated from the real reference."""
import numpy as np
import pytest
import torch

from conftest import case_inputs, golden_ct_cases, load_golden
from oracle import pbr_oracle as O


def _t(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


@pytest.mark.parametrize("name", golden_ct_cases())
def test_oracle_matches_reference_forward_and_grads(name):
    z = load_golden(name)
    maps, view, lights, inten, p, multi, per_light = case_inputs(z)
    leaves = {k: _t(v).requires_grad_(True) for k, v in maps.items()}
    lt = _t(z["lights"])
    it = _t(z["intensity"])
    out = O.render(leaves, _t(view), lt, it, p["light_size"], p["light_type"], p["albedo_is_srgb"],
                   p.get("specular_is_srgb", True), p["return_srgb"], p["accumulate"])
    ref = _t(z["out32"])
    assert out.shape == ref.shape
    # bit-identical where the fixtures were generated; other CPUs may differ in the last ulp of pow
    assert torch.allclose(out, ref, rtol=2e-6, atol=2e-7)
    out.backward(_t(z["grad_out"]))
    for k in maps:
        g = leaves[k].grad
        r = _t(z["g32_" + k])
        assert torch.allclose(g, r, rtol=1e-4, atol=1e-4 * float(r.abs().mean()))


def test_oracle_fp64_arbiter_close_to_fp32():
    z = load_golden("ct_metal_point_37x53")
    maps, view, lights, inten, p, *_ = case_inputs(z)
    out64 = O.render({k: _t(v, torch.float64) for k, v in maps.items()}, _t(view, torch.float64),
                     _t(z["lights"], torch.float64), _t(z["intensity"], torch.float64), p["light_size"], p["light_type"])
    assert np.allclose(out64.numpy(), z["out64"], rtol=1e-12, atol=1e-14)
    assert np.abs(out64.numpy() - z["out32"]).max() < 1e-5


def test_oracle_conversions_and_blend_fixture():
    z = load_golden("convert_31x45")
    for srgb in (True, False):
        d, s = O.metallic_to_specular(_t(z["in_m_albedo"]), _t(z["in_m_metallic"]), srgb)
        assert torch.allclose(d, _t(z[f"m2s_diffuse_srgb{int(srgb)}"]), rtol=2e-6, atol=1e-7)
        assert torch.allclose(s, _t(z[f"m2s_specular_srgb{int(srgb)}"]), rtol=2e-6, atol=1e-7)
        b, m = O.specular_to_metallic(_t(z["in_s_albedo"]), _t(z["in_s_specular"]), srgb)
        assert torch.allclose(b, _t(z[f"s2m_basecolor_srgb{int(srgb)}"]), rtol=2e-6, atol=1e-7)
        assert torch.allclose(m, _t(z[f"s2m_metallic_srgb{int(srgb)}"]), rtol=2e-6, atol=1e-7)
    z = load_golden("blend_29x43")
    m1 = {k[4:]: _t(v) for k, v in z.items() if k.startswith("in1_")}
    m2 = {k[4:]: _t(v) for k, v in z.items() if k.startswith("in2_")}
    out = O.blend_maps(m1, m2, _t(z["mask"]))
    out["normal"] = O.process_normal_map(out["normal"])
    for k, v in out.items():
        assert torch.equal(v, _t(z["mask_" + k])), k
    hm = O.sigmoid_mask(m1["height"], m2["height"], 0.1)
    assert torch.allclose(hm, _t(z["height_mask"]), rtol=1e-6, atol=1e-7)
    assert torch.equal(O.gradient_mask(29, 43, "horizontal"), _t(z["grad_h_mask"]))
    assert torch.equal(O.gradient_mask(29, 43, "vertical"), _t(z["grad_v_mask"]))


def test_oracle_config1_fixture_anchor():
    """BASELINE.json configs[0]: tiles fixture at 256x256; SURVEY.md §8c anchor mean 0.4920087."""
    z = load_golden("config1_tiles_256")
    assert abs(float(z["out32"].mean()) - 0.4920087) < 1e-6
    assert abs(float(z["out32"].min()) - 0.27467) < 1e-4 and abs(float(z["out32"].max()) - 0.67585) < 1e-4
    maps, view, lights, inten, p, *_ = case_inputs(z)
    out = O.render({k: _t(v) for k, v in maps.items()}, _t(view), _t(z["lights"]), _t(z["intensity"]),
                   p["light_size"], p["light_type"], p["albedo_is_srgb"])
    assert torch.allclose(out, _t(z["out32"]), rtol=2e-6, atol=2e-7)
