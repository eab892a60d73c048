"""The reference-side binding, compiled and called: oracle/_ref/libvimg_ref_binding.so links the reference's
own sources + v-img_b200/adapter/vimg_adapter.cpp (WITH scene_integrator_cuda, i.e. without
VIMG_ADAPTER_FLATTEN_ONLY) + libvimg_cuda.so -- the build a v-img maintainer gets by following INTEGRATION.md.
One process calls the reference's scene_integrator<F>() and the adapter's scene_integrator_cuda() on the SAME
reference scene objects (the seam of src/main.cpp:219-248) and compares the two std::vector<glm::vec3>.

(The scene objects are rebuilt from a golden .vscene by the oracle because /root/reference's scene files do not
exist on the GPU box; the adapter then flattens those reference objects again, which is the path under test.)
"""
import numpy as np
import pytest

from helpers import golden_scene_path

pytestmark = pytest.mark.gpu


def _binding(refcpu):
    if not refcpu.available("binding"):
        pytest.skip("oracle/_ref/libvimg_ref_binding.so not built (make -C oracle binding)")
    return refcpu


@pytest.mark.parametrize("name,integ,spp", [("cornell_small", "mis", 16), ("cornell_small", "material", 8),
                                            ("disney_small", "mis", 16), ("glass_small", "mis", 16),
                                            ("cornell_small", "s_normal", 4), ("synthetic_small", "mis", 8),
                                            ("cornell_box_spheres", "mis", 4)])
def test_scene_integrator_cuda_equals_scene_integrator(vimg, man, refcpu, name, integ, spp):
    r = _binding(refcpu)
    d = vimg.LoadedDesc(golden_scene_path(name))
    rs = r.RefScene.from_desc(d.ptr, spp, man[name]["depth"], "mis", flavour="binding")
    cpu, _ = rs.render(integ, spp=spp)
    gpu = rs.render_cuda(integ, spp=spp)
    assert np.array_equal(cpu, gpu), f"{int((cpu != gpu).any(-1).sum())} pixels differ"
    rs.close()
    d.close()


def test_scene_integrator_cuda_multi(vimg, man, refcpu):
    r = _binding(refcpu)
    if vimg.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    name = "disney_small"
    d = vimg.LoadedDesc(golden_scene_path(name))
    rs = r.RefScene.from_desc(d.ptr, 16, man[name]["depth"], "mis", flavour="binding")
    cpu, _ = rs.render("mis", spp=16)
    gpu = rs.render_cuda("mis", spp=16, devices=list(range(min(vimg.device_count(), 4))))
    assert np.array_equal(cpu, gpu)
    rs.close()
    d.close()


def test_binding_reports_errors(vimg, man, refcpu):
    r = _binding(refcpu)
    d = vimg.LoadedDesc(golden_scene_path("cornell_small"))
    rs = r.RefScene.from_desc(d.ptr, 4, 64, "mis", flavour="binding")
    with pytest.raises(RuntimeError) as e:
        rs.render_cuda("mis", spp=4, devices=[99])
    assert "device" in str(e.value)
    rs.close()
    d.close()
