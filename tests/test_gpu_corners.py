"""Corners of the bit-exact path that no golden scene reaches (VERDICT r1 "What's weak" 2, 3), each against the
compiled reference rebuilt from the SAME descriptor on this host (oracle/ref_driver.cpp: vref_scene_from_desc
uses the descriptor's BVH as it is, so a hand-made BVH is traversed by both sides):

  * the double-precision fallback of the watertight triangle test when an edge function is exactly 0
    (include/geometry/triangle.h:127-134): a vertex moved onto the pinhole camera's origin makes two edge
    functions 0 for EVERY camera ray;
  * a BVH deeper than 30 levels: one traversal needs max_depth + 2 stack entries, so the stale entries an
    occluded shadow ray leaves (include/bvh.h:143,165-168) no longer fit below the next traversal in the
    device's 64-entry stack and are dropped -- what that costs is measured here;
  * a non-square TextureRG, where the reference's row-stride slip (texture_RG.h:47,52) reads other texels.
"""
import ctypes as C

import numpy as np
import pytest

from helpers import golden_scene_path, rel_mse

pytestmark = pytest.mark.gpu


class _Mesh(C.Structure):  # vimg_mesh (include/vimg_cuda.h)
    _fields_ = [("indices", C.POINTER(C.c_uint32)), ("n_tris", C.c_uint64), ("vertices", C.POINTER(C.c_float)),
                ("n_verts", C.c_uint64), ("normals", C.POINTER(C.c_float)), ("texcoords", C.c_void_p),
                ("n_uv_sets", C.c_uint32), ("color_tex_uv", C.c_uint8), ("normal_tex_uv", C.c_uint8),
                ("mr_tex_uv", C.c_uint8), ("pad_", C.c_uint8), ("material", C.c_uint32)]


class _Prim(C.Structure):  # vimg_prim
    _fields_ = [("type", C.c_uint32), ("a", C.c_uint32), ("b", C.c_uint32)]


class _Sphere(C.Structure):  # vimg_sphere
    _fields_ = [("center", C.c_float * 3), ("radius", C.c_float), ("material", C.c_uint32)]


def _desc(vimg, d):
    return vimg.renderer._Desc.from_address(d.ptr)


def _open_all_boxes(desc):
    """every box of the BVH becomes [-1e30, 1e30]^3: no ray is culled, every primitive is tested in BVH order"""
    bb = np.ctypeslib.as_array(C.cast(desc.bvh.bb_mins_maxes, C.POINTER(C.c_float)), shape=(desc.bvh.n_bb_slots, 3))
    for s in range(desc.bvh.n_bb_slots):
        bb[s, :] = -1e30 if s % 4 < 2 else 1e30  # slots 4k, 4k+1: mins; 4k+2, 4k+3: maxes (include/vimg_cuda.h)


def test_edge_function_exactly_zero_takes_the_double_fallback(vimg, refcpu):
    d = vimg.LoadedDesc(golden_scene_path("cornell_small"))
    desc = _desc(vimg, d)
    assert desc.camera.aperture_radius == 0  # pinhole: every camera ray starts at cam_to_world's translation
    origin = [desc.camera.cam_to_world[12 + k] for k in range(3)]
    meshes = C.cast(desc.meshes, C.POINTER(_Mesh))
    moved = 0
    for m in range(desc.n_meshes):
        if meshes[m].n_tris == 0:
            continue
        v0 = meshes[m].indices[0]  # first vertex of the mesh's first triangle
        for k in range(3):
            meshes[m].vertices[3 * v0 + k] = origin[k]
        moved += 1
        if moved == 2:
            break
    assert moved >= 1
    _open_all_boxes(desc)
    rs = refcpu.RefScene.from_desc(d.ptr, 4, 64, "mis")
    sc = vimg.Scene(d.ptr)
    ids, t = sc.trace_ids(4)
    rids, rt = rs.trace_ids(4)
    assert np.array_equal(ids, rids) and np.array_equal(t, rt)
    for integ in ("s_normal", "g_normal"):
        g, _ = sc.scene_integrator(integ, 4, 64)
        r, _ = rs.render(integ, spp=4)
        assert np.array_equal(g, r), integ
    g, st = sc.scene_integrator("mis", 16, 64)
    r, _ = rs.render("mis", spp=16)
    same = float((g == r).all(-1).mean())
    print(f"vertex on the camera origin, open boxes: mis 16 spp bit-identical pixels {same:.4f}, nan {st['nan_samples']}")
    assert same == 1.0
    sc.close()
    rs.close()
    d.close()


def _chain_bvh(vimg, d, n_leaves):
    """Replace the descriptor's BVH by a chain of n_leaves one-primitive leaves (leaf k holds primitive
    k % n_prims: primitives may repeat, both sides just test them again): node 2k is inner with children
    (2k+1 = leaf k, 2k+2 = the rest), so the tree is n_leaves levels deep.
    Returns the arrays that back the new pointers (keep them alive)."""
    desc = _desc(vimg, d)
    prims = C.cast(desc.prims, C.POINTER(_Prim))
    meshes = C.cast(desc.meshes, C.POINTER(_Mesh))
    spheres = C.cast(desc.spheres, C.POINTER(_Sphere))
    lo = np.empty((n_leaves, 3), np.float32)
    hi = np.empty((n_leaves, 3), np.float32)
    for k in range(n_leaves):
        p = prims[k % desc.n_prims]
        if p.type == 1:
            sp = spheres[p.a]
            c = np.array(list(sp.center), np.float32)
            lo[k], hi[k] = c - np.float32(sp.radius), c + np.float32(sp.radius)
            continue
        m = meshes[p.a]
        idx = [m.indices[3 * p.b + j] for j in range(3)]
        v = np.array([[m.vertices[3 * i + c] for c in range(3)] for i in idx], np.float32)
        lo[k], hi[k] = v.min(0), v.max(0)
    n_nodes = 2 * n_leaves - 1
    nodes = np.zeros((n_nodes, 2), np.uint32)
    bb = np.zeros((2 * n_nodes + 2, 3), np.float32)
    # suffix boxes: rest[k] = union of leaves k..n-1
    rest_lo, rest_hi = lo.copy(), hi.copy()
    for k in range(n_leaves - 2, -1, -1):
        rest_lo[k] = np.minimum(lo[k], rest_lo[k + 1])
        rest_hi[k] = np.maximum(hi[k], rest_hi[k + 1])
    bb[0], bb[2] = rest_lo[0], rest_hi[0]
    for k in range(n_leaves - 1):
        inner, left, right = 2 * k, 2 * k + 1, 2 * k + 2
        nodes[inner] = (left, 0)
        nodes[left] = (k, 1)
        last = k == n_leaves - 2
        if last:
            nodes[right] = (k + 1, 1)
        s = 2 * left + 2  # Lmin, Rmin, Lmax, Rmax
        bb[s], bb[s + 2] = lo[k], hi[k]
        bb[s + 1], bb[s + 3] = (lo[k + 1], hi[k + 1]) if last else (rest_lo[k + 1], rest_hi[k + 1])
    obj = (np.arange(n_leaves) % desc.n_prims).astype(np.uint64)
    desc.bvh.nodes = nodes.ctypes.data
    desc.bvh.n_nodes = n_nodes
    desc.bvh.bb_mins_maxes = bb.ctypes.data
    desc.bvh.n_bb_slots = bb.shape[0]
    desc.bvh.obj_indices = obj.ctypes.data
    desc.bvh.n_obj_indices = n_leaves
    desc.bvh.max_depth = n_leaves
    return nodes, bb, obj


@pytest.mark.parametrize("n_leaves", [24, 40])
def test_bvh_deeper_than_thirty_levels(vimg, refcpu, n_leaves):
    """24 levels: two traversals fit the device's 64-entry stack, stale entries are kept, mis is bit-identical.
    40 levels: they do not fit; ids / distances / normal images (fresh traversals) stay bit-exact, and the mis
    image shows what dropping the stale entries does."""
    d = vimg.LoadedDesc(golden_scene_path("cornell_skylight"))  # half of its shadow rays are occluded: stale entries are common
    keep = _chain_bvh(vimg, d, n_leaves)
    rs = refcpu.RefScene.from_desc(d.ptr, 4, 64, "mis")
    sc = vimg.Scene(d.ptr)
    ids, t = sc.trace_ids(2)
    rids, rt = rs.trace_ids(2)
    assert np.array_equal(ids, rids) and np.array_equal(t, rt)
    assert (ids >= 0).any()
    for integ in ("s_normal", "g_normal"):
        g, _ = sc.scene_integrator(integ, 4, 64)
        r, _ = rs.render(integ, spp=4)
        assert np.array_equal(g, r), integ
    g, st = sc.scene_integrator("mis", 16, 64)
    r, _ = rs.render("mis", spp=16)
    differing = int((g != r).any(-1).sum())
    err = rel_mse(g, r)
    print(f"chain BVH, {n_leaves} levels: mis 16 spp, {differing} of {g.shape[0] * g.shape[1]} pixels differ, relMSE {err:.3e}")
    assert st["nan_samples"] == 0
    if 2 * (n_leaves + 2) <= 64:
        assert differing == 0
    else:
        assert err < 1e-3  # stale entries dropped: a path may miss a self-hit the reference finds
    del keep
    sc.close()
    rs.close()
    d.close()


def test_non_square_texture_rg(vimg, refcpu):
    """TextureRG 64 x 48 (wide): `next_x + y * height` of texture_RG.h:47,52 is not `next_x + y * width`."""
    rs = refcpu.RefScene.synthetic(G=40, width=160, height=90, tex_size=256, env_w=256, env_h=128, spp=32, depth=64,
                                   bvh="binned", with_rg=2, aperture=0.0)
    sc = vimg.Scene(rs.desc_ptr())
    for integ, spp in (("mis", 32), ("material", 16)):
        g, st = sc.scene_integrator(integ, spp, 64)
        r, _ = rs.render(integ, spp=spp)
        same = float((g == r).all(-1).mean())
        print(f"non-square TextureRG, {integ} {spp} spp: bit-identical pixels {same:.4f}")
        assert same == 1.0 and st["nan_samples"] == 0
    # and the slip is really exercised: the square texture of the same scene gives another image
    rs2 = refcpu.RefScene.synthetic(G=40, width=160, height=90, tex_size=256, env_w=256, env_h=128, spp=32, depth=64,
                                    bvh="binned", with_rg=1, aperture=0.0)
    r2, _ = rs2.render("mis", spp=8)
    r1, _ = rs.render("mis", spp=8)
    assert not np.array_equal(r1, r2)
    sc.close()
    rs.close()
    rs2.close()


def test_scene_create_rejects_a_bvh_that_is_not_a_tree_of_the_stated_depth(vimg):
    """The device's traversal stacks hold max_depth + 2 entries and are not bounds checked: vimg_scene_create walks
    the tree once and refuses an understated depth, a shared child and a cycle."""
    d = vimg.LoadedDesc(golden_scene_path("cornell_skylight"))
    keep = _chain_bvh(vimg, d, 12)
    desc = _desc(vimg, d)
    vimg.Scene(d.ptr).close()  # the chain itself is fine
    desc.bvh.max_depth = 5     # 12 levels deep
    with pytest.raises(vimg.VimgError) as e:
        vimg.Scene(d.ptr)
    assert "deeper" in str(e.value)
    desc.bvh.max_depth = 12
    nodes = keep[0]
    saved = nodes[4].copy()
    nodes[4] = (1, 0)          # inner node 4 now points back at children (1, 2): node 2 is its own ancestor
    with pytest.raises(vimg.VimgError) as e:
        vimg.Scene(d.ptr)
    assert "tree" in str(e.value)
    nodes[4] = saved
    vimg.Scene(d.ptr).close()
    d.close()


def test_scene_create_rejects_an_empty_rg_texture(vimg, refcpu):
    rs = refcpu.RefScene.synthetic(G=8, width=32, height=32, tex_size=16, env_w=16, env_h=8, spp=1, depth=4,
                                   bvh="binned", with_rg=1, aperture=0.0)
    desc = vimg.renderer._Desc.from_address(rs.desc_ptr())

    class _TexRG(C.Structure):  # vimg_texture_rg
        _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("wrap_u", C.c_uint32), ("wrap_v", C.c_uint32),
                    ("data", C.c_void_p)]
    rg = C.cast(desc.textures_rg, C.POINTER(_TexRG))
    assert desc.n_textures_rg == 1
    vimg.Scene(rs.desc_ptr()).close()
    w = rg[0].width
    rg[0].width = 0
    with pytest.raises(vimg.VimgError):
        vimg.Scene(rs.desc_ptr())
    rg[0].width = w
    rs.close()
