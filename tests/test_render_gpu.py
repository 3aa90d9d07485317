"""GPU parity of the normal-sketch rasteriser (fusg_render_normals through the C ABI) against its software oracle
(oracle/render_oracle.py restates render_open3d.py:29-50): the uint8 sketch and the background mask are bit-exact."""
from types import SimpleNamespace

import numpy as np
import pytest

from future_urban_scene_generation_b200 import synth

pytestmark = pytest.mark.gpu


def _cams(idx, h, w):
    p = synth.make_pose_pair(idx, h, w)
    return p["E_src"], p["E_dst"], p["K"]


@pytest.mark.parametrize("hw", [(256, 256), (180, 320), (720, 1280)])
def test_render_matches_oracle(cuda, hw):
    from future_urban_scene_generation_b200.warp_learn.render import render_normals_batch, MeshOnDevice
    from oracle import render_oracle as RO
    h, w = hw
    V, T = synth.make_car_mesh(1)
    mesh = MeshOnDevice(V, T)
    n = 4 if h <= 256 else 2
    E, K = [], []
    for i in range(n):
        es, ed, k = _cams(20 + i, h, w)
        E += [es, ed]
        K += [k, k]
    normals, mask = render_normals_batch(mesh, np.stack(E), np.stack(K), h, w)
    normals, mask = normals.cpu().numpy(), mask.cpu().numpy()
    assert normals.shape == (2 * n, h, w, 3) and mask.dtype == np.bool_
    for b in range(2 * n):
        ref_img, ref_mask = RO.render_normals(V, T, E[b], K[b], h, w)
        assert np.array_equal(mask[b], ref_mask), b
        assert np.array_equal(normals[b], ref_img), (b, int((normals[b] != ref_img).sum()))
        assert 0.005 * h * w < (~ref_mask).sum() < 0.6 * h * w          # the vehicle is really in the picture


def test_render_with_rigid_move_matches_oracle(cuda):
    """The trajectory loop moves the mesh before every render (`orig_vertices @ z_rot(theta) + tr`,
    trajectory_inference.py:363): normals follow the moved vertices; one item partly leaves the frame."""
    from future_urban_scene_generation_b200.kinematics import z_rot
    from future_urban_scene_generation_b200.warp_learn.render import render_normals_batch
    from oracle import render_oracle as RO
    h = w = 256
    V, T = synth.make_car_mesh(2)
    es, _, k = _cams(31, h, w)
    thetas = [0.0, 0.2, -0.35, 0.1]
    trs = [[0, 0, 0], [0.3, -1.0, 0], [-0.5, -2.5, 0], [3.5, 1.0, 0]]
    rot = np.stack([z_rot(t) for t in thetas])
    tr = np.asarray(trs, np.float64)
    normals, mask = render_normals_batch((V, T), np.stack([es] * 4), k, h, w, rot=rot, tr=tr)
    normals, mask = normals.cpu().numpy(), mask.cpu().numpy()
    for b in range(4):
        ref_img, ref_mask = RO.render_normals(V, T, es, k, h, w, rot=rot[b], tr=tr[b])
        assert np.array_equal(normals[b], ref_img) and np.array_equal(mask[b], ref_mask), b
    assert not np.array_equal(normals[0], normals[1])


def test_get_rendered_dropin(cuda):
    from future_urban_scene_generation_b200.warp_learn.render import get_rendered
    from oracle import render_oracle as RO
    V, T = synth.make_car_mesh(0)
    es, _, k = _cams(5, 256, 256)
    ply = SimpleNamespace(vertices=V, triangles=T)
    src_normal, object_mask = get_rendered(ply, 256, 256, es, k)
    assert src_normal.dtype == np.uint8 and src_normal.shape == (256, 256, 3)
    assert object_mask.dtype == np.bool_ and object_mask.shape == (256, 256)
    ref_img, ref_mask = RO.render_normals(V, T, es, k, 256, 256)
    assert np.array_equal(src_normal, ref_img) and np.array_equal(object_mask, ref_mask)
    assert np.array_equal(object_mask, np.all(src_normal == 0, axis=-1))           # render_open3d.py:48


def test_render_chunked_equals_single_call(cuda):
    from future_urban_scene_generation_b200.warp_learn.render import render_normals_batch
    V, T = synth.make_car_mesh(3)
    E, K = [], []
    for i in range(5):
        es, ed, k = _cams(40 + i, 128, 128)
        E.append(es)
        K.append(k)
    a, ma = render_normals_batch((V, T), np.stack(E), np.stack(K), 128, 128)
    b, mb = render_normals_batch((V, T), np.stack(E), np.stack(K), 128, 128, max_items_per_call=2)
    assert cuda.equal(a, b) and cuda.equal(ma, mb)
