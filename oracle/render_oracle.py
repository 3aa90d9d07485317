"""Software oracle of the normal-sketch renderer -- TEST INFRASTRUCTURE ONLY.

The reference renders the "2.5D normal sketch" of a CAD mesh with Open3D's OpenGL visualiser
(warp_learn/render_open3d.py:29-50, called from warp_learn/vehicle_utils.py:18-19):

    model_ply.compute_vertex_normals()
    model_ply.vertex_colors = (vertex_normals + 1) / 2           # lighting off, black background
    src_normal = (capture_screen_float_buffer() * 255).astype(np.uint8)
    object_mask = np.all(src_normal == 0, axis=-1)                 # True on the BACKGROUND

Open3D is an absent third-party dependency (requirements.txt, "open3d", unpinned) and a windowed GL context cannot
exist here, so the GL pipeline is restated -- PARITY UNPINNED against Open3D itself; what is pinned is that the CUDA
rasteriser equals THIS restatement bit for bit (tests/test_render_gpu.py):

  * vertex normals: Open3D's ComputeVertexNormals -- unnormalised triangle normals (v1-v0) x (v2-v0), summed per vertex
    in ascending triangle order (i.e. area weighted), then normalised; fp64 like Open3D's Eigen::Vector3d;
  * camera: the pinhole model align_view installs -- focal lengths from the intrinsic matrix, principal point at the
    window centre (w/2 - 0.5, h/2 - 0.5) ("this must be left as they are", render_open3d.py:21), extrinsic world->camera,
    camera looking down +z; pixel (i, j) is sampled at image-plane position (u, v) = (i, j);
  * rasterisation: every triangle with all three vertices in front of the camera (z > 1e-6: no near-plane clipping of
    partially visible triangles -- a vehicle is in front of the camera), both orientations (equivalent to Open3D's
    back-face culling on a closed mesh), sample-in-triangle by edge functions with the top-left rule, depth test on
    the perspective-correct camera-space z rounded to float32 (ties: the lower triangle index, GL_LESS in draw order);
  * shading: perspective-correct interpolation of the vertex colours, GL's float -> 8-bit conversion
    round(c * 255); the float readback * 255 -> astype(uint8) of the reference is the identity on those values
    (checked for all 256 levels).

Everything is fp64 with the operation order spelled out in `render_normals`, which the kernel follows
(csrc/render.cu, compiled with -fmad=false).  Only tests/ and __graft_entry__.smoke() import this module.
"""
import numpy as np


def vertex_adjacency(triangles, n_vertices):
    """CSR vertex -> incident triangles (ascending triangle index per vertex): (offsets (Nv+1,), tri_ids (3*Nt,))."""
    tri = np.asarray(triangles, np.int64)
    flat = tri.ravel()
    order = np.argsort(flat, kind="stable")              # stable: ascending (triangle, corner) inside each vertex
    counts = np.bincount(flat, minlength=n_vertices)
    off = np.zeros(n_vertices + 1, np.int32)
    off[1:] = np.cumsum(counts)
    return off, (order // 3).astype(np.int32)


def vertex_normals(vertices, triangles):
    V = np.asarray(vertices, np.float64)
    T = np.asarray(triangles, np.int64)
    e1 = V[T[:, 1]] - V[T[:, 0]]
    e2 = V[T[:, 2]] - V[T[:, 0]]
    tn = np.stack([e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1],
                   e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2],
                   e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]], 1)
    off, ids = vertex_adjacency(T, len(V))
    n = np.zeros_like(V)
    for v in range(len(V)):                               # sequential sums in ascending triangle order
        acc = np.zeros(3)
        for t in ids[off[v]:off[v + 1]]:
            acc = acc + tn[t]
        n[v] = acc
    norm = np.sqrt(n[:, 0] * n[:, 0] + n[:, 1] * n[:, 1] + n[:, 2] * n[:, 2])
    norm = np.where(norm == 0.0, 1.0, norm)
    return n / norm[:, None]


def transform_vertices(vertices, rot=None, tr=None):
    """trajectory_inference.py:363: `orig_vertices @ z_rot(theta) + tr`, every entry as the left-to-right sum
    (v0*M0c + v1*M1c) + v2*M2c in fp64."""
    V = np.asarray(vertices, np.float64)
    if rot is None:
        return V.copy()
    M = np.asarray(rot, np.float64)
    out = np.empty_like(V)
    for c in range(3):
        out[:, c] = (V[:, 0] * M[0, c] + V[:, 1] * M[1, c]) + V[:, 2] * M[2, c]
    if tr is not None:
        out = out + np.asarray(tr, np.float64)[None, :]
    return out


def project_vertices(Vw, extrinsic, intrinsic, h, w):
    E = np.asarray(extrinsic, np.float64)[:3]
    K = np.asarray(intrinsic, np.float64)
    x = ((E[0, 0] * Vw[:, 0] + E[0, 1] * Vw[:, 1]) + E[0, 2] * Vw[:, 2]) + E[0, 3]
    y = ((E[1, 0] * Vw[:, 0] + E[1, 1] * Vw[:, 1]) + E[1, 2] * Vw[:, 2]) + E[1, 3]
    z = ((E[2, 0] * Vw[:, 0] + E[2, 1] * Vw[:, 1]) + E[2, 2] * Vw[:, 2]) + E[2, 3]
    cx, cy = w / 2.0 - 0.5, h / 2.0 - 0.5
    zs = np.where(z > 1e-6, z, 1.0)
    u = K[0, 0] * x / zs + cx
    v = K[1, 1] * y / zs + cy
    return u, v, z


def _top_left(A, B):
    return (A > 0) | ((A == 0) & (B < 0))


def render_normals(vertices, triangles, extrinsic, intrinsic, h, w, rot=None, tr=None, return_ids=False):
    """-> (normal sketch (h,w,3) uint8 RGB, object_mask (h,w) bool: True on the background)."""
    T = np.asarray(triangles, np.int64)
    Vw = transform_vertices(vertices, rot, tr)
    col = (vertex_normals(Vw, T) + 1.0) / 2.0
    u, v, z = project_vertices(Vw, extrinsic, intrinsic, h, w)
    key_depth = np.full((h, w), np.inf, np.float32)
    key_tri = np.full((h, w), -1, np.int64)
    for t in range(len(T)):
        i0, i1, i2 = T[t]
        if not (z[i0] > 1e-6 and z[i1] > 1e-6 and z[i2] > 1e-6):
            continue
        x0, y0, x1, y1, x2, y2 = u[i0], v[i0], u[i1], v[i1], u[i2], v[i2]
        area = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0)
        if area == 0.0 or not np.isfinite(area):
            continue
        if area < 0:                                      # normalise the orientation: swap vertices 1 and 2
            i1, i2 = i2, i1
            x1, y1, x2, y2 = x2, y2, x1, y1
            area = -area
        xmin = max(int(np.ceil(min(x0, x1, x2))), 0)
        xmax = min(int(np.floor(max(x0, x1, x2))), w - 1)
        ymin = max(int(np.ceil(min(y0, y1, y2))), 0)
        ymax = min(int(np.floor(max(y0, y1, y2))), h - 1)
        if xmin > xmax or ymin > ymax:
            continue
        px, py = np.meshgrid(np.arange(xmin, xmax + 1, dtype=np.float64), np.arange(ymin, ymax + 1, dtype=np.float64))
        # edge functions: e_k is the one opposite vertex k
        e0 = (x2 - x1) * (py - y1) - (y2 - y1) * (px - x1)
        e1 = (x0 - x2) * (py - y2) - (y0 - y2) * (px - x2)
        e2 = (x1 - x0) * (py - y0) - (y1 - y0) * (px - x0)
        in0 = (e0 > 0) | ((e0 == 0) & _top_left(y1 - y2, x2 - x1))
        in1 = (e1 > 0) | ((e1 == 0) & _top_left(y2 - y0, x0 - x2))
        in2 = (e2 > 0) | ((e2 == 0) & _top_left(y0 - y1, x1 - x0))
        inside = in0 & in1 & in2
        if not inside.any():
            continue
        q0, q1, q2 = 1.0 / (area * z[i0]), 1.0 / (area * z[i1]), 1.0 / (area * z[i2])
        iz = (e0 * q0 + e1 * q1) + e2 * q2
        depth = (1.0 / iz).astype(np.float32)
        sub_d = key_depth[ymin:ymax + 1, xmin:xmax + 1]
        sub_t = key_tri[ymin:ymax + 1, xmin:xmax + 1]
        win = inside & (depth < sub_d)                   # ties keep the earlier (lower-index) triangle
        sub_d[win] = depth[win]
        sub_t[win] = t
    img = np.zeros((h, w, 3), np.uint8)
    ys, xs = np.nonzero(key_tri >= 0)
    for y, x in zip(ys, xs):
        t = key_tri[y, x]
        i0, i1, i2 = T[t]
        x0, y0, x1, y1, x2, y2 = u[i0], v[i0], u[i1], v[i1], u[i2], v[i2]
        area = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0)
        if area < 0:
            i1, i2 = i2, i1
            x1, y1, x2, y2 = x2, y2, x1, y1
            area = -area
        px, py = float(x), float(y)
        e0 = (x2 - x1) * (py - y1) - (y2 - y1) * (px - x1)
        e1 = (x0 - x2) * (py - y2) - (y0 - y2) * (px - x2)
        e2 = (x1 - x0) * (py - y0) - (y1 - y0) * (px - x0)
        w0, w1, w2 = e0 * (1.0 / (area * z[i0])), e1 * (1.0 / (area * z[i1])), e2 * (1.0 / (area * z[i2]))
        iz = (w0 + w1) + w2
        for c in range(3):
            val = ((w0 * col[i0, c] + w1 * col[i1, c]) + w2 * col[i2, c]) / iz
            k = np.floor(val * 255.0 + 0.5)
            img[y, x, c] = int(min(max(k, 0.0), 255.0))
    mask = np.all(img == 0, axis=-1)
    return (img, mask, key_tri) if return_ids else (img, mask)
