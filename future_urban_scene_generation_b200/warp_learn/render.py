"""Normal "2.5D sketch" renderer on the GPU: the drop-in for the reference's warp_learn/render_open3d.py (an Open3D /
OpenGL window per call) behind include/fusg.h: fusg_render_normals.

    get_rendered(model_ply, w, h, extrinsic, intrinsic) -> (src_normal (h,w,3) uint8, object_mask (h,w) bool)

keeps the reference's name, argument order and return types (render_open3d.py:29-50); `model_ply` is anything with
`.vertices` and `.triangles` (an open3d TriangleMesh, or a SimpleNamespace of arrays).  `render_normals_batch` renders
many (vehicle, step) items of one mesh in one call, each with its own camera and rigid move
(trajectory_inference.py:363), and leaves the results on the device for `pack_vunet_inputs_batch` / `get_icn_inputs_batch`.
Rasterisation rules: oracle/render_oracle.py (bit-identical; Open3D itself is not available to pin against).
"""
import numpy as np

from .. import _lib


def vertex_adjacency(triangles, n_vertices):
    """CSR vertex -> incident triangles, ascending triangle index per vertex (Open3D's summation order)."""
    tri = np.asarray(triangles, np.int64)
    flat = tri.ravel()
    order = np.argsort(flat, kind="stable")
    counts = np.bincount(flat, minlength=n_vertices)
    off = np.zeros(n_vertices + 1, np.int32)
    off[1:] = np.cumsum(counts)
    return off, (order // 3).astype(np.int32)


class MeshOnDevice:
    """Vertices, triangles and the vertex adjacency of one CAD mesh, uploaded once."""

    def __init__(self, vertices, triangles, device=None):
        torch = _lib.require_cuda()
        dev = torch.device(device if device is not None else "cuda")
        V = np.ascontiguousarray(np.asarray(vertices, np.float64).reshape(-1, 3))
        T = np.ascontiguousarray(np.asarray(triangles, np.int32).reshape(-1, 3))
        if len(T) and (T.min() < 0 or T.max() >= len(V)):
            raise ValueError("triangle index out of range")
        off, ids = vertex_adjacency(T, len(V))
        self.nv, self.nt = len(V), len(T)
        self.verts = torch.from_numpy(V).to(dev)
        self.tris = torch.from_numpy(T).to(dev)
        self.adj_off = torch.from_numpy(off).to(dev)
        self.adj_tri = torch.from_numpy(ids).to(dev)
        self.device = dev


def _f64(torch, a, shape, dev):
    t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(a))
    return t.to(device=dev, dtype=torch.float64).reshape(shape).contiguous()


def render_normals_batch(mesh, extrinsics, intrinsics, h, w, rot=None, tr=None, max_items_per_call=None, out=None):
    """mesh: MeshOnDevice or (vertices, triangles); extrinsics (B,3,4)|(B,4,4), intrinsics (B,3,3)|(3,3); rot (B,3,3) and
    tr (B,3), optional: item b renders `vertices @ rot[b] + tr[b]`.  -> (normals (B,h,w,3) u8, mask (B,h,w) bool) on the
    device; mask is True on the BACKGROUND like the reference's `object_mask`.  Asynchronous on the current stream.
    out = (normals u8 (B,h,w,3), mask u8 or bool (B,h,w)): contiguous device tensors (or slices) to render into."""
    torch = _lib.require_cuda()
    if not isinstance(mesh, MeshOnDevice):
        mesh = MeshOnDevice(*mesh)
    dev = mesh.device
    E = extrinsics if isinstance(extrinsics, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(extrinsics))
    if E.shape[-2:] == (4, 4):
        E = E[..., :3, :]
    B = E.shape[0]
    E = _f64(torch, E, (B, 12), dev)
    Kt = intrinsics if isinstance(intrinsics, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(intrinsics))
    if Kt.dim() == 2:
        Kt = Kt.unsqueeze(0).expand(B, 3, 3)
    Kt = _f64(torch, Kt, (B, 9), dev)
    R = _f64(torch, rot, (B, 9), dev) if rot is not None else None
    T = _f64(torch, tr, (B, 3), dev) if tr is not None else None
    if T is not None and R is None:
        raise ValueError("tr needs rot")
    if out is None:
        normals = torch.empty((B, h, w, 3), dtype=torch.uint8, device=dev)
        mask = torch.empty((B, h, w), dtype=torch.uint8, device=dev)
    else:
        normals, mask = out
        if tuple(normals.shape) != (B, h, w, 3) or tuple(mask.shape) != (B, h, w) or normals.dtype != torch.uint8 or \
                mask.dtype not in (torch.uint8, torch.bool) or not (normals.is_contiguous() and mask.is_contiguous()):
            raise ValueError("out must be contiguous (B,h,w,3) uint8 and (B,h,w) uint8/bool device tensors")
        mask = mask.view(torch.uint8)
    L = _lib.lib()
    # the z-buffer costs 8 bytes per pixel and item: bound the workspace (default 1 GiB) by rendering in chunks
    per_item = L.fusg_render_workspace_bytes(1, mesh.nv, h, w)
    chunk = max_items_per_call or max(1, min(B, (1 << 30) // max(per_item, 1)))
    need = L.fusg_render_workspace_bytes(min(chunk, B), mesh.nv, h, w)
    ws = getattr(mesh, "_ws", None)                        # the workspace is reused across calls on the same mesh (stream ordered)
    if ws is None or ws.numel() < need:
        ws = torch.empty((need,), dtype=torch.uint8, device=dev)
        mesh._ws = ws
    with torch.cuda.device(dev):
        for b0 in range(0, B, chunk):
            n = min(chunk, B - b0)
            rc = L.fusg_render_normals(_lib.ptr(mesh.verts), _lib.ptr(mesh.tris), _lib.ptr(mesh.adj_off), _lib.ptr(mesh.adj_tri), mesh.nv, mesh.nt,
                                       _lib.ptr(R[b0:b0 + n]) if R is not None else None, _lib.ptr(T[b0:b0 + n]) if T is not None else None,
                                       _lib.ptr(E[b0:b0 + n]), _lib.ptr(Kt[b0:b0 + n]), _lib.ptr(normals[b0:b0 + n]), _lib.ptr(mask[b0:b0 + n]),
                                       _lib.ptr(ws), ws.numel(), n, int(h), int(w), _lib.stream_ptr(torch))
            _lib.check(rc, "fusg_render_normals")
    normals._keep = (ws, E, Kt, R, T, mesh)
    return normals, mask.view(torch.bool)                  # the kernel writes 0 / 1


def get_rendered(model_ply, w, h, extrinsic, intrinsic):
    """render_open3d.py:29-50, same signature and return types: (src_normal (h,w,3) uint8 RGB, object_mask (h,w) bool,
    True where nothing was drawn)."""
    extrinsic = np.asarray(extrinsic, np.float64)
    assert extrinsic.shape == (3, 4) or extrinsic.shape == (4, 4)          # align_view, render_open3d.py:8
    intrinsic = np.asarray(intrinsic, np.float64)
    assert intrinsic.shape == (3, 3)
    normals, mask = render_normals_batch((np.asarray(model_ply.vertices), np.asarray(model_ply.triangles)),
                                         extrinsic[None], intrinsic[None], int(h), int(w))
    return normals[0].cpu().numpy(), mask[0].cpu().numpy()
