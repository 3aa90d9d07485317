// render.cu -- the normal "2.5D sketch" of a CAD mesh, rasterised on the device (second half).
//
// Reference (file:line in the reference repo): warp_learn/render_open3d.py:29-50 `get_rendered`, called from
// warp_learn/vehicle_utils.py:18-19 once per vehicle and pose: Open3D's OpenGL visualiser draws the mesh with
// vertex colours (vertex_normal + 1) / 2, lighting off, black background, and the 8-bit frame is read back;
// object_mask = all channels zero.  Inside the trajectory loop the mesh is first moved,
// `orig_vertices @ z_rot(theta) + tr` (trajectory_inference.py:363), so the normals are recomputed per item.
//
// Open3D / OpenGL are not available to this repository, so the pipeline is restated (oracle/render_oracle.py spells out
// every rule: Open3D's area-weighted vertex normals, the pinhole camera align_view installs, sample-in-triangle with the
// top-left rule, z-buffer on the perspective-correct depth, perspective-correct colour interpolation, round(c * 255)).
// This file follows the oracle operation for operation in fp64 (compiled with -fmad=false) and is bit-identical to it.
//
// One call renders B items of one mesh.  A vehicle covers a small part of a 1080p frame, so only the screen bounding box of
// each item's projected vertices is z-buffered and shaded; the rest of the outputs is a plain background fill:
//   k_render_background frame-sized outputs <- background (sketch 0, mask 1), per-item bounding boxes <- empty
//   k_render_vertices   (vertex, item): world position (optional rigid move), camera space, image-plane position; bounding box
//   k_render_colours    (vertex, item): vertex normal from the incident triangles in ascending order (CSR), colour
//   k_render_clear      z-buffer keys of the bounding box <- +inf
//   k_render_raster     warp per (triangle, item): edge-function coverage over the triangle's bounding box, 64-bit atomicMin of
//                       (float32 depth bits << 32 | triangle index): nearest wins, ties go to the lower index
//   k_render_resolve    (bounding-box pixel, item): winner triangle -> interpolated colour -> uint8 sketch + mask
#include <cuda_runtime.h>
#include <climits>
#include <cstdint>
#include "../../include/fusg.h"
#include "fusg_common.h"

namespace fusg {

__global__ void __launch_bounds__(256) k_render_vertices(const double *__restrict__ verts, const double *__restrict__ rot, const double *__restrict__ tr,
                                                         const double *__restrict__ E, const double *__restrict__ K, double *__restrict__ Vw,
                                                         double *__restrict__ uvz, int *__restrict__ bbox, int Nv, int H, int W) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (v >= Nv) return;
    double x = verts[3 * v], y = verts[3 * v + 1], z = verts[3 * v + 2];
    if (rot) {
        const double *M = rot + 9 * b;                   // row vector times matrix: out_c = (v0*M0c + v1*M1c) + v2*M2c
        const double ox = (x * M[0] + y * M[3]) + z * M[6];
        const double oy = (x * M[1] + y * M[4]) + z * M[7];
        const double oz = (x * M[2] + y * M[5]) + z * M[8];
        x = ox; y = oy; z = oz;
        if (tr) { x = x + tr[3 * b]; y = y + tr[3 * b + 1]; z = z + tr[3 * b + 2]; }
    }
    double *w = Vw + ((size_t)b * Nv + v) * 3;
    w[0] = x; w[1] = y; w[2] = z;
    const double *e = E + 12 * b, *k = K + 9 * b;
    const double cx = ((e[0] * x + e[1] * y) + e[2] * z) + e[3];
    const double cy = ((e[4] * x + e[5] * y) + e[6] * z) + e[7];
    const double cz = ((e[8] * x + e[9] * y) + e[10] * z) + e[11];
    const double pcx = W / 2.0 - 0.5, pcy = H / 2.0 - 0.5;
    const double zs = cz > 1e-6 ? cz : 1.0;
    double *o = uvz + ((size_t)b * Nv + v) * 3;
    const double pu = k[0] * cx / zs + pcx, pv = k[4] * cy / zs + pcy;
    o[0] = pu;
    o[1] = pv;
    o[2] = cz;
    if (cz > 1e-6) {                                    // triangles with a vertex behind the camera are not drawn
        int *bb = bbox + 4 * b;
        const double lim = 1e9;
        atomicMin(bb + 0, (int)floor(fmax(fmin(pu, lim), -lim)));
        atomicMin(bb + 1, (int)floor(fmax(fmin(pv, lim), -lim)));
        atomicMax(bb + 2, (int)ceil(fmax(fmin(pu, lim), -lim)));
        atomicMax(bb + 3, (int)ceil(fmax(fmin(pv, lim), -lim)));
    }
}

__global__ void __launch_bounds__(256) k_render_background(uint8_t *__restrict__ img, uint8_t *__restrict__ mask, int *__restrict__ bbox, size_t npix, int B) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    // 16-byte stores where the buffers allow it (torch allocations do), bytes for the rest
    const bool vec = ((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(mask)) & 15) == 0;
    const size_t n_img = vec ? (npix * 3) / 16 : 0, n_mask = vec ? npix / 16 : 0;
    const int4 z4 = make_int4(0, 0, 0, 0), o4 = make_int4(0x01010101, 0x01010101, 0x01010101, 0x01010101);
    for (size_t i = i0; i < n_img; i += stride) reinterpret_cast<int4 *>(img)[i] = z4;
    for (size_t i = i0; i < n_mask; i += stride) reinterpret_cast<int4 *>(mask)[i] = o4;
    for (size_t i = n_img * 16 + i0; i < npix * 3; i += stride) img[i] = 0;
    for (size_t i = n_mask * 16 + i0; i < npix; i += stride) mask[i] = 1;
    if (i0 < (size_t)B) { bbox[4 * i0] = INT_MAX; bbox[4 * i0 + 1] = INT_MAX; bbox[4 * i0 + 2] = INT_MIN; bbox[4 * i0 + 3] = INT_MIN; }
}

// the part of item b's bounding box inside the frame: x0..x1, y0..y1 (empty: x1 < x0)
__device__ __forceinline__ void item_box(const int *__restrict__ bbox, int b, int H, int W, int &x0, int &y0, int &x1, int &y1) {
    x0 = max(bbox[4 * b], 0); y0 = max(bbox[4 * b + 1], 0);
    x1 = min(bbox[4 * b + 2], W - 1); y1 = min(bbox[4 * b + 3], H - 1);
}

__global__ void __launch_bounds__(256) k_render_colours(const double *__restrict__ Vw, const int32_t *__restrict__ tris, const int32_t *__restrict__ adj_off,
                                                        const int32_t *__restrict__ adj_tri, double *__restrict__ col, int Nv) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (v >= Nv) return;
    const double *P = Vw + (size_t)b * Nv * 3;
    double ax = 0, ay = 0, az = 0;
    for (int k = adj_off[v]; k < adj_off[v + 1]; ++k) {
        const int t = adj_tri[k];
        const double *p0 = P + 3 * tris[3 * t], *p1 = P + 3 * tris[3 * t + 1], *p2 = P + 3 * tris[3 * t + 2];
        const double e1x = p1[0] - p0[0], e1y = p1[1] - p0[1], e1z = p1[2] - p0[2];
        const double e2x = p2[0] - p0[0], e2y = p2[1] - p0[1], e2z = p2[2] - p0[2];
        ax = ax + (e1y * e2z - e1z * e2y);
        ay = ay + (e1z * e2x - e1x * e2z);
        az = az + (e1x * e2y - e1y * e2x);
    }
    double norm = sqrt((ax * ax + ay * ay) + az * az);
    if (norm == 0.0) norm = 1.0;
    double *c = col + ((size_t)b * Nv + v) * 3;
    c[0] = (ax / norm + 1.0) / 2.0;
    c[1] = (ay / norm + 1.0) / 2.0;
    c[2] = (az / norm + 1.0) / 2.0;
}

// grid (blocks, B): grid-stride over the item's bounding box
__global__ void __launch_bounds__(256) k_render_clear(unsigned long long *__restrict__ zbuf, const int *__restrict__ bbox, int H, int W) {
    const int b = blockIdx.y;
    int x0, y0, x1, y1;
    item_box(bbox, b, H, W, x0, y0, x1, y1);
    if (x1 < x0 || y1 < y0) return;
    const int bw = x1 - x0 + 1, n = bw * (y1 - y0 + 1);
    unsigned long long *zb = zbuf + (size_t)b * H * W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) zb[(size_t)(y0 + i / bw) * W + x0 + i % bw] = ~0ull;
}

struct TriSetup {
    double x0, y0, x1, y1, x2, y2, z0, z1, z2, area;
    double q0, q1, q2;                                 // 1 / (area * z_k): barycentric / depth weights are e_k * q_k
    int i0, i1, i2;
    bool ok;
};

__device__ __forceinline__ TriSetup tri_setup(const double *__restrict__ uvz, const int32_t *__restrict__ tris, int t) {
    TriSetup s;
    s.i0 = tris[3 * t]; s.i1 = tris[3 * t + 1]; s.i2 = tris[3 * t + 2];
    s.x0 = uvz[3 * s.i0]; s.y0 = uvz[3 * s.i0 + 1]; s.z0 = uvz[3 * s.i0 + 2];
    s.x1 = uvz[3 * s.i1]; s.y1 = uvz[3 * s.i1 + 1]; s.z1 = uvz[3 * s.i1 + 2];
    s.x2 = uvz[3 * s.i2]; s.y2 = uvz[3 * s.i2 + 1]; s.z2 = uvz[3 * s.i2 + 2];
    s.ok = s.z0 > 1e-6 && s.z1 > 1e-6 && s.z2 > 1e-6;
    s.area = (s.x1 - s.x0) * (s.y2 - s.y0) - (s.x2 - s.x0) * (s.y1 - s.y0);
    if (s.area == 0.0 || !(fabs(s.area) < 1e300)) s.ok = false;
    if (s.area < 0) {                                  // normalise the orientation: swap vertices 1 and 2
        const int ti = s.i1; s.i1 = s.i2; s.i2 = ti;
        double td = s.x1; s.x1 = s.x2; s.x2 = td;
        td = s.y1; s.y1 = s.y2; s.y2 = td;
        td = s.z1; s.z1 = s.z2; s.z2 = td;
        s.area = -s.area;
    }
    s.q0 = 1.0 / (s.area * s.z0); s.q1 = 1.0 / (s.area * s.z1); s.q2 = 1.0 / (s.area * s.z2);
    return s;
}

__device__ __forceinline__ bool top_left(double A, double B) { return A > 0 || (A == 0 && B < 0); }

// One WARP per (triangle, item): the lanes tile the triangle's bounding box 8 x 4 pixels at a time.  (A thread per triangle
// left the kernel waiting for the few threads that own large triangles: 79 % of the renderer's time.)  atomicMin is
// commutative, so the result does not depend on how the pixels are dealt out.
__global__ void __launch_bounds__(128) k_render_raster(const double *__restrict__ uvz_all, const int32_t *__restrict__ tris, unsigned long long *__restrict__ zbuf,
                                                       int Nv, int Nt, int H, int W) {
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), b = blockIdx.y, lane = threadIdx.x & 31;
    if (t >= Nt) return;
    const TriSetup s = tri_setup(uvz_all + (size_t)b * Nv * 3, tris, t);
    if (!s.ok) return;
    const double fx0 = fmin(s.x0, fmin(s.x1, s.x2)), fx1 = fmax(s.x0, fmax(s.x1, s.x2));
    const double fy0 = fmin(s.y0, fmin(s.y1, s.y2)), fy1 = fmax(s.y0, fmax(s.y1, s.y2));
    if (!(fx1 >= 0.0 && fy1 >= 0.0 && fx0 <= (double)(W - 1) && fy0 <= (double)(H - 1))) return;
    const int xmin = max((int)ceil(fmax(fx0, 0.0)), 0), xmax = min((int)floor(fmin(fx1, (double)(W - 1))), W - 1);
    const int ymin = max((int)ceil(fmax(fy0, 0.0)), 0), ymax = min((int)floor(fmin(fy1, (double)(H - 1))), H - 1);
    const bool tl0 = top_left(s.y1 - s.y2, s.x2 - s.x1), tl1 = top_left(s.y2 - s.y0, s.x0 - s.x2), tl2 = top_left(s.y0 - s.y1, s.x1 - s.x0);
    unsigned long long *zb = zbuf + (size_t)b * H * W;
    const int lx = lane & 7, ly = lane >> 3;
    for (int y = ymin + ly; y <= ymax; y += 4) {
        const double py = (double)y;
        for (int x = xmin + lx; x <= xmax; x += 8) {
            const double px = (double)x;
            const double e0 = (s.x2 - s.x1) * (py - s.y1) - (s.y2 - s.y1) * (px - s.x1);
            const double e1 = (s.x0 - s.x2) * (py - s.y2) - (s.y0 - s.y2) * (px - s.x2);
            const double e2 = (s.x1 - s.x0) * (py - s.y0) - (s.y1 - s.y0) * (px - s.x0);
            const bool in = (e0 > 0 || (e0 == 0 && tl0)) && (e1 > 0 || (e1 == 0 && tl1)) && (e2 > 0 || (e2 == 0 && tl2));
            if (!in) continue;
            const double iz = (e0 * s.q0 + e1 * s.q1) + e2 * s.q2;
            const float depth = (float)(1.0 / iz);
            const unsigned long long key = ((unsigned long long)__float_as_uint(depth) << 32) | (unsigned)t;
            atomicMin(zb + (size_t)y * W + x, key);
        }
    }
}

__global__ void __launch_bounds__(256) k_render_resolve(const double *__restrict__ uvz_all, const double *__restrict__ col_all, const int32_t *__restrict__ tris,
                                                        const unsigned long long *__restrict__ zbuf, const int *__restrict__ bbox,
                                                        uint8_t *__restrict__ img, uint8_t *__restrict__ mask, int Nv, int H, int W) {
    const int b = blockIdx.y;
    int bx0, by0, bx1, by1;
    item_box(bbox, b, H, W, bx0, by0, bx1, by1);
    if (bx1 < bx0 || by1 < by0) return;
    const int bw = bx1 - bx0 + 1, n = bw * (by1 - by0 + 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int p = (by0 + i / bw) * W + bx0 + i % bw;
    const unsigned long long key = zbuf[(size_t)b * H * W + p];
    if (key == ~0ull) continue;                        // background: already filled
    uint8_t r = 0, g = 0, bl = 0;
    {
        const int t = (int)(key & 0xffffffffull);
        const TriSetup s = tri_setup(uvz_all + (size_t)b * Nv * 3, tris, t);
        const double *col = col_all + (size_t)b * Nv * 3;
        const double px = (double)(p % W), py = (double)(p / W);
        const double e0 = (s.x2 - s.x1) * (py - s.y1) - (s.y2 - s.y1) * (px - s.x1);
        const double e1 = (s.x0 - s.x2) * (py - s.y2) - (s.y0 - s.y2) * (px - s.x2);
        const double e2 = (s.x1 - s.x0) * (py - s.y0) - (s.y1 - s.y0) * (px - s.x0);
        const double w0 = e0 * s.q0, w1 = e1 * s.q1, w2 = e2 * s.q2;
        const double iz = (w0 + w1) + w2;
        uint8_t out[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double val = ((w0 * col[3 * s.i0 + c] + w1 * col[3 * s.i1 + c]) + w2 * col[3 * s.i2 + c]) / iz;
            const double k = floor(val * 255.0 + 0.5);
            out[c] = (uint8_t)fmin(fmax(k, 0.0), 255.0);
        }
        r = out[0]; g = out[1]; bl = out[2];
    }
    uint8_t *o = img + ((size_t)b * H * W + p) * 3;
    o[0] = r; o[1] = g; o[2] = bl;
    mask[(size_t)b * H * W + p] = (r == 0 && g == 0 && bl == 0) ? 1 : 0;
    }
}

}  // namespace fusg

using namespace fusg;

extern "C" size_t fusg_render_workspace_bytes(int B, int Nv, int H, int W) {
    if (B <= 0 || Nv <= 0 || H <= 0 || W <= 0) return 0;
    return (size_t)B * ((size_t)Nv * 9 * sizeof(double) + (size_t)H * W * sizeof(unsigned long long) + 4 * sizeof(int));
}

extern "C" int fusg_render_normals(const double *verts, const int32_t *tris, const int32_t *adj_off, const int32_t *adj_tri, int Nv, int Nt,
                                   const double *rot, const double *tr, const double *E, const double *K, uint8_t *normals, uint8_t *mask,
                                   void *workspace, size_t workspace_bytes, int B, int H, int W, void *stream) {
    if (!verts || !tris || !adj_off || !adj_tri || !E || !K || !normals || !mask || !workspace) return FUSG_ERR_ARG;
    if (B <= 0 || Nv <= 0 || Nt <= 0 || H <= 0 || W <= 0 || (tr && !rot)) return FUSG_ERR_ARG;
    if (B > 65535 || (long long)H * W > 0x7fffffffLL) return FUSG_ERR_UNSUPPORTED;
    if (workspace_bytes < fusg_render_workspace_bytes(B, Nv, H, W)) return FUSG_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    double *Vw = reinterpret_cast<double *>(workspace);
    double *uvz = Vw + (size_t)B * Nv * 3;
    double *col = uvz + (size_t)B * Nv * 3;
    unsigned long long *zbuf = reinterpret_cast<unsigned long long *>(col + (size_t)B * Nv * 3);
    int *bbox = reinterpret_cast<int *>(zbuf + (size_t)B * H * W);
    const dim3 gv((Nv + 255) / 256, B);
    const size_t nz = (size_t)B * H * W;
    const int bgrid = (int)((nz * 3 / 16 + 255) / 256 < (size_t)fusg_num_sms() * 16 ? (nz * 3 / 16 + 255) / 256 + 1 : (size_t)fusg_num_sms() * 16);
    k_render_background<<<bgrid, 256, 0, st>>>(normals, mask, bbox, nz, B);
    k_render_vertices<<<gv, 256, 0, st>>>(verts, rot, tr, E, K, Vw, uvz, bbox, Nv, H, W);
    k_render_colours<<<gv, 256, 0, st>>>(Vw, tris, adj_off, adj_tri, col, Nv);
    // per item: enough blocks for a bounding box of a quarter of the frame in one sweep, grid-stride beyond
    const int pgrid = (H * W / 4 + 255) / 256 < 64 ? ((H * W / 4 + 255) / 256 > 0 ? (H * W / 4 + 255) / 256 : 1) : 64;
    k_render_clear<<<dim3(pgrid, B), 256, 0, st>>>(zbuf, bbox, H, W);
    k_render_raster<<<dim3((Nt + 3) / 4, B), 128, 0, st>>>(uvz, tris, zbuf, Nv, Nt, H, W);
    k_render_resolve<<<dim3(pgrid, B), 256, 0, st>>>(uvz, col, tris, zbuf, bbox, normals, mask, Nv, H, W);
    fusg_count_launch(6);
    return fusg_check_launch();
}
