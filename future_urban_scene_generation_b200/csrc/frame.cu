// frame.cu -- paste-back of completed vehicle crops into video frames (include/fusg.h: fusg_resize_u8, fusg_paste_back).
//
// Reference semantics: trajectory_inference.py:236-250 (also :184-198, :393-407, :428-442), per vehicle
//     crop_inv = cv2.resize(net_image, crop_size_orig[::-1]);  crop_inv = crop_inv[pad_before : -pad_after]
//     out_frame = zeros;  out_frame[crop_xy_min ...] = crop_inv;  img_output[dst_sketch_mask] = out_frame[dst_sketch_mask]
// with crop_info from warp_learn/models.py:334-342.  Vehicles are pasted in sequence, so where masks overlap the last
// vehicle wins.  Here the whole batch is two launches: k_paste_owner resolves "last writer" per frame pixel with an
// atomicMax over the item index, k_paste_apply evaluates cv2.resize (INTER_LINEAR, 8-bit) pointwise for exactly the
// pixels that survive -- no intermediate resized crop, no per-vehicle frame-sized temporaries.
//
// cv2.resize is restated from its observable behaviour (OpenCV 4.13; oracle/frame_oracle.py holds the same
// statement in numpy, pinned to cv2 by tests/golden/frame_golden.json): double scale = 1/(dst/src), float coordinate,
// 11-bit coefficients rounded half-to-even, int32 horizontal pass, (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2
// vertically; an exact halving in both axes is a 2x2 box average.  Compiled with -fmad=false: the coordinate
// arithmetic must round like the host code it mirrors.
#include <cstdint>

#include "../../include/fusg.h"
#include "fusg_common.h"

namespace fusg {

struct Tap { int i0, i1, a0, a1; };

// coefficient pair of destination index d on an axis of sn source and dn destination samples
__device__ __forceinline__ Tap resize_tap(int dn, int sn, int d, bool horizontal) {
    const double scale = 1.0 / ((double)dn / (double)sn);
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f = f - (float)s;
    Tap t;
    if (horizontal) {
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= sn - 1) { s = sn - 1; f = 0.f; }
        t.i0 = s;
        t.i1 = min(s + 1, sn - 1);
    } else {
        t.i0 = min(max(s, 0), sn - 1);
        t.i1 = min(max(s + 1, 0), sn - 1);
    }
    t.a0 = __float2int_rn((1.f - f) * 2048.f);
    t.a1 = __float2int_rn(f * 2048.f);
    return t;
}

// pixel (y, x) of cv2.resize(src[sh, sw, 3], (dw, dh)); three channels
__device__ __forceinline__ void resize_pixel(const uint8_t *__restrict__ src, int sh, int sw, int dh, int dw, int y, int x, uint8_t out[3]) {
    if (sw == 2 * dw && sh == 2 * dh) {
        const uint8_t *p = src + ((size_t)(2 * y) * sw + 2 * x) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) out[c] = (uint8_t)((p[c] + p[3 + c] + p[(size_t)sw * 3 + c] + p[(size_t)sw * 3 + 3 + c] + 2) >> 2);
        return;
    }
    const Tap tx = resize_tap(dw, sw, x, true), ty = resize_tap(dh, sh, y, false);
    const uint8_t *r0 = src + (size_t)ty.i0 * sw * 3, *r1 = src + (size_t)ty.i1 * sw * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int s0 = r0[tx.i0 * 3 + c] * tx.a0 + r0[tx.i1 * 3 + c] * tx.a1;
        const int s1 = r1[tx.i0 * 3 + c] * tx.a0 + r1[tx.i1 * 3 + c] * tx.a1;
        const int v = (((ty.a0 * (s0 >> 4)) >> 16) + ((ty.a1 * (s1 >> 4)) >> 16) + 2) >> 2;
        out[c] = (uint8_t)min(max(v, 0), 255);
    }
}

// one thread per destination pixel; blockIdx.y = item
__global__ void __launch_bounds__(256) k_resize_u8(const uint8_t *__restrict__ src, const long long *__restrict__ src_off,
                                                   const int *__restrict__ src_hw, uint8_t *__restrict__ dst,
                                                   const long long *__restrict__ dst_off, const int *__restrict__ dst_hw) {
    const int b = blockIdx.y;
    const int sh = src_hw[2 * b], sw = src_hw[2 * b + 1], dh = dst_hw[2 * b], dw = dst_hw[2 * b + 1];
    const uint8_t *s = src + src_off[b];
    uint8_t *d = dst + dst_off[b];
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < (long long)dh * dw; p += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(p / dw), x = (int)(p - (long long)y * dw);
        uint8_t o[3];
        resize_pixel(s, sh, sw, dh, dw, y, x, o);
        d[p * 3] = o[0]; d[p * 3 + 1] = o[1]; d[p * 3 + 2] = o[2];
    }
}

constexpr int PASTE_INFO = 9;   // frame, h_orig, w_orig, pad_x0, pad_y0, pad_x1, pad_y1, x_min, y_min

// owner[frame][y][x] = highest item index whose mask covers the pixel (-1: nobody)
__global__ void __launch_bounds__(256) k_paste_owner(const uint8_t *__restrict__ masks, const long long *__restrict__ mask_off,
                                                     const int *__restrict__ mask_rect, const int *__restrict__ info,
                                                     int *__restrict__ owner, int Hf, int Wf) {
    const int b = blockIdx.y;
    const int x0 = mask_rect[4 * b], y0 = mask_rect[4 * b + 1], w = mask_rect[4 * b + 2], h = mask_rect[4 * b + 3];
    const int frame = info[PASTE_INFO * b];
    const uint8_t *m = masks + mask_off[b];
    int *own = owner + (size_t)frame * Hf * Wf;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < (long long)w * h; p += (long long)gridDim.x * blockDim.x) {
        if (!m[p]) continue;
        const int yy = (int)(p / w), xx = (int)(p - (long long)yy * w);
        const int fy = y0 + yy, fx = x0 + xx;
        if (fy >= 0 && fy < Hf && fx >= 0 && fx < Wf) atomicMax(&own[(size_t)fy * Wf + fx], b);
    }
}

__global__ void __launch_bounds__(256) k_paste_apply(uint8_t *__restrict__ frames, const uint8_t *__restrict__ crops,
                                                     const int *__restrict__ info, const int *__restrict__ owner, int Hf, int Wf, int S) {
    const int frame = blockIdx.y;
    const size_t fbase = (size_t)frame * Hf * Wf;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < (long long)Hf * Wf; p += (long long)gridDim.x * blockDim.x) {
        const int o = owner[fbase + p];
        if (o < 0) continue;
        const int *in = info + PASTE_INFO * o;
        const int h = in[1], w = in[2], px0 = in[3], py0 = in[4], px1 = in[5], py1 = in[6], xmin = in[7], ymin = in[8];
        const int y = (int)(p / Wf), x = (int)(p - (long long)y * Wf);
        const int ry = y - ymin, rx = x - xmin;
        uint8_t v[3] = {0, 0, 0};                              // out_frame is zero outside the pasted rectangle
        if (ry >= 0 && ry < h - py0 - py1 && rx >= 0 && rx < w - px0 - px1)
            resize_pixel(crops + (size_t)o * S * S * 3, S, S, h, w, ry + py0, rx + px0, v);
        uint8_t *d = frames + (fbase + p) * 3;
        d[0] = v[0]; d[1] = v[1]; d[2] = v[2];
    }
}

}  // namespace fusg

using namespace fusg;

extern "C" int fusg_resize_u8(const uint8_t *src, const long long *src_off, const int32_t *src_hw, uint8_t *dst, const long long *dst_off,
                              const int32_t *dst_hw, int B, int max_dst_pixels, void *stream) {
    if (!src || !src_off || !src_hw || !dst || !dst_off || !dst_hw || B <= 0 || max_dst_pixels <= 0) return FUSG_ERR_ARG;
    if (B > 65535) return FUSG_ERR_UNSUPPORTED;
    int gx = (max_dst_pixels + 255) / 256;
    if (gx > 4096) gx = 4096;
    k_resize_u8<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(src, src_off, src_hw, dst, dst_off, dst_hw);
    fusg_count_launch(1);
    return fusg_check_launch();
}

extern "C" size_t fusg_paste_workspace_bytes(int F, int Hf, int Wf) {
    return F <= 0 || Hf <= 0 || Wf <= 0 ? 0 : (size_t)F * Hf * Wf * sizeof(int);
}

extern "C" int fusg_paste_back(uint8_t *frames, const uint8_t *crops, const uint8_t *masks, const long long *mask_off, const int32_t *mask_rect,
                               const int32_t *info, void *workspace, size_t workspace_bytes, int B, int F, int Hf, int Wf, int S,
                               int max_mask_pixels, void *stream) {
    if (!frames || !crops || !masks || !mask_off || !mask_rect || !info || !workspace) return FUSG_ERR_ARG;
    if (B <= 0 || F <= 0 || Hf <= 0 || Wf <= 0 || S <= 0 || max_mask_pixels <= 0) return FUSG_ERR_ARG;
    if (B > 65535 || F > 65535) return FUSG_ERR_UNSUPPORTED;
    if (workspace_bytes < fusg_paste_workspace_bytes(F, Hf, Wf)) return FUSG_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    int *owner = reinterpret_cast<int *>(workspace);
    if (fusg_record_cuda(cudaMemsetAsync(owner, 0xFF, (size_t)F * Hf * Wf * sizeof(int), st)) != FUSG_OK) return FUSG_ERR_CUDA;
    int gx = (max_mask_pixels + 255) / 256;
    if (gx > 8192) gx = 8192;
    k_paste_owner<<<dim3(gx, B), 256, 0, st>>>(masks, mask_off, mask_rect, info, owner, Hf, Wf);
    int gf = (int)(((long long)Hf * Wf + 255) / 256);
    if (gf > 16384) gf = 16384;
    k_paste_apply<<<dim3(gf, F), 256, 0, st>>>(frames, crops, info, owner, Hf, Wf, S);
    fusg_count_launch(2);
    return fusg_check_launch();
}
