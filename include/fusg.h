/*
 * fusg.h -- C ABI of the B200-native novel-view-completion hot path.
 *
 * The reference (alexj94/future_urban_scene_generation) is pure Python; it has no FFI for this
 * path.  Each entry point below names the reference Python interface it stands behind; the
 * Python host layer (future_urban_scene_generation_b200/{warp_learn,vunet}) binds them with
 * ctypes and re-exposes the reference's own module/function names.  INTEGRATION.md shows the
 * binding a reference maintainer would add.
 *
 * Conventions: plain pointers and sizes only; every pointer is a DEVICE pointer unless its name
 * ends in _host; `stream` is a cudaStream_t passed as void*; calls are asynchronous on that
 * stream, never allocate, never synchronise; the return value is 0 or a negative FUSG_ERR_*.
 */
#ifndef FUSG_H_
#define FUSG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FUSG_OK 0
#define FUSG_ERR_ARG (-1)          /* null pointer / non-positive size                    */
#define FUSG_ERR_UNSUPPORTED (-2)  /* shape outside what the kernels cover                */
#define FUSG_ERR_CUDA (-3)         /* a CUDA runtime call failed (see fusg_last_error)    */
#define FUSG_ERR_WORKSPACE (-4)    /* workspace too small                                 */

/* Version / diagnostics ---------------------------------------------------------------- */
int fusg_version(void);
const char *fusg_last_error(void);     /* text of the last CUDA error seen by this library  */
int fusg_kernel_launches(void);        /* number of kernels this library launched so far     */

/* ====================================================================================== */
/* Warp stage                                                                              */
/* ====================================================================================== */

/* Bytes of device workspace fusg_warp_fused needs for a batch of B crops. */
size_t fusg_warp_workspace_bytes(int B);

/*
 * Fused planar warp for a batch of B vehicle crops.  Stands behind
 *     warp_learn/online_visibility.py:105-150   compute_visibility  (source and destination pose)
 *     warp_learn/planes_utils.py:11-37          get_planes          (polygon masks, never materialised)
 *     warp_learn/planes_utils.py:40-82          warp_unwarp_planes  (first return value only)
 * i.e. the sequence trajectory_inference.py:165-174 runs per vehicle and per future step.
 *
 *   src      [B,H,W,3] u8   source crops (HWC, as cv2 images)
 *   src_kp   [B,12,2]  i32  source-pose plane vertices: the int32-truncated 2D keypoints of
 *                           planes_utils.py:22-27, utils/keypoint_utils.py:9-13 (_KP_NAMES) order
 *   dst_kp   [B,12,2]  i32  destination-pose plane vertices, same convention
 *   K        [B,3,3]   f64  intrinsics
 *   E_src    [B,3,4]   f64  source extrinsic (world->camera), row major
 *   E_dst    [B,3,4]   f64  destination extrinsic
 *   kp3d     [B,12,3]  f64  CAD keypoints, _KP_NAMES order
 *   warped   [B,5,H,W,3] u8 out: planes_warped (left,right,roof,front,back); fully written
 *   vis      [B,2,7]   u8   out: visibility of (left,right,roof,front,back,front_bt,back_bt)
 *                           for the source (index 0) and destination (index 1) pose
 *   plane_j  [B,5]     i8   out: target plane index j of source plane i, or -1 if skipped
 *   H12      [B,5,9]   f64  out (may be NULL): homography of source plane i (zeros if skipped)
 *   workspace, workspace_bytes: device scratch, >= fusg_warp_workspace_bytes(B)
 *
 * Supported: 8 <= H,W <= 256 (the crop is staged in shared memory); every vertex must lie inside
 * the frame, else that crop's outputs are zero and plane_j = -2 (the reference's clipped-polygon
 * regime is not covered).
 */
int fusg_warp_fused(const uint8_t *src, const int32_t *src_kp, const int32_t *dst_kp,
                    const double *K, const double *E_src, const double *E_dst, const double *kp3d,
                    uint8_t *warped, uint8_t *vis, int8_t *plane_j, double *H12,
                    void *workspace, size_t workspace_bytes,
                    int B, int H, int W, void *stream);

/*
 * Visibility only (compute_visibility, online_visibility.py:105-150) for B poses.
 *   K [B,3,3], E [B,3,4], kp3d [B,12,3] f64 -> vis [B,7] u8, pts [B,12,2] i32 (int()-truncated
 *   projections; may be NULL), areas [B,7,2] i32 (absolute, occluded; may be NULL).
 */
int fusg_visibility(const double *K, const double *E, const double *kp3d,
                    uint8_t *vis, int32_t *pts, int32_t *areas, int B, int H, int W, void *stream);

/*
 * get_planes (planes_utils.py:11-37): planes[b,p] = image[b] * fillPoly(mask of plane p).
 *   img [B,H,W,3] u8, kp [B,12,2] i32 -> planes [B,5,H,W,3] u8.   Any H,W >= 1.
 */
int fusg_get_planes(const uint8_t *img, const int32_t *kp, uint8_t *planes, int B, int H, int W, void *stream);

/*
 * Homographies of cv2.findHomography(src, dst) (method 0) for N point sets of n = 4 or 6 points.
 *   src,dst [N,n,2] i32 -> H [N,9] f64, ok [N] u8 (0 where OpenCV returns None).
 */
int fusg_find_homography(const int32_t *src, const int32_t *dst, int n, double *H, uint8_t *ok, int N, void *stream);

/*
 * cv2.warpPerspective(img, H, (W,H)) with default flags for N images (planes_utils.py:76-77),
 * gathered from global memory (any H,W >= 1).   img [N,H,W,3] u8, Hm [N,9] f64 -> out [N,H,W,3] u8.
 */
int fusg_warp_perspective(const uint8_t *img, const double *Hm, uint8_t *out, int N, int H, int W, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FUSG_H_ */
