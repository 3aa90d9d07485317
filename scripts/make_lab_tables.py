"""OpenCV's 8-bit RGB -> CIE L*a*b* (cv2.cvtColor(..., COLOR_RGB2LAB / COLOR_BGR2LAB) on uint8, the colour space of every ICN
input: warp_learn/models.py:354-358, warp_learn/planes_utils.py:88) restated and pinned EXHAUSTIVELY against the installed cv2.

OpenCV (opencv-python, unpinned in requirements.txt:5; 4.13.0 here) documents the integer pipeline
    R,G,B -> sRGB gamma table (x 2^3) -> XYZ/white-point matrix in 2^12 fixed point -> cube-root table (2^15) ->
    L = (296*fY - 1336935 + 2^14) >> 15,  a = (500*(fX - fY) + 128*2^15 + 2^14) >> 15,  b = (200*(fY - fZ) + ...) >> 15
but for default coefficients it evaluates a trilinearly interpolated 33^3 table of that function instead, which differs from
it by +-1 in a or b for ~1e-4 of the 2^24 colours (never in L).  This script
  1. builds the two tables of the exact pipeline in float64 and checks that pipeline against cv2 on ALL 16,777,216 colours,
  2. records every colour where cv2 deviates as an exception (key = R<<16 | G<<8 | B, value = cv2's a<<8 | b),
  3. verifies that pipeline + exceptions == cv2 on all colours, for RGB2LAB and BGR2LAB,
  4. writes future_urban_scene_generation_b200/data/lab8.npz (tables + exceptions: what the device kernel and the oracle read)."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAB_SHIFT, GAMMA_SHIFT = 12, 3
LAB_SHIFT2 = LAB_SHIFT + GAMMA_SHIFT


def descale(x, n):
    return (x + (1 << (n - 1))) >> n


def gamma(x):
    return x / 12.92 if x <= 0.04045 else ((x + 0.055) / 1.055) ** 2.4


def cbrt_lab(x):
    return x * 7.787 + 0.13793103448275862 if x < 0.008856 else x ** (1.0 / 3.0)


gamma_tab = np.array([min(65535, int(round(255.0 * (1 << GAMMA_SHIFT) * gamma(i / 255.0)))) for i in range(256)], dtype=np.int64)
n_cbrt = 256 * 3 // 2 * (1 << GAMMA_SHIFT)
cbrt_tab = np.array([min(65535, int(round((1 << LAB_SHIFT2) * cbrt_lab(i / (255.0 * (1 << GAMMA_SHIFT)))))) for i in range(n_cbrt)], dtype=np.int64)
M = np.array([[0.412453, 0.357580, 0.180423], [0.212671, 0.715160, 0.072169], [0.019334, 0.119193, 0.950227]])
WP = np.array([0.950456, 1.0, 1.088754])
COEFFS = np.array([[int(round(M[i, j] / WP[i] * (1 << LAB_SHIFT))) for j in range(3)] for i in range(3)], dtype=np.int64)
assert COEFFS.tolist() == [[1777, 1541, 778], [871, 2929, 296], [73, 448, 3575]]
LSCALE = (116 * 255 + 50) // 100
LSHIFT = -((16 * 255 * (1 << LAB_SHIFT2) + 50) // 100)
assert (LSCALE, LSHIFT) == (296, -1336934), (LSCALE, LSHIFT)


def pipeline(rgb):
    R, G, B = gamma_tab[rgb[..., 0]], gamma_tab[rgb[..., 1]], gamma_tab[rgb[..., 2]]
    fX = cbrt_tab[descale(R * COEFFS[0, 0] + G * COEFFS[0, 1] + B * COEFFS[0, 2], LAB_SHIFT)]
    fY = cbrt_tab[descale(R * COEFFS[1, 0] + G * COEFFS[1, 1] + B * COEFFS[1, 2], LAB_SHIFT)]
    fZ = cbrt_tab[descale(R * COEFFS[2, 0] + G * COEFFS[2, 1] + B * COEFFS[2, 2], LAB_SHIFT)]
    L = descale(LSCALE * fY + LSHIFT, LAB_SHIFT2)
    a = descale(500 * (fX - fY) + 128 * (1 << LAB_SHIFT2), LAB_SHIFT2)
    b = descale(200 * (fY - fZ) + 128 * (1 << LAB_SHIFT2), LAB_SHIFT2)
    return np.stack([np.clip(L, 0, 255), np.clip(a, 0, 255), np.clip(b, 0, 255)], -1).astype(np.uint8)


r, g, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
img = np.ascontiguousarray(np.stack([r, g, b], -1).reshape(4096, 4096, 3))
ref = cv2.cvtColor(img, cv2.COLOR_RGB2LAB)
out = pipeline(img)
diff = (out != ref)
assert not diff[..., 0].any(), "L must match everywhere"
bad = diff.any(-1)
keys = (img[..., 0].astype(np.uint32) << 16 | img[..., 1].astype(np.uint32) << 8 | img[..., 2].astype(np.uint32))[bad]
vals = (ref[..., 1].astype(np.uint16) << 8 | ref[..., 2].astype(np.uint16))[bad]
order = np.argsort(keys)
keys, vals = keys[order], vals[order]
print(f"cv2 {cv2.__version__}: exact pipeline matches on {100 * (1 - bad.mean()):.5f} % of 2^24 colours; {len(keys)} exceptions "
      f"(max |delta| = {np.abs(out.astype(int) - ref.astype(int)).max()})")
# 3. pipeline + exceptions == cv2 everywhere, both channel orders
fixed = out.copy().reshape(-1, 3)
flat_keys = (img[..., 0].astype(np.uint32) << 16 | img[..., 1].astype(np.uint32) << 8 | img[..., 2].astype(np.uint32)).reshape(-1)
pos = np.searchsorted(keys, flat_keys)
pos[pos >= len(keys)] = 0
hit = keys[pos] == flat_keys
fixed[hit, 1] = (vals[pos[hit]] >> 8).astype(np.uint8)
fixed[hit, 2] = (vals[pos[hit]] & 0xFF).astype(np.uint8)
assert np.array_equal(fixed.reshape(ref.shape), ref)
ref_bgr = cv2.cvtColor(np.ascontiguousarray(img[..., ::-1]), cv2.COLOR_BGR2LAB)
assert np.array_equal(ref_bgr, ref), "BGR2LAB is RGB2LAB on the swapped channels"
path = os.path.join(ROOT, "future_urban_scene_generation_b200", "data", "lab8.npz")
np.savez_compressed(path, gamma_tab=gamma_tab.astype(np.uint16), cbrt_tab=cbrt_tab.astype(np.uint16), exc_keys=keys.astype(np.uint32),
                    exc_vals=vals.astype(np.uint16), coeffs=COEFFS.astype(np.int32), cv2_version=np.array(cv2.__version__))
print("wrote", path, os.path.getsize(path), "bytes")
