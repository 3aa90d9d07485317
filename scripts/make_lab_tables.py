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
  4. does the same for the inverse, COLOR_LAB2RGB / COLOR_LAB2BGR (see the second half of this file),
  5. writes future_urban_scene_generation_b200/data/lab8.npz (tables + exceptions: what the device kernels and the oracle read)."""
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
# ---- the inverse: cv2.cvtColor(..., COLOR_LAB2BGR / COLOR_LAB2RGB) on uint8 (to_image(from_LAB=True), warp_learn/planes_utils.py:117) ------
# OpenCV's integer pipeline (Lab2RGBinteger): L -> (Y, fY) table at 2^14 (linear branch up to L*100/255 <= 8, i.e. index 20),
# a, b -> fX, fZ by fixed-point division, f -> X, Z through the abToXZ table (an integer formula, recomputed on the fly), XYZ -> linear
# RGB with 2^12 coefficients, descale by 14, clamp to the 4096-entry inverse sRGB gamma table.  Everything but the last table is
# integer arithmetic on float32-built tables and matches as computed here; the inverse gamma table depends on OpenCV's softfloat pow,
# so it is READ BACK from cv2 (for every table index all cv2 outputs agree) and the whole pipeline is then verified on all 2^24 triples.
BASE = 1 << 14
f32 = np.float32


def cdiv(a, b):
    return np.sign(a) * (np.abs(a) // b)


lab_to_yf = np.zeros(512, np.int64)
for i in range(256):
    li = f32(i) * f32(100.0) / f32(255.0)
    if i <= 20:
        yy = li / f32(903.3)
        fy = f32(7.787) * yy + f32(16.0) / f32(116.0)
    else:
        fy = (li + f32(16.0)) / f32(116.0)
        yy = fy * fy * fy
    lab_to_yf[2 * i], lab_to_yf[2 * i + 1] = int(np.rint(float(yy) * BASE)), int(np.rint(float(fy) * BASE))
MIN_AB = -8145
idx = np.arange(MIN_AB, BASE * 9 // 4 + MIN_AB, dtype=np.int64)
ab_to_xz = np.where(idx <= 3390, cdiv(idx * 108, 841) - BASE * 16 // 116 * 108 // 841, cdiv(cdiv(idx * idx, BASE) * idx, BASE))
M_INV = np.array([3.240479, -1.53715, -0.498535, -0.969256, 1.875991, 0.041556, 0.055648, -0.204043, 1.057311]).reshape(3, 3)
C_INV = np.array([[int(np.rint((1 << LAB_SHIFT) * M_INV[i, j] * WP[j])) for j in range(3)] for i in range(3)], dtype=np.int64)
assert C_INV.tolist() == [[12615, -6296, -2223], [-3773, 7684, 185], [217, -836, 4715]]


def lab2rgb_indices(lab):
    LL, aa, bb = lab[..., 0].astype(np.int64), lab[..., 1].astype(np.int64), lab[..., 2].astype(np.int64)
    yv, ify = lab_to_yf[LL * 2], lab_to_yf[LL * 2 + 1]
    adiv = ((5 * aa * 53687 + (1 << 7)) >> 13) - 128 * BASE // 500
    bdiv = ((bb * 41943 + (1 << 4)) >> 9) - 128 * BASE // 200 + 1
    x, z = ab_to_xz[ify + adiv - MIN_AB], ab_to_xz[ify - bdiv - MIN_AB]
    return [np.clip(descale(C_INV[k, 0] * x + C_INV[k, 1] * yv + C_INV[k, 2] * z, 14), 0, 4095) for k in range(3)]


ref_inv = cv2.cvtColor(img, cv2.COLOR_LAB2RGB)               # `img` doubles as the cube of all (L, a, b) triples
ind = lab2rgb_indices(img)
inv_gamma = np.full(4096, -1, np.int64)
for k in range(3):
    lo, hi = np.full(4096, 999), np.full(4096, -1)
    np.minimum.at(lo, ind[k].reshape(-1), ref_inv[..., k].reshape(-1).astype(np.int64))
    np.maximum.at(hi, ind[k].reshape(-1), ref_inv[..., k].reshape(-1).astype(np.int64))
    assert (hi >= 0).all() and np.array_equal(lo, hi), "every table index must map to ONE cv2 output"
    assert k == 0 or np.array_equal(hi, inv_gamma)
    inv_gamma = hi
xs = np.arange(4096) / 4095.0
guess = np.clip(np.rint(255.0 * np.where(xs <= 0.0031308, xs * 12.92, 1.055 * np.power(xs, 1 / 2.4) - 0.055)), 0, 255)
print(f"inverse gamma table read back from cv2: {(guess != inv_gamma).sum()} of 4096 entries differ from the float64 formula (by +-1)")
out_inv = np.stack([inv_gamma[v] for v in ind], -1).astype(np.uint8)
assert np.array_equal(out_inv, ref_inv), "Lab -> RGB pipeline must equal cv2 on all 2^24 triples"
assert np.array_equal(cv2.cvtColor(img, cv2.COLOR_LAB2BGR), ref_inv[..., ::-1]), "LAB2BGR is LAB2RGB with the channels swapped"
print("Lab -> RGB / BGR: pipeline == cv2 on all 2^24 (L, a, b) triples")

path = os.path.join(ROOT, "future_urban_scene_generation_b200", "data", "lab8.npz")
np.savez_compressed(path, gamma_tab=gamma_tab.astype(np.uint16), cbrt_tab=cbrt_tab.astype(np.uint16), exc_keys=keys.astype(np.uint32),
                    exc_vals=vals.astype(np.uint16), coeffs=COEFFS.astype(np.int32), lab_to_yf=lab_to_yf.astype(np.uint16),
                    inv_gamma=inv_gamma.astype(np.uint8), coeffs_inv=C_INV.astype(np.int32), cv2_version=np.array(cv2.__version__))
print("wrote", path, os.path.getsize(path), "bytes")
