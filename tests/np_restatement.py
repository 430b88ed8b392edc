"""Second, independent restatement of the hot path in vectorised numpy (float32 ops are IEEE, the
fused multiply-add is emulated exactly through float64 + round-to-odd).  Used by the CPU tests to
cross-check oracle/oracle.c; cites the same reference lines as the oracle."""
from __future__ import annotations

import numpy as np

F = np.float32


def fma32(a, b, c):
    """Correctly rounded float32 fma(a, b, c) for float32 arrays."""
    a64, b64, c64 = (np.asarray(v, dtype=np.float64) for v in (a, b, c))
    p = a64 * b64                      # exact: 24 + 24 significant bits
    s = p + c64                        # RN in float64
    bb = s - p                         # TwoSum error term (exact)
    e = (p - (s - bb)) + (c64 - bb)
    # round-to-odd: if inexact and the float64 significand is even, step one ulp towards the error
    bits = s.view(np.int64) if isinstance(s, np.ndarray) else np.asarray(s).view(np.int64)
    even = (bits & 1) == 0
    toward = np.where(e > 0, np.inf, -np.inf)
    s_odd = np.where((e != 0) & even & np.isfinite(s), np.nextafter(s, toward), s)
    return s_odd.astype(F)


def geometry(sw, sh, dw, dh, aspect):
    """fk::Resize::build, resize.cuh:100-161,191-216."""
    def rnd(x):  # cxp::round
        x = F(x)
        return F(int(x + F(0.5))) if x > 0 else F(int(x - F(0.5)))
    tw, th = dw, dh
    if aspect != 1:
        sf = F(dh) / F(sh)
        wt = int(rnd(sf * F(sw)))
        even = aspect == 2
        if even:
            wt -= wt % 2
        if wt > dw:
            sf2 = F(dw) / F(sw)
            ht = int(rnd(sf2 * F(sh)))
            if even:
                ht -= ht % 2
            tw, th = dw, ht
        else:
            tw, th = wt, dh
    fx = F(1.0 / (float(tw) / float(sw)))
    fy = F(1.0 / (float(th) / float(sh)))
    if aspect == 1:
        return fx, fy, 0, 0, dw - 1, dh - 1
    x1 = 0 if aspect == 3 else (dw - tw) // 2
    y1 = (dh - th) // 2
    return fx, fy, x1, y1, x1 + tw - 1, y1 + th - 1


def resize_crop(img3, dw, dh, aspect, bg):
    """img3: [h, w, 3] uint8 crop. Returns [dh, dw, 3] float32 (Resize::exec + Interpolate::exec,
    resize.cuh:70-82,178-189; interpolation.cuh:57-92; FP order = reference SASS)."""
    h, w, _ = img3.shape
    fx, fy, bx1, by1, bx2, by2 = geometry(w, h, dw, dh, aspect)
    out = np.empty((dh, dw, 3), dtype=F)
    out[:] = np.asarray(bg, dtype=F)
    xs = np.arange(bx1, bx2 + 1)
    ys = np.arange(by1, by2 + 1)
    sx = (xs - bx1).astype(F) * fx
    sy = (ys - by1).astype(F) * fy
    x1 = np.floor(sx).astype(np.int64)
    y1 = np.floor(sy).astype(np.int64)
    x2r = np.minimum(x1 + 1, w - 1)
    y2r = np.minimum(y1 + 1, h - 1)
    wx1 = sx - x1.astype(F)
    wx0 = (x1 + 1).astype(F) - sx
    wy1 = sy - y1.astype(F)
    wy0 = (y1 + 1).astype(F) - sy
    w00 = wx0[None, :] * wy0[:, None]
    w10 = wx1[None, :] * wy0[:, None]
    w01 = wx0[None, :] * wy1[:, None]
    w11 = wx1[None, :] * wy1[:, None]
    f = img3.astype(F)
    p00 = f[y1][:, x1]
    p10 = f[y1][:, x2r]
    p01 = f[y2r][:, x1]
    p11 = f[y2r][:, x2r]
    t = p10 * w10[..., None]
    t = fma32(p00, np.broadcast_to(w00[..., None], p00.shape), t)
    t = fma32(p01, np.broadcast_to(w01[..., None], p01.shape), t)
    t = fma32(p11, np.broadcast_to(w11[..., None], p11.shape), t)
    out[by1:by2 + 1, bx1:bx2 + 1] = t
    return out


def chain(v, ops, fused=True, round_u8=False):
    """v: [..., 3] float32. ops as in tests.util (kind, values)."""
    v = v.astype(F)
    if round_u8:
        v = np.where(v > 0, np.minimum(np.rint(v), F(255)), F(0)).astype(F)
    i = 0
    while i < len(ops):
        k, val = ops[i]
        if k == "mul" and fused:
            j = i + 1
            perm = [0, 1, 2]
            while j < len(ops) and ops[j][0] == "reorder":
                perm = [perm[c] for c in ops[j][1]]
                j += 1
            if j < len(ops) and ops[j][0] in ("sub", "add"):
                m = np.asarray(val, dtype=F)
                a = np.asarray(ops[j][1], dtype=F)
                if ops[j][0] == "sub":
                    a = -a
                src = v[..., perm]
                v = fma32(src, np.broadcast_to(m[perm], src.shape), np.broadcast_to(a, src.shape))
                i = j + 1
                continue
        if k == "reorder":
            v = v[..., list(val)]
        else:
            c = np.asarray(val, dtype=F)
            v = {"mul": v * c, "sub": v - c, "add": v + c, "div": v / c}[k].astype(F)
        i += 1
    return v


def preproc(image, width, rects, dsize, ops, aspect=1, bg=(0, 0, 0), n_planes=None, used=None, fused=True,
            round_u8=False):
    """Whole batch -> NCHW float32 [n_planes, 3, H, W]."""
    dw, dh = dsize
    n_planes = len(rects) if n_planes is None else n_planes
    used = len(rects) if used is None else used
    out = np.empty((n_planes, 3, dh, dw), dtype=F)
    for z in range(n_planes):
        if z < used:
            x, y, w, h = rects[z]
            crop = image[y:y + h, 3 * x:3 * (x + w)].reshape(h, w, 3)
            r = resize_crop(crop, dw, dh, aspect, bg)
        else:
            r = np.broadcast_to(np.asarray(bg, dtype=F), (dh, dw, 3))
        out[z] = np.moveaxis(chain(r, ops, fused, round_u8), -1, 0)
    return out
