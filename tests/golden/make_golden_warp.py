"""Golden vectors for the warp path, produced by the REFERENCE's own kernel (oracle/_ref/libfkref_16.so, fk::Warping)
on a GPU box:   python tests/golden/make_golden_warp.py gpurun_out/   then copy warp_*.npz into tests/golden/.
Small on purpose (a 72x56 image, 48x40 destinations)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import cvgpuspeedup_b200 as cvgs  # noqa: E402
from tests import gpu_util, util  # noqa: E402
from tests.test_warp_gpu import perspective_from_points  # noqa: E402


def main(out_dir):
    rng = np.random.default_rng(2025)
    w, h, pitch = 72, 56, 256
    img = util.make_image(rng, w, h, pitch)
    cases = {
        "affine_shift": (cvgs.WARP_AFFINE, np.array([[1, 0, 5], [0, 1, 10]], dtype=np.float64)),
        "affine_rot": (cvgs.WARP_AFFINE, np.array([[0.9, 0.35, -6.5], [-0.3, 1.05, 4.25]])),
        "persp": (cvgs.WARP_PERSPECTIVE, perspective_from_points([(9, 11), (61, 8), (5, 50), (66, 52)],
                                                               [(0, 0), (48, 0), (0, 40), (48, 40)])),
    }
    mul = (0.5, 1.25, 1 / 255.0)
    for name, (wt, m) in cases.items():
        inv = cvgs.api.invert_warp_matrix(m, wt)
        f = gpu_util.run_fkref_warp(img, w, h, wt, inv, (48, 40), mul=mul)
        u = gpu_util.run_fkref_warp(img, w, h, wt, inv, (48, 40), mul=None)
        np.savez_compressed(os.path.join(out_dir, f"warp_{name}.npz"), image=img, width=w, height=h, warp_type=wt,
                            inverse=inv, mul=np.array(mul, dtype=np.float32), out_f32=f, out_u8=u)
        print(name, f.shape, u.shape, int(np.count_nonzero(u)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out")
