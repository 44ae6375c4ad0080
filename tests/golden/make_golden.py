"""Generates the golden fixtures under tests/golden/ by running the UNMODIFIED reference extension
(baseline/_ref, built from /root/reference/diff-gaussian-rasterization-w-depth.git by
__graft_entry__.build_reference()) on a B200.  Run on the GPU box:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'

then copy gpurun_out/golden/*.npz into tests/golden/.  The reference repository ships no golden
vectors of its own, so these pin the oracle and the CUDA path.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scenes  # noqa: E402
from gpu_harness import GRADS, random_dL, run_ref  # noqa: E402


def dump(scene, path, w2c=None):
    dL = random_dL(scene)
    r = run_ref(scene, dL)
    d = dict(name=scene.name, w=scene.w, h=scene.h, fx=scene.fx, fy=scene.fy, cx=scene.cx, cy=scene.cy,
             means3D=scene.means3D.numpy(), colors=scene.colors.numpy(), opacities=scene.opacities.numpy(),
             scales=scene.scales.numpy(), rotations=scene.rotations.numpy(), bg=scene.bg.numpy(),
             viewmatrix=scene.view.numpy(), projmatrix=scene.proj.numpy(),
             w2c=scene.view[0].T.contiguous().numpy(), dL_dcolor=dL.astype(np.float32))
    for k in ["color", "depth", "radii", "depths", "means2D", "conic_opacity", "cov3D", "tiles_touched", "keys",
              "point_list", "ranges", "final_T", "n_contrib"]:
        d["ref_" + k] = r[k]
    for k in GRADS:
        d["ref_grad_" + k] = r["grad_" + k]
    np.savez_compressed(path, **d)
    print("wrote", path, "P", scene.P, "R", r["num_rendered"])


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    dump(scenes.config1(), os.path.join(out, "config1_p256_64x64.npz"))
    dump(scenes.edge_cases(), os.path.join(out, "edge_p300_97x45.npz"))
    dump(scenes.dense_opaque(P=600, w=64, h=48), os.path.join(out, "dense_p600_64x48.npz"))
