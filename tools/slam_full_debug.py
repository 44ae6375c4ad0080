"""Diagnostics for the full RGB-D SLAM loop (tests/test_slam_gpu.py::test_full_loop_from_rgbd_only): per-variant ATE,
PSNR, Gaussian counts and the silhouette coverage after the first-frame mapping.  Run on the GPU box."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import splatam_b200 as S
from splatam_b200 import mapping as M, slam
from test_slam_gpu import _opaque_sequence

dev = torch.device("cuda:0")
cam, frames, rots_gt, trans_gt = _opaque_sequence(dev, S.GaussianRasterizationSettings)
K = torch.tensor([[160.0, 0, 159.5], [0, 160.0, 95.5], [0, 0, 1]])
print("observed frame0: sil>0.99 frac %.3f  sil mean %.3f  valid depth frac %.3f  depth range %.2f..%.2f" %
      (float((frames[0]["sil"] > 0.99).float().mean()), float(frames[0]["sil"].mean()),
       float((frames[0]["depth"] > 0).float().mean()), float(frames[0]["depth"].min()), float(frames[0]["depth"].max())))
init, radius = slam.initialize_map(frames[0], K)
print("P0", init["means3D"].shape[0], "scene_radius", radius)

# coverage after first-frame mapping
for iters in (0, 20, 60):
    m = M.ShardedMapper(init, torch.tensor([[[1.0], [0], [0], [0]]], device=dev).repeat(1, 1, 9),
                        torch.zeros(1, 3, 9, device=dev), fused=True)
    for _ in range(iters):
        loss, _, _ = m.step([frames[0]])
    with torch.no_grad():
        p = m.params()
        tg = M.transform_to_frame(p, 0, False, False)
        ds, _, _ = M.default_render(cam, **M.depth_sil_rendervar(p, torch.eye(4, device=dev), tg))
        im, _, _ = M.default_render(cam, **M.rgb_rendervar(p, tg))
    print("first-frame mapping iters %3d: sil>0.99 frac %.3f  sil mean %.3f  depth L1 %.4f  psnr %.2f  opacity mean %.3f" %
          (iters, float((ds[1] > 0.99).float().mean()), float(ds[1].mean()),
           float((ds[0] - frames[0]["depth"][0]).abs().mean()), slam.psnr(im.clamp(0, 1), frames[0]["im"]),
           float(torch.sigmoid(p["logit_opacities"]).mean())))

prune = dict(start_after=0, remove_big_after=0, stop_after=20, prune_every=4, removal_opacity_threshold=0.005,
             final_removal_opacity_threshold=0.005, reset_opacities=False, reset_opacities_every=500)
travelled = float(trans_gt[0, :, -1].norm())
variants = {
    "test config": dict(add_new_gaussians=True, prune_dict=prune, select_keyframes=True, first_frame_iters=60),
    "track80": dict(add_new_gaussians=True, prune_dict=prune, select_keyframes=True, first_frame_iters=60, tracking_iters=80),
    "track40": dict(add_new_gaussians=True, prune_dict=prune, select_keyframes=True, first_frame_iters=60, tracking_iters=40),
    "no grow": dict(add_new_gaussians=False, prune_dict=None, select_keyframes=False, first_frame_iters=60),
}
for name, kw in variants.items():
    torch.manual_seed(3); np.random.seed(3)
    t0 = time.time()
    out = slam.run_slam(init, frames, cam, fused=True, intrinsics=K, sil_thres=0.5, scene_radius=radius, window=4,
                        mapping_iters=12, **kw)
    def centre(r, t, i):
        R = M.build_rotation(torch.nn.functional.normalize(r[..., i]))[0]
        return -(R.T @ t[0, :, i])
    errs = [float((centre(out["rots"], out["trans"], i) - centre(rots_gt, trans_gt, i)).norm()) for i in range(9)]
    print("%-18s ATE %.4f (%.1f%% of path) PSNR %.2f counts %s  per-frame err mm %s  %.1fs" %
          (name, slam.ate_rmse(out["rots"], out["trans"], rots_gt, trans_gt), 100 * slam.ate_rmse(out["rots"], out["trans"], rots_gt, trans_gt) / travelled,
           out["psnr"], out["counts"], [round(1000 * e, 1) for e in errs], time.time() - t0))
