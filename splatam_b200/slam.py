"""Slim SplaTAM-style RGB-D SLAM loop over a synthetic sequence (BASELINE.json configs 4/5 stand-in: the
Replica / TUM datasets are not available offline).  It is a HARNESS around the operator, not a
re-implementation of R/scripts/splatam.py: per frame it runs SplaTAM's tracking inner loop (camera-only Adam,
silhouette-masked L1 sums; `mapping.track_frame`) from a constant-velocity initialisation
(R/scripts/splatam.py:423-444) and then `mapping_iters` keyframe-sharded mapping steps (`ShardedMapper`) over
a sliding window of keyframes, and reports ATE-RMSE of the estimated trajectory and the PSNR of re-rendered
keyframes -- the two end metrics the reference prints (R/utils/eval_helpers.py:23-77,569-592).
Densification / pruning are out of scope; the map starts from a perturbed copy of the generating scene.
"""
import math

import torch
import torch.nn.functional as F

from . import mapping as M


def look_trajectory(num_frames, device, step=(0.012, -0.004, 0.008), rot_step=(0.003, -0.002, 0.0015)):
    """Smooth camera trajectory: per-frame relative w2c as (unnormalised quaternion [1,4,T], translation [1,3,T]);
    frame 0 is the identity (SplaTAM makes all poses relative to frame 0)."""
    rots = torch.zeros(1, 4, num_frames, device=device)
    trans = torch.zeros(1, 3, num_frames, device=device)
    rots[:, 0] = 1.0
    for t in range(num_frames):
        w = 1.0 + 0.15 * math.sin(0.7 * t)
        rots[0, 1:, t] = torch.tensor(rot_step, device=device) * t * w
        trans[0, :, t] = torch.tensor(step, device=device) * t * w
    return rots, trans


def render_frame(gauss, rots, trans, t, cam, render=None):
    """RGB-D observation of frame t from the generating scene (no gradient)."""
    with torch.no_grad():
        p = dict(gauss, cam_unnorm_rots=rots, cam_trans=trans)
        tg = M.transform_to_frame(p, t, gaussians_grad=False, camera_grad=False)
        w2c0 = torch.eye(4, device=rots.device)
        render = M.default_render if render is None else render
        im, _, _ = render(cam, **M.rgb_rendervar(p, tg))
        ds, _, _ = render(cam, **M.depth_sil_rendervar(p, w2c0, tg))
    return dict(id=t, cam=cam, w2c=w2c0, im=im.clone(), depth=ds[0:1].clone())


def psnr(a, b):
    mse = ((a - b) ** 2).mean()
    return float(20 * torch.log10(1.0 / torch.sqrt(mse)))


def run_slam(gauss_init, frames, cam, render=None, tracking_iters=20, mapping_iters=8, keyframe_every=2,
             window=4, fused=None, seed=0):
    """Tracks every frame and maps on keyframes.  gauss_init: dict of the five Gaussian tensors (the map's
    starting point); frames: list of dict(id, cam, w2c, im, depth).  Returns dict(rots, trans, psnr, gauss)."""
    dev = gauss_init["means3D"].device
    T = len(frames)
    rots = torch.zeros(1, 4, T, device=dev); rots[:, 0] = 1.0
    trans = torch.zeros(1, 3, T, device=dev)
    kw = {} if render is None else {"render": render}
    mapper = M.ShardedMapper(gauss_init, rots, trans, seed=seed, fused=fused, **kw)
    keyframes = [frames[0]]
    for t in range(1, T):
        with torch.no_grad():           # constant-velocity initialisation (splatam.py:423-444)
            if t > 1:
                r1, r2 = F.normalize(rots[0, :, t - 1], dim=0), F.normalize(rots[0, :, t - 2], dim=0)
                rots[0, :, t] = F.normalize(r1 + (r1 - r2), dim=0)
                trans[0, :, t] = trans[0, :, t - 1] + (trans[0, :, t - 1] - trans[0, :, t - 2])
            else:
                rots[0, :, t] = rots[0, :, t - 1]
                trans[0, :, t] = trans[0, :, t - 1]
        params = dict({k: v.detach() for k, v in mapper.g.params.items()}, cam_unnorm_rots=rots, cam_trans=trans)
        M.track_frame(params, frames[t], render=render, num_iters=tracking_iters, fused=fused)
        rots, trans = params["cam_unnorm_rots"].detach(), params["cam_trans"].detach()
        mapper.cam = dict(cam_unnorm_rots=rots, cam_trans=trans)
        if t % keyframe_every == 0:
            keyframes.append(frames[t])
            win = keyframes[-window:]
            for _ in range(mapping_iters):
                mapper.step(win)
    # PSNR of the keyframes re-rendered from the final map at the estimated poses
    final = {k: v.detach() for k, v in mapper.g.params.items()}
    vals = []
    for fr in keyframes:
        est = render_frame(final, rots, trans, fr["id"], cam, render)
        vals.append(psnr(est["im"].clamp(0, 1), fr["im"].clamp(0, 1)))
    return dict(rots=rots, trans=trans, psnr=sum(vals) / len(vals), gauss=final)


def ate_rmse(rots_est, trans_est, rots_gt, trans_gt):
    """RMSE of camera-centre positions (all poses share frame 0, so no alignment is needed)."""
    def centres(rots, trans):
        out = []
        for t in range(rots.shape[-1]):
            R = M.build_rotation(F.normalize(rots[..., t]))[0]
            out.append(-(R.T @ trans[0, :, t]))
        return torch.stack(out)
    d = centres(rots_est, trans_est) - centres(rots_gt, trans_gt)
    return float(torch.sqrt((d ** 2).sum(-1).mean()))
