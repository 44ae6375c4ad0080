"""Keyframe-sharded mapping step -- NEW functionality; the reference mapping
loop is single-GPU and renders ONE keyframe per Adam step (R/scripts/splatam.py:828-885).

One process per GPU.  Every rank holds the full Gaussian parameter set; in a K-rank step rank r
renders keyframe ``perm[step*K + r]`` of the selected window (2 raster calls, RGB and
depth/silhouette, exactly SplaTAM's ``get_loss(mapping=True)``: R/scripts/splatam.py:214-347), the
per-Gaussian gradients of all ranks are summed with ONE all-reduce over a flat, packed gradient
bucket (NCCL over NVLink on the box, gloo in the CPU tests), and every rank applies the same Adam
update so the replicas stay bit-identical without any parameter broadcast.  Tracking stays
single-GPU.  A K-rank step is a K-keyframe minibatch: gradients equal a 1-process accumulation over
the same K keyframes up to float summation order (tests/test_mapping_gloo.py).

The PyTorch glue below restates R/utils/slam_helpers.py:124-139,196-304 and
R/utils/slam_external.py:25-97 (transform_to_frame, rendervars, L1, SSIM); it is host-side plumbing
around the operator, not part of the rasterizer hot path.
"""
import math

import torch
import torch.nn.functional as F

GAUSSIAN_KEYS = ("means3D", "rgb_colors", "unnorm_rotations", "logit_opacities", "log_scales")


def build_rotation(q):
    """Unit-quaternion (w,x,y,z) rows -> rotation matrices (R/utils/slam_external.py:25-42)."""
    q = q / torch.sqrt((q * q).sum(-1, keepdim=True))
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)


def quat_mult(q1, q2):
    """Hamilton product, (w,x,y,z) convention (R/utils/slam_helpers.py:24-31)."""
    w1, x1, y1, z1 = q1.T
    w2, x2, y2, z2 = q2.T
    return torch.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                        w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2]).T


def transform_to_frame(params, time_idx, gaussians_grad, camera_grad):
    """World -> camera-frame Gaussians for frame `time_idx` (R/utils/slam_helpers.py:252-304)."""
    rot, tran = params["cam_unnorm_rots"][..., time_idx], params["cam_trans"][..., time_idx]
    if not camera_grad:
        rot, tran = rot.detach(), tran.detach()
    cam_rot = F.normalize(rot)
    dev = params["means3D"].device
    rel_w2c = torch.eye(4, device=dev, dtype=torch.float32)
    rel_w2c[:3, :3] = build_rotation(cam_rot)
    rel_w2c[:3, 3] = tran
    pts, unnorm = params["means3D"], params["unnorm_rotations"]
    if not gaussians_grad:
        pts, unnorm = pts.detach(), unnorm.detach()
    pts4 = torch.cat((pts, torch.ones(pts.shape[0], 1, device=dev)), dim=1)
    out = {"means3D": (rel_w2c @ pts4.T).T[:, :3]}
    if params["log_scales"].shape[1] == 1:            # isotropic: rotations irrelevant
        out["unnorm_rotations"] = unnorm
    else:
        out["unnorm_rotations"] = quat_mult(cam_rot, F.normalize(unnorm))
    return out


def _scales(params):
    ls = params["log_scales"]
    return torch.exp(torch.tile(ls, (1, 3)) if ls.shape[1] == 1 else ls)


def rgb_rendervar(params, tg):
    """R/utils/slam_helpers.py:124-139."""
    return dict(means3D=tg["means3D"], colors_precomp=params["rgb_colors"],
                rotations=F.normalize(tg["unnorm_rotations"]), opacities=torch.sigmoid(params["logit_opacities"]),
                scales=_scales(params), means2D=torch.zeros_like(params["means3D"], requires_grad=True) + 0)


def depth_sil_rendervar(params, w2c, tg):
    """colours = [z, 1, z^2] in the camera frame (R/utils/slam_helpers.py:196-249)."""
    pts = tg["means3D"]
    pts4 = torch.cat((pts, torch.ones_like(pts[:, :1])), dim=-1)
    z = (w2c @ pts4.T).T[:, 2:3]
    col = torch.cat((z, torch.ones_like(z), z * z), dim=1)
    return dict(means3D=pts, colors_precomp=col, rotations=F.normalize(tg["unnorm_rotations"]),
                opacities=torch.sigmoid(params["logit_opacities"]), scales=_scales(params),
                means2D=torch.zeros_like(params["means3D"], requires_grad=True) + 0)


_POSE_CACHE = {}


def _pose_cache_get(rots_all, trans_all, time_idx):
    """Cache of (rel_w2c, cam_rot) per (pose tensors, their versions, frame).  Keys use id(): every entry holds weak
    references to the two tensors and is only honoured while both are the very same live objects, so a recycled
    id can never return another tensor's pose."""
    key = (id(rots_all), id(trans_all), rots_all._version, trans_all._version, int(time_idx))
    hit = _POSE_CACHE.get(key)
    if hit is not None and hit[0]() is rots_all and hit[1]() is trans_all:
        return key, hit[2]
    return key, None


def _pose_cache_put(key, rots_all, trans_all, value):
    import weakref
    if len(_POSE_CACHE) > 256:
        _POSE_CACHE.clear()
    _POSE_CACHE[key] = (weakref.ref(rots_all), weakref.ref(trans_all), value)


def fused_pose_cached(params, time_idx):
    """Cached (rel_w2c, cam_rot) of a frame whose pose is not being optimised."""
    rots_all, trans_all = params["cam_unnorm_rots"], params["cam_trans"]
    key, hit = _pose_cache_get(rots_all, trans_all, time_idx)
    if hit is None:
        hit = pose_matrices(params, time_idx)
        _pose_cache_put(key, rots_all, trans_all, hit)
    return hit


def _last_num_rendered():
    from . import rasterizer
    return rasterizer._LAST_SYNC_R[0]


def pose_matrices(params, time_idx, camera_grad=False):
    """(rel_w2c [4,4], normalised camera quaternion [1,4]) of frame `time_idx` (slam_helpers.py:266-275)."""
    rot, tran = params["cam_unnorm_rots"][..., time_idx], params["cam_trans"][..., time_idx]
    if not camera_grad:
        rot, tran = rot.detach(), tran.detach()
    cam_rot = F.normalize(rot)
    rel_w2c = torch.eye(4, device=params["means3D"].device, dtype=torch.float32)
    rel_w2c[:3, :3] = build_rotation(cam_rot)
    rel_w2c[:3, 3] = tran
    return rel_w2c, cam_rot


def fused_rendervars(params, time_idx, w2c0, camera_grad, pose=None):
    """Both rendervars of one frame from ONE fused kernel (splatam_b200/prepare.py); equals
    transform_to_frame + rgb_rendervar + depth_sil_rendervar above to float rounding.  `pose` = precomputed
    (rel_w2c, cam_rot) tensors (the CUDA-graph path keeps them in static buffers)."""
    from .prepare import prepare_gaussians
    rots_all, trans_all = params["cam_unnorm_rots"], params["cam_trans"]
    dev = params["means3D"].device
    if pose is not None:
        rel_w2c, cam_rot = pose
    elif not camera_grad:     # poses are constants of the mapping loop: build each frame's matrix once
        rel_w2c, cam_rot = fused_pose_cached(params, time_idx)
    else:
        rel_w2c, cam_rot = pose_matrices(params, time_idx, camera_grad)
    means_cam, rots, opac, sc3, dcols = prepare_gaussians(params["means3D"], params["unnorm_rotations"],
                                                          params["logit_opacities"], params["log_scales"], rel_w2c,
                                                          cam_rot, w2c0)
    z2 = lambda: torch.zeros_like(params["means3D"], requires_grad=True) + 0
    rgb = dict(means3D=means_cam, colors_precomp=params["rgb_colors"], rotations=rots, opacities=opac, scales=sc3,
               means2D=z2())
    dep = dict(means3D=means_cam, colors_precomp=dcols, rotations=rots, opacities=opac, scales=sc3, means2D=z2())
    return rgb, dep


def _ssim_window(channel, device, size=11, sigma=1.5):
    g = torch.tensor([math.exp(-(x - size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(size)])
    g = (g / g.sum()).unsqueeze(1)
    return g.mm(g.t()).float()[None, None].expand(channel, 1, size, size).contiguous().to(device)


def calc_ssim(img1, img2, window_size=11):
    """R/utils/slam_external.py:66-97 (mean SSIM, 11x11 Gaussian window, sigma 1.5)."""
    ch = img1.size(-3)
    w = _ssim_window(ch, img1.device, window_size).type_as(img1)
    pad = window_size // 2
    conv = lambda t: F.conv2d(t, w, padding=pad, groups=ch)
    mu1, mu2 = conv(img1), conv(img2)
    mu1_sq, mu2_sq, mu12 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1, s2, s12 = conv(img1 * img1) - mu1_sq, conv(img2 * img2) - mu2_sq, conv(img1 * img2) - mu12
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu12 + c1) * (2 * s12 + c2)) / ((mu1_sq + mu2_sq + c1) * (s1 + s2 + c2))).mean()


def mapping_loss(params, frame, render, loss_weights=(0.5, 1.0), ignore_outlier_depth_loss=False, fused_loss=False,
                 max_rendered=None, keep=None):
    """SplaTAM get_loss(mapping=True): Gaussians get gradient, camera does not
    (R/scripts/splatam.py:214-347 with tracking=False, mapping=True, do_ba=False, use_l1=True).
    frame: dict(im [3,H,W], depth [1,H,W], cam settings, w2c [4,4] first-frame w2c, id time index).
    render(settings, **rendervar) -> (image, radii, depth)."""
    if fused_loss:      # fused glue: one kernel builds both rendervars (csrc/prepare.cu)
        rv_rgb, rv_depth = fused_rendervars(params, frame["id"], frame["w2c"], camera_grad=False,
                                            pose=frame.get("pose"))
    else:
        tg = transform_to_frame(params, frame["id"], gaussians_grad=True, camera_grad=False)
        rv_rgb, rv_depth = rgb_rendervar(params, tg), depth_sil_rendervar(params, frame["w2c"], tg)
    if keep is not None:       # the caller wants d loss / d means2D of the RGB render (densification statistic)
        rv_rgb["means2D"].retain_grad()
        keep["means2D"] = rv_rgb["means2D"]
    if fused_loss and render is default_render:      # N1: both colour sets in one raster pass
        from .rasterizer import GaussianRasterizer
        im, depth_sil, radius, _ = GaussianRasterizer(raster_settings=frame["cam"], max_rendered=max_rendered).forward_fused(
            means3D=rv_rgb["means3D"], means2D=rv_rgb["means2D"], opacities=rv_rgb["opacities"],
            colors_precomp=rv_rgb["colors_precomp"], colors_extra=rv_depth["colors_precomp"],
            scales=rv_rgb["scales"], rotations=rv_rgb["rotations"])
    else:
        im, radius, _ = render(frame["cam"], **rv_rgb)
        depth_sil, _, _ = render(frame["cam"], **rv_depth)
    if fused_loss and not ignore_outlier_depth_loss:
        from .train_ops import masked_l1        # validity-masked mean |gt - depth| in one kernel (+ one backward)
        l_depth, _ = masked_l1(depth_sil, frame["depth"], depth_mean=True)
    else:
        depth = depth_sil[0:1]
        uncertainty = (depth_sil[2:3] - depth ** 2).detach()
        mask = (frame["depth"] > 0) & (~torch.isnan(depth)) & (~torch.isnan(uncertainty))
        if ignore_outlier_depth_loss:
            err = torch.abs(frame["depth"] - depth) * (frame["depth"] > 0)
            mask = mask & (err < 10 * err.median())
        mask = mask.detach()
        l_depth = torch.abs(frame["depth"] - depth)[mask].mean()
    if fused_loss:      # one forward + one backward kernel instead of 5 depthwise convs + autograd (train_ops.cu)
        from .train_ops import image_loss
        l_im = image_loss(im, frame["im"], 0.8, 0.2)
    else:
        l_im = 0.8 * torch.abs(im - frame["im"]).mean() + 0.2 * (1.0 - calc_ssim(im, frame["im"]))
    return loss_weights[0] * l_im + loss_weights[1] * l_depth, radius


def tracking_loss(params, frame, render, sil_thres=0.99, loss_weights=(0.5, 1.0), fused=False):
    """SplaTAM get_loss(tracking=True, use_sil_for_loss=True, use_l1=True): only the camera pose of frame
    `frame["id"]` gets gradient; L1 sums over pixels with silhouette > sil_thres and valid depth
    (R/scripts/splatam.py:220-224,254-288; weights R/configs/replica/splatam.py:65-68)."""
    if fused:
        rv_rgb, rv_depth = fused_rendervars(params, frame["id"], frame["w2c"], camera_grad=True)
    else:
        tg = transform_to_frame(params, frame["id"], gaussians_grad=False, camera_grad=True)
        rv_rgb, rv_depth = rgb_rendervar(params, tg), depth_sil_rendervar(params, frame["w2c"], tg)
    if fused and render is default_render:
        from .rasterizer import GaussianRasterizer
        im, depth_sil, radius, _ = GaussianRasterizer(raster_settings=frame["cam"]).forward_fused(
            means3D=rv_rgb["means3D"], means2D=rv_rgb["means2D"], opacities=rv_rgb["opacities"],
            colors_precomp=rv_rgb["colors_precomp"], colors_extra=rv_depth["colors_precomp"],
            scales=rv_rgb["scales"], rotations=rv_rgb["rotations"])
    else:
        im, radius, _ = render(frame["cam"], **rv_rgb)
        depth_sil, _, _ = render(frame["cam"], **rv_depth)
    if fused:
        from .train_ops import masked_l1
        l_depth, l_im = masked_l1(depth_sil, frame["depth"], im, frame["im"], sil_thres=sil_thres, use_sil=True,
                                  depth_mean=False)
    else:
        depth, sil = depth_sil[0:1], depth_sil[1]
        uncertainty = (depth_sil[2:3] - depth ** 2).detach()
        mask = (frame["depth"] > 0) & (~torch.isnan(depth)) & (~torch.isnan(uncertainty)) & (sil > sil_thres)
        mask = mask.detach()
        # boolean indexing, as the reference: a NaN depth under a cleared mask bit must not reach the sum
        l_depth = torch.abs(frame["depth"] - depth)[mask].sum()
        l_im = torch.abs(frame["im"] - im)[torch.tile(mask, (3, 1, 1))].sum()
    return loss_weights[0] * l_im + loss_weights[1] * l_depth, radius


def track_frame(params, frame, render=None, num_iters=40, lr_rot=0.0004, lr_trans=0.002, fused=None, graph=None):
    """SplaTAM's tracking inner loop for one frame (R/scripts/splatam.py:676-744): Adam on the frame's camera
    quaternion and translation only, keeping the best-loss pose.  `params` holds detached Gaussian tensors and
    cam_unnorm_rots [1,4,T] / cam_trans [1,3,T] leaf tensors.  Returns the list of per-iteration losses.
    graph: the iteration -- loss forward + backward over the sync-free rasterizer, the Adam step and the best-pose
    bookkeeping, all on the device -- is captured once and replayed, so a frame costs one host synchronisation instead of
    one per iteration (`_TrackingGraph`); None = re-use a cached graph, build one only for loops of >= 100 iterations."""
    render = default_render if render is None else render
    if fused is None:
        fused = params["means3D"].is_cuda and render is default_render
    can_graph = bool(fused) and params["means3D"].is_cuda and render is default_render
    if can_graph and graph is not False:
        # default policy: always re-use a cached graph; BUILD one only where it pays -- long loops (TUM: 200 iterations
        # per frame) or on request (graph=True: tracking-only runs over a map that stays put).  Building costs about as
        # much as ~40 eager iterations (capacity probe, warm-up, capture into a fresh memory pool).
        out = _track_frame_graphed(params, frame, num_iters, lr_rot, lr_trans, build=(graph is True or num_iters >= 100))
        if out is not None:
            return out              # None: no graph, or the instance capacity overflowed mid-frame -> run eagerly
    rots, trans = params["cam_unnorm_rots"], params["cam_trans"]
    rots.requires_grad_(True); trans.requires_grad_(True)
    opt = torch.optim.Adam([{"params": [rots], "lr": lr_rot}, {"params": [trans], "lr": lr_trans}], lr=0.0, eps=1e-15)
    t = frame["id"]
    best, best_rot, best_tran, losses = None, None, None, []
    for _ in range(num_iters):
        opt.zero_grad(set_to_none=True)
        loss, _ = tracking_loss(params, frame, render, fused=fused)
        loss.backward()
        opt.step()
        with torch.no_grad():
            # as the reference (splatam.py:703-712): the candidate is the pose AFTER the step, ranked by the loss that
            # was evaluated before it
            lv = float(loss)
            if best is None or lv < best:
                best, best_rot, best_tran = lv, rots[..., t].detach().clone(), trans[..., t].detach().clone()
        losses.append(lv)
    with torch.no_grad():
        rots[..., t] = best_rot
        trans[..., t] = best_tran
    return losses


def _capture(body):
    """Captures body() on the CURRENT (non-default) stream into a CUDA graph.  capture_begin / capture_end directly: the
    torch.cuda.graph context manager also runs gc.collect() and torch.cuda.empty_cache(), which costs tens of
    milliseconds and throws the allocator's cached blocks away."""
    g = torch.cuda.CUDAGraph()
    g.capture_begin()
    try:
        body()
    finally:
        g.capture_end()
    return g


class _TrackingGraph:
    """One captured tracking iteration (loss forward + backward over the sync-free rasterizer, capturable Adam step,
    best-pose bookkeeping, all on the device) bound to ONE map (the parameter tensors' storage), camera and iteration
    count -- not to a pose column: the pose travels through two small leaf tensors.  Re-used frame after frame while the map tensors stay the same objects (tracking-only runs; SLAM phases
    without map growth): a frame then costs two small image copies, `num_iters` graph launches and one host
    synchronisation."""

    def __init__(self, params, frame, num_iters, lr_rot, lr_trans, slack=1.5):
        from .rasterizer import GaussianRasterizer
        from .train_ops import masked_l1
        dev = params["means3D"].device
        self.dev, self.t, self.num_iters = dev, frame["id"], num_iters
        t = frame["id"]
        rots, trans = params["cam_unnorm_rots"], params["cam_trans"]
        self.gp = dict({k: params[k].detach() for k in GAUSSIAN_KEYS}, cam_unnorm_rots=rots.detach(), cam_trans=trans.detach())
        self.key = _tracking_key(params, frame, num_iters)
        gp = self.gp
        with torch.no_grad():       # one synchronous forward at the starting pose sizes the capacity
            rgb, _ = fused_rendervars(gp, t, frame["w2c"], camera_grad=False, pose=pose_matrices(gp, t))
            GaussianRasterizer(frame["cam"])(means3D=rgb["means3D"], means2D=rgb["means2D"], opacities=rgb["opacities"],
                                             colors_precomp=rgb["colors_precomp"], scales=rgb["scales"],
                                             rotations=rgb["rotations"])
        cap = int(_last_num_rendered() * slack) + 4096
        self.side = torch.cuda.Stream(dev)
        self.side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.side):
            self.im, self.depth, self.w2c = frame["im"].clone(), frame["depth"].clone(), frame["w2c"].clone()
            self.q = rots[..., t].detach().clone().requires_grad_(True)      # [1,4], [1,3]: leaves born on the capture stream
            self.tr = trans[..., t].detach().clone().requires_grad_(True)
            q, tr = self.q, self.tr
            q.grad, tr.grad = torch.zeros_like(q), torch.zeros_like(tr)
            self.opt = torch.optim.Adam([{"params": [q], "lr": lr_rot}, {"params": [tr], "lr": lr_trans}], lr=0.0,
                                        eps=1e-15, capturable=True)
            self.best_loss = torch.full((), float("inf"), device=dev)
            self.best_q, self.best_tr = q.detach().clone(), tr.detach().clone()
            self.hist = torch.zeros(num_iters, device=dev)
            self.it = torch.zeros(1, dtype=torch.long, device=dev)
            self.bad_any = torch.zeros(1, device=dev)
            cam = frame["cam"]

            def body():
                q.grad.zero_(); tr.grad.zero_()
                cam_rot = F.normalize(q)
                rel = torch.eye(4, device=dev, dtype=torch.float32)
                rel[:3, :3] = build_rotation(cam_rot)
                rel[:3, 3] = tr
                rgb, dep = fused_rendervars(gp, t, self.w2c, camera_grad=True, pose=(rel, cam_rot))
                im, depth_sil, _, _ = GaussianRasterizer(cam, max_rendered=cap).forward_fused(
                    means3D=rgb["means3D"], means2D=rgb["means2D"], opacities=rgb["opacities"],
                    colors_precomp=rgb["colors_precomp"], colors_extra=dep["colors_precomp"], scales=rgb["scales"],
                    rotations=rgb["rotations"])
                l_depth, l_im = masked_l1(depth_sil, self.depth, im, self.im, sil_thres=0.99, use_sil=True, depth_mean=False)
                loss = 0.5 * l_im + 1.0 * l_depth
                loss.backward()
                lossd = loss.detach().reshape(1)
                bad = (GaussianRasterizer.last_state(dev).header()[2:3] != 0) | ~torch.isfinite(lossd)
                q.grad.copy_(torch.where(bad, torch.zeros_like(q.grad), q.grad))
                tr.grad.copy_(torch.where(bad, torch.zeros_like(tr.grad), tr.grad))
                self.bad_any.add_(bad.float())
                self.opt.step()
                # the reference's rule (splatam.py:703-712): candidate = pose AFTER the step, ranked by the loss evaluated
                # before it; the first strictly smaller loss wins
                better = (lossd < self.best_loss) & ~bad
                self.best_loss.copy_(torch.where(better, lossd, self.best_loss.reshape(1)).reshape(()))
                self.best_q.copy_(torch.where(better, q.detach(), self.best_q))
                self.best_tr.copy_(torch.where(better, tr.detach(), self.best_tr))
                self.hist.scatter_(0, self.it, lossd)
                self.it.add_(1)
            self.body = body
            for _ in range(3):              # warm-up (initialises the capturable Adam state); undone by reset()
                body()
            self.graph = _capture(body)
        torch.cuda.current_stream(dev).wait_stream(self.side)

    def reset(self, params, frame):
        """Loads a frame: images, starting pose, fresh optimizer state (the reference builds a new Adam per frame,
        splatam.py:680), empty best-pose record."""
        t = frame["id"]
        with torch.no_grad():
            self.im.copy_(frame["im"]); self.depth.copy_(frame["depth"]); self.w2c.copy_(frame["w2c"])
            self.q.copy_(params["cam_unnorm_rots"][..., t]); self.tr.copy_(params["cam_trans"][..., t])
            for st in self.opt.state.values():
                st["exp_avg"].zero_(); st["exp_avg_sq"].zero_(); st["step"].zero_()
            self.best_loss.fill_(float("inf")); self.best_q.copy_(self.q); self.best_tr.copy_(self.tr)
            self.hist.zero_(); self.it.zero_(); self.bad_any.zero_()

    def run(self, params, frame):
        dev = self.dev
        self.side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.side):
            self.reset(params, frame)
            for _ in range(self.num_iters):
                self.graph.replay()
        torch.cuda.current_stream(dev).wait_stream(self.side)
        if float(self.bad_any) != 0.0:                                   # the one host synchronisation of the frame
            return None
        with torch.no_grad():
            params["cam_unnorm_rots"][..., frame["id"]] = self.best_q
            params["cam_trans"][..., frame["id"]] = self.best_tr
        return self.hist.tolist()


_TRACK_GRAPHS = {}


def _tracking_key(params, frame, num_iters):
    return (params["means3D"].data_ptr(), tuple(params["means3D"].shape), params["log_scales"].shape[1],
            id(frame["cam"]), tuple(frame["im"].shape), num_iters)


def _track_frame_graphed(params, frame, num_iters, lr_rot, lr_trans, build=True):
    """Graph-replayed tracking of one frame, or None when no graph is available / an iteration overflowed the instance
    capacity (the caller then runs the frame eagerly).  Graphs are cached per (map storage, camera, pose column,
    iteration count); `build` creates one on a miss (worth it for long loops or when the map stays put)."""
    key = _tracking_key(params, frame, num_iters)
    tg = _TRACK_GRAPHS.get(key)
    if tg is None:
        if not build:
            return None
        if len(_TRACK_GRAPHS) >= 2:
            _TRACK_GRAPHS.clear()
        tg = _TRACK_GRAPHS[key] = _TrackingGraph(params, frame, num_iters, lr_rot, lr_trans)
    return tg.run(params, frame)


class FlatGaussians:
    """The five Gaussian parameter tensors as views into ONE flat fp32 buffer, their gradients as
    views into ONE flat gradient bucket of the same layout -- the packed send buffer of the
    all-reduce (48-56 B per Gaussian)."""

    def __init__(self, tensors):
        self.shapes = {k: tuple(tensors[k].shape) for k in GAUSSIAN_KEYS}
        n = sum(math.prod(s) for s in self.shapes.values())
        dev = tensors["means3D"].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        # ONE all-reduce bucket: [per-Gaussian gradients (n) | per-Gaussian "seen" flags as 0/1 floats (P) | loss (1) |
        # capacity-overflow flag of the sync-free rasterizer (1): non-zero on ANY rank => every rank skips the update]
        P_ = tensors["means3D"].shape[0]
        self.bucket = torch.zeros(n + P_ + 2, dtype=torch.float32, device=dev)
        self.flat_grad = self.bucket[:n]
        self.seen_f = self.bucket[n:n + P_]
        self.loss_slot = self.bucket[n + P_:n + P_ + 1]
        self.overflow_slot = self.bucket[n + P_ + 1:]
        self.params, off = {}, 0
        for k in GAUSSIAN_KEYS:
            m = math.prod(self.shapes[k])
            self.flat[off:off + m].copy_(tensors[k].detach().reshape(-1))
            p = self.flat[off:off + m].view(self.shapes[k]).requires_grad_(True)
            p.grad = self.flat_grad[off:off + m].view(self.shapes[k])   # autograd accumulates in place
            self.params[k] = p
            off += m

    def zero_grad(self):
        self.bucket.zero_()


def default_render(settings, **rendervar):
    from .rasterizer import GaussianRasterizer
    return GaussianRasterizer(raster_settings=settings)(**rendervar)


class ShardedMapper:
    """Data-parallel mapping over keyframes.  `lrs` follow R/configs/replica/splatam.py:92-100."""

    DEFAULT_LRS = dict(means3D=0.0001, rgb_colors=0.0025, unnorm_rotations=0.001, logit_opacities=0.05,
                       log_scales=0.001)

    def __init__(self, gaussians, cam_unnorm_rots, cam_trans, lrs=None, render=default_render, group=None, seed=0,
                 fused=None):
        import torch.distributed as dist
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.group = group
        self.rank = self.dist.get_rank(group) if self.dist else 0
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.g = FlatGaussians(gaussians)
        self.cam = dict(cam_unnorm_rots=cam_unnorm_rots.detach(), cam_trans=cam_trans.detach())
        lrs = dict(self.DEFAULT_LRS, **(lrs or {}))
        self.lrs = lrs
        # fused = hand-written Adam over the flat buffer + fused L1/SSIM loss (CUDA only); the torch path is
        # the reference formulation (R/scripts/splatam.py:160-166) and the one the CPU/gloo test exercises
        self.fused = self.g.flat.is_cuda if fused is None else bool(fused)
        self._make_optimizer()
        self.render = render
        self.gen = torch.Generator().manual_seed(seed)   # shared seed -> identical schedule on every rank
        self.step_idx = 0
        # graph mode, world > 1: the bucket is all-reduced in this many slices, the guarded Adam update of slice k
        # overlapping the transfer of slice k+1 (1 = one collective, then one Adam launch)
        import os as _os
        self.overlap_chunks = int(_os.environ.get("SPLATAM_OVERLAP_CHUNKS", "1"))

    def params(self):
        return dict(self.g.params, **self.cam)

    # ---- what stays single-GPU: rank `src` does it, the result is broadcast ---------------------
    def sync_camera(self, src=0):
        """Tracking runs on one rank (its backward uses float atomics, so two ranks would not end bit-identical);
        the estimated poses are broadcast so every replica maps with the same cameras."""
        if self.dist and self.world > 1:
            for k in ("cam_unnorm_rots", "cam_trans"):
                t = self.cam[k].detach().clone().contiguous()
                self.dist.broadcast(t, src=src, group=self.group)
                self.cam[k] = t
        return self.cam

    def broadcast_frame(self, frame, src=0):
        """A keyframe's images are loaded by one rank and broadcast once when the keyframe is created (13 MB at
        1200x680); `frame` must hold tensors of the right shape on every rank (contents are overwritten)."""
        if self.dist and self.world > 1:
            for k in ("im", "depth", "w2c"):
                t = frame[k].contiguous()
                self.dist.broadcast(t, src=src, group=self.group)
                frame[k] = t
        return frame

    # ---- map maintenance (prune / grow); every rank applies the same deterministic edit --------------------------
    def _make_optimizer(self, m=None, v=None, t=0):
        if self.fused:
            from .train_ops import FusedAdam
            sizes = [math.prod(self.g.shapes[k]) for k in GAUSSIAN_KEYS]
            self.opt = FusedAdam(self.g.flat, self.g.flat_grad, sizes, [self.lrs[k] for k in GAUSSIAN_KEYS], eps=1e-15,
                                 exp_avg=m, exp_avg_sq=v, step=t)
        else:
            self.opt = torch.optim.Adam([{"params": [self.g.params[k]], "name": k, "lr": self.lrs[k]}
                                         for k in GAUSSIAN_KEYS], lr=0.0, eps=1e-15)

    def reset_optimizer(self):
        """Fresh Adam state, as the reference re-creates its optimizer at the start of every frame's mapping
        phase (R/scripts/splatam.py:822)."""
        self._make_optimizer()

    def _widths(self):
        return [self.g.shapes[k][1] if len(self.g.shapes[k]) > 1 else 1 for k in GAUSSIAN_KEYS]

    def _rebuild(self, flat, P_new, m=None, v=None, t=0, state=None):
        tensors, off = {}, 0
        for k, w in zip(GAUSSIAN_KEYS, self._widths()):
            tensors[k] = flat[off:off + w * P_new].view(P_new, w)
            off += w * P_new
        self.g = FlatGaussians(tensors)
        self._make_optimizer(m, v, t)
        if state is not None:                   # torch.optim path: carry the moments over
            for k in GAUSSIAN_KEYS:
                self.opt.state[self.g.params[k]] = state[k]
        self._graph = None                      # shapes changed: a captured graph is stale

    def remove_points(self, to_remove):
        """Drop the rows where `to_remove` is set from the parameters and both Adam moments
        (remove_points, R/utils/slam_external.py:144-167).  Returns the number of Gaussians kept."""
        P = self.g.shapes["means3D"][0]
        keep = ~to_remove.reshape(-1).bool()
        if self.fused:
            from . import map_ops
            m8, dst, P_new = map_ops.compact_plan(keep)
            if P_new == P:
                return P
            w = self._widths()
            flat = map_ops.compact_flat(self.g.flat, P, w, m8, dst, P_new)
            m = map_ops.compact_flat(self.opt.m, P, w, m8, dst, P_new)
            v = map_ops.compact_flat(self.opt.v, P, w, m8, dst, P_new)
            self._rebuild(flat, P_new, m, v, self.opt.applied_steps()[0])
            return P_new
        P_new = int(keep.sum())
        if P_new == P:
            return P
        state = {}
        for k in GAUSSIAN_KEYS:
            st = self.opt.state.get(self.g.params[k], None)
            state[k] = {} if not st else dict(step=st["step"], exp_avg=st["exp_avg"][keep].clone(),
                                              exp_avg_sq=st["exp_avg_sq"][keep].clone())
        flat = torch.cat([self.g.params[k].detach()[keep].reshape(-1) for k in GAUSSIAN_KEYS])
        self._rebuild(flat, P_new, state={k: s for k, s in state.items() if s} or None)
        return P_new

    def prune_gaussians(self, iter, prune_dict, scene_radius):
        """prune_gaussians of the reference (R/utils/slam_external.py:170-197): same schedule keys
        (start_after, stop_after, prune_every, removal_opacity_threshold, final_removal_opacity_threshold,
        remove_big_after, reset_opacities, reset_opacities_every).  Returns the Gaussian count afterwards."""
        P = self.g.shapes["means3D"][0]
        if iter > prune_dict["stop_after"]:
            return P
        if iter >= prune_dict["start_after"] and iter % prune_dict["prune_every"] == 0:
            thr = (prune_dict["final_removal_opacity_threshold"] if iter == prune_dict["stop_after"]
                   else prune_dict["removal_opacity_threshold"])
            big = 0.1 * float(scene_radius) if iter >= prune_dict["remove_big_after"] else None
            lo, ls = self.g.params["logit_opacities"].detach(), self.g.params["log_scales"].detach()
            if self.fused:
                from . import map_ops
                keep = map_ops.prune_mask(lo, ls, thr, big)
            else:
                remove = (torch.sigmoid(lo) < thr).squeeze(-1)
                if big is not None:
                    remove = remove | (torch.exp(ls).max(dim=1).values > big)
                keep = ~remove
            P = self.remove_points(~keep)
        if iter > 0 and iter % prune_dict["reset_opacities_every"] == 0 and prune_dict["reset_opacities"]:
            with torch.no_grad():                                  # inverse_sigmoid(0.01), moments zeroed
                self.g.params["logit_opacities"].fill_(math.log(0.01 / 0.99))
            if self.fused:
                off = sum(math.prod(self.g.shapes[k]) for k in GAUSSIAN_KEYS[:3])
                self.opt.m[off:off + P].zero_(); self.opt.v[off:off + P].zero_()
            else:
                st = self.opt.state.get(self.g.params["logit_opacities"], None)
                if st:
                    st["exp_avg"].zero_(); st["exp_avg_sq"].zero_()
        return P

    def add_gaussians(self, new):
        """Append rows (dict of the five tensors) to the map; Adam moments of the new rows start at zero
        (cat_params_to_optimizer, R/utils/slam_external.py:121-141).  Returns the new Gaussian count."""
        P, n_new = self.g.shapes["means3D"][0], new["means3D"].shape[0]
        if n_new == 0:
            return P
        dev = self.g.flat.device
        flat = torch.cat([torch.cat([self.g.params[k].detach(), new[k].to(dev).float().reshape(n_new, -1)], 0).reshape(-1)
                          for k in GAUSSIAN_KEYS])
        if self.fused:
            def grown(buf):
                parts, off = [], 0
                for k, w in zip(GAUSSIAN_KEYS, self._widths()):
                    parts += [buf[off:off + w * P], torch.zeros(w * n_new, device=dev)]
                    off += w * P
                return torch.cat(parts)
            self._rebuild(flat, P + n_new, grown(self.opt.m), grown(self.opt.v), self.opt.applied_steps()[0])
        else:
            state = {}
            for k in GAUSSIAN_KEYS:
                st = self.opt.state.get(self.g.params[k], None)
                if st:
                    z = torch.zeros(n_new, *st["exp_avg"].shape[1:], device=dev)
                    state[k] = dict(step=st["step"], exp_avg=torch.cat([st["exp_avg"], z], 0),
                                    exp_avg_sq=torch.cat([st["exp_avg_sq"], z.clone()], 0))
            self._rebuild(flat, P + n_new, state=state or None)
        return P + n_new

    def add_new_gaussians(self, frame, time_idx, intrinsics, sil_thres):
        """add_new_gaussians of the reference (R/scripts/splatam.py:378-420): one depth/silhouette render of the
        current map at the frame's estimated pose, non-presence mask, back-projection of the selected pixels,
        append.  CUDA only.  Returns the number of Gaussians added."""
        from . import map_ops
        p = self.params()
        with torch.no_grad():
            tg = transform_to_frame(p, time_idx, gaussians_grad=False, camera_grad=False)
            depth_sil, _, _ = self.render(frame["cam"], **depth_sil_rendervar(p, frame["w2c"], tg))
            rot = F.normalize(p["cam_unnorm_rots"][..., time_idx].detach())
            curr_w2c = torch.eye(4, device=depth_sil.device)
            curr_w2c[:3, :3] = build_rotation(rot)[0]
            curr_w2c[:3, 3] = p["cam_trans"][0, :, time_idx].detach()
        new, n = map_ops.new_gaussians_from_frame(depth_sil, frame, intrinsics, curr_w2c, sil_thres,
                                                  scale_dim=self._widths()[4])
        if n:
            self.add_gaussians(new)
        return n

    # ---- CUDA-graph mode -----------------------------------------------------------------------------------
    def enable_graph(self, window, slack=1.3, capacity=None):
        """Capture loss forward + backward of one keyframe into a CUDA graph (sync-free rasterizer with a fixed
        instance capacity = slack x the largest num_rendered over `window`, or `capacity`).  Afterwards `step`
        copies the chosen keyframe into static buffers and replays the graph: ~6 launches per step instead of ~60,
        and no host synchronisation, so the step no longer depends on host speed.  Frames must share one camera.

        Safety: if the map moves or grows until a render needs more instances than the capacity, the rasterizer
        raises its device-side overflow flag and writes NaN images; the flag travels in the all-reduce bucket, the
        guarded Adam step (device-side step clock) is a no-op on every rank for that step, and `step` -- which
        polls the flag with a one-step lag, never blocking on the step in flight -- re-captures with a larger
        capacity.  An overflowed step therefore costs time, never correctness."""
        assert self.fused and self.g.flat.is_cuda and self.render is default_render
        from .rasterizer import GaussianRasterizer
        dev = self.g.flat.device
        f0 = window[0]
        if capacity is None:
            worst = 0
            probe = window if len(window) <= 3 else [window[0], window[len(window) // 2], window[-1]]
            for fr in probe:             # a few synchronous forwards size the capacity (an overflow later is safe: it
                                         # skips the step and re-captures, see the docstring)
                with torch.no_grad():
                    rgb, dep = fused_rendervars(self.params(), fr["id"], fr["w2c"], camera_grad=False)
                    c, _, _ = GaussianRasterizer(fr["cam"])(**rgb)
                worst = max(worst, _last_num_rendered())
            capacity = int(worst * slack) + 4096
        self._cap = int(capacity)
        self._slack = slack
        self._polls = []                 # (event, pinned copy of the overflow slot) of steps in flight
        self.overflow_events = getattr(self, "overflow_events", 0)
        self._pin = [torch.empty(1, dtype=torch.float32, pin_memory=True) for _ in range(8)]
        self._static = dict(id=f0["id"], cam=f0["cam"], w2c=f0["w2c"].clone(), im=f0["im"].clone(), depth=f0["depth"].clone(),
                            pose=tuple(t.clone() for t in pose_matrices(self.params(), f0["id"])))
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                self.g.zero_grad()
                loss, radius = mapping_loss(self.params(), self._static, self.render, fused_loss=True, max_rendered=self._cap)
                loss.backward()
        def body():
            self.g.bucket.zero_()
            loss, radius = mapping_loss(self.params(), self._static, self.render, fused_loss=True, max_rendered=self._cap)
            loss.backward()
            self.g.seen_f.copy_(radius > 0)
            self.g.loss_slot.copy_(loss.detach().reshape(1))
            self.g.overflow_slot.copy_(GaussianRasterizer.last_state(dev).header()[2:3])     # int32 0/1 -> float
        with torch.cuda.stream(side):
            self._graph = _capture(body)
        torch.cuda.current_stream(dev).wait_stream(side)
        return self._cap

    def _replay(self, frame):
        st = self._static
        st["im"].copy_(frame["im"], non_blocking=True)
        st["depth"].copy_(frame["depth"], non_blocking=True)
        st["w2c"].copy_(frame["w2c"], non_blocking=True)
        rel, cr = fused_pose_cached(self.params(), frame["id"])
        st["pose"][0].copy_(rel, non_blocking=True)
        st["pose"][1].copy_(cr, non_blocking=True)
        self._graph.replay()

    def check_capacity(self):
        """(num_rendered, overflowed) of the last graph replay; synchronises."""
        from .rasterizer import GaussianRasterizer
        return GaussianRasterizer.last_counts(self.g.flat.device)

    def effective_steps(self):
        """(Adam steps applied, steps skipped because a render overflowed the instance capacity); synchronises."""
        return self.opt.applied_steps() if self.fused else (self.step_idx, 0)

    def _poll_overflow(self, window, block=False):
        """Looks at the overflow flags of COMPLETED earlier steps (one-step lag: never waits for the step in flight
        unless `block`); on overflow grows the capacity from the measured instance count and re-captures."""
        hit = False
        while self._polls and (block or len(self._polls) > 1 or self._polls[0][0].query()):
            ev, host = self._polls.pop(0)
            ev.synchronize()
            hit = hit or float(host[0]) != 0.0
        if hit:
            torch.cuda.synchronize(self.g.flat.device)
            self._polls = []
            n_r, _ = self.check_capacity()
            self.overflow_events += 1
            self.enable_graph(window, slack=self._slack, capacity=max(int(n_r * self._slack), int(self._cap * 1.5)) + 4096)
        return hit

    def timing_breakdown(self, window, reps=10):
        """Device time (CUDA events, ms) of the step's collective alone -- the all-reduce of the packed bucket -- measured
        outside any timed benchmark region; {} on a single rank."""
        if not (self.dist and self.world > 1):
            return {}
        dev = self.g.flat.device
        buf = torch.zeros_like(self.g.bucket)
        for _ in range(3):
            self.dist.all_reduce(buf, op=self.dist.ReduceOp.SUM, group=self.group)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            self.dist.all_reduce(buf, op=self.dist.ReduceOp.SUM, group=self.group)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX, group=self.group)
        return {"allreduce_ms": float(ms.item()), "allreduce_algbw_gbs": buf.numel() * 4 / (float(ms.item()) * 1e-3) / 1e9}

    def schedule(self, window_size):
        """Keyframe index (into the window) rendered by each rank this step: a shared-seed permutation
        dealt round-robin, so ranks draw distinct keyframes whenever the window holds >= world frames."""
        perm = torch.randperm(window_size, generator=self.gen).tolist()
        return [perm[r % window_size] for r in range(self.world)]

    def accumulate(self, frame):
        """Local backward of one keyframe into the flat gradient bucket (no communication)."""
        keep = {} if getattr(self, "track_means2D", False) else None
        loss, radius = mapping_loss(self.params(), frame, self.render, fused_loss=self.fused, keep=keep)
        loss.backward()
        if keep is not None:      # gradient of the RGB render's 2D means: the densification statistic (splatam.py:250)
            self.last_means2D_grad = keep["means2D"].grad
            self.last_seen_local = (radius > 0)
        return loss.detach(), radius

    # ---- gradient-based densification (3DGS-style; off in every shipped SLAM config) -----------------------------
    def densify(self, iter, densify_dict, scene_radius, means2D_grad=None, seen=None, generator=None):
        """densify of the reference (R/utils/slam_external.py:191-243), same schedule keys (start_after, stop_after,
        densify_every, grad_thresh, num_to_split_into, removal_opacity_threshold, final_removal_opacity_threshold,
        remove_big_after, reset_opacities, reset_opacities_every): accumulates |d loss / d means2D| of the seen
        Gaussians, clones the small high-gradient ones, splits the large ones into num_to_split_into samples of
        their own distribution, drops transparent / oversized ones and optionally resets opacities.  Parameters and both
        Adam moments follow (new rows start with zero moments).  Every rank must call it with the same all-reduced
        statistics; `generator` (default: the mapper's shared-seed generator when distributed, else torch's global
        RNG, as the reference) draws the split samples.  Returns the Gaussian count afterwards."""
        P = self.g.shapes["means3D"][0]
        dev = self.g.flat.device
        if not hasattr(self, "_dens") or self._dens["accum"].shape[0] != P:
            self._dens = dict(accum=torch.zeros(P, device=dev), denom=torch.zeros(P, device=dev))
        if iter > densify_dict["stop_after"]:
            return P
        means2D_grad = self.last_means2D_grad if means2D_grad is None else means2D_grad
        if seen is None:       # the Gaussians THIS rank's render saw (the bucket's seen flags are already summed over ranks)
            seen = self.last_seen_local if getattr(self, "last_seen_local", None) is not None else (self.g.seen_f > 0)
        if self.dist and self.world > 1:
            # K ranks rendered K keyframes this step: the statistic of the step is the sum over the keyframes, exactly
            # what K sequential single-keyframe steps of the reference would have accumulated; summing it over the ranks
            # also keeps the replicas' clone / split decisions identical
            st = torch.stack([torch.norm(means2D_grad[:, :2], dim=-1) * seen, seen.to(torch.float32)])
            self.dist.all_reduce(st, op=self.dist.ReduceOp.SUM, group=self.group)
            self._dens["accum"] += st[0]
            self._dens["denom"] += st[1]
        else:
            self._dens["accum"][seen] += torch.norm(means2D_grad[seen, :2], dim=-1)      # accumulate_mean2d_gradient
            self._dens["denom"][seen] += 1
        if not (iter >= densify_dict["start_after"] and iter % densify_dict["densify_every"] == 0):
            return P
        grads = self._dens["accum"] / self._dens["denom"]
        grads[grads.isnan()] = 0.0
        thr = densify_dict["grad_thresh"]
        cur = lambda: {k: self.g.params[k].detach() for k in GAUSSIAN_KEYS}
        big = lambda: torch.max(torch.exp(self.g.params["log_scales"].detach()), dim=1).values
        to_clone = torch.logical_and(grads >= thr, big() <= 0.01 * scene_radius)
        self.add_gaussians({k: v[to_clone] for k, v in cur().items()})
        num = self.g.shapes["means3D"][0]
        padded = torch.zeros(num, device=dev)
        padded[:grads.shape[0]] = grads
        to_split = torch.logical_and(padded >= thr, big() > 0.01 * scene_radius)
        n = int(densify_dict["num_to_split_into"])
        p = cur()
        new = {k: v[to_split].repeat(n, 1) for k, v in p.items()}
        stds = torch.exp(p["log_scales"])[to_split].repeat(n, 3)
        if generator is None and self.world > 1:
            generator = self._device_generator()
        samples = torch.normal(mean=torch.zeros((stds.size(0), 3), device=dev), std=stds, generator=generator)
        rots = build_rotation(p["unnorm_rotations"][to_split]).repeat(n, 1, 1)
        new["means3D"] = new["means3D"] + torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1)
        new["log_scales"] = torch.log(torch.exp(new["log_scales"]) / (0.8 * n))
        self.add_gaussians(new)
        self.remove_points(torch.cat((to_split, torch.zeros(n * int(to_split.sum()), dtype=torch.bool, device=dev))))
        rthr = (densify_dict["final_removal_opacity_threshold"] if iter == densify_dict["stop_after"]
                else densify_dict["removal_opacity_threshold"])
        to_remove = (torch.sigmoid(self.g.params["logit_opacities"].detach()) < rthr).squeeze(-1)
        if iter >= densify_dict["remove_big_after"]:
            to_remove = torch.logical_or(to_remove, big() > 0.1 * scene_radius)
        P = self.remove_points(to_remove)
        self._dens = dict(accum=torch.zeros(P, device=dev), denom=torch.zeros(P, device=dev))
        if iter > 0 and iter % densify_dict["reset_opacities_every"] == 0 and densify_dict["reset_opacities"]:
            with torch.no_grad():
                self.g.params["logit_opacities"].fill_(math.log(0.01 / 0.99))
            if self.fused:
                off = sum(math.prod(self.g.shapes[k]) for k in GAUSSIAN_KEYS[:3])
                self.opt.m[off:off + P].zero_(); self.opt.v[off:off + P].zero_()
            else:
                st = self.opt.state.get(self.g.params["logit_opacities"], None)
                if st:
                    st["exp_avg"].zero_(); st["exp_avg_sq"].zero_()
        return P

    def _device_generator(self):
        if getattr(self, "_dev_gen", None) is None:
            self._dev_gen = torch.Generator(device=self.g.flat.device)
            self._dev_gen.manual_seed(12345)
        return self._dev_gen

    def step(self, window):
        """One sharded mapping step over `window` (list of keyframe dicts).  Returns the mean loss."""
        picks = self.schedule(len(window))
        graphed = getattr(self, "_graph", None) is not None
        if graphed:
            self._poll_overflow(window)
            self._replay(window[picks[self.rank]])
        else:
            self.g.zero_grad()
            loss, radius = self.accumulate(window[picks[self.rank]])
            self.g.seen_f.copy_(radius > 0)
            self.g.loss_slot.copy_(loss.reshape(1))
        chunked = graphed and self.dist is not None and self.world > 1 and self.overlap_chunks > 1
        if chunked:
            # The bucket travels as K slices, the TAIL first (it carries the seen flags, the loss and the overflow flag
            # that every later Adam chunk needs); the guarded Adam update of slice k runs on the compute stream as soon as
            # its all-reduce has landed, while NCCL moves slice k+1: only the last slice's update is exposed.
            n = self.g.flat.numel()
            K = int(self.overlap_chunks)
            cuts = [n * i // K for i in range(K)] + [n]
            works = [self.dist.all_reduce(self.g.bucket[cuts[K - 1]:], op=self.dist.ReduceOp.SUM, group=self.group,
                                          async_op=True)]
            works += [self.dist.all_reduce(self.g.bucket[cuts[i]:cuts[i + 1]], op=self.dist.ReduceOp.SUM, group=self.group,
                                           async_op=True) for i in range(K - 1)]
            works[0].wait()
            self.opt.advance_clock(self.g.overflow_slot)
            self.opt.apply_range(cuts[K - 1], n - cuts[K - 1], self.g.overflow_slot)
            for i in range(K - 1):
                works[i + 1].wait()
                self.opt.apply_range(cuts[i], cuts[i + 1] - cuts[i], self.g.overflow_slot)
        elif self.dist and self.world > 1:
            # ONE collective per step: gradients, seen flags (sum of 0/1 > 0 == OR) and the loss travel together
            self.dist.all_reduce(self.g.bucket, op=self.dist.ReduceOp.SUM, group=self.group)
        if graphed:
            if not chunked:
                self.opt.step_guarded(self.g.overflow_slot)    # no-op on every rank if any rank's render overflowed
            host = self._pin[self.step_idx % len(self._pin)]
            host.copy_(self.g.overflow_slot, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._polls.append((ev, host))
        else:
            self.opt.step()
        self.step_idx += 1
        seen = self.g.seen_f > 0
        loss = self.g.loss_slot[0] / self.world
        if graphed:
            return loss.clone(), seen, picks      # device scalar: the graph path never synchronises
        return float(loss), seen, picks
