#!/usr/bin/env python
"""bench.py -- fwd+bwd renders/sec of the Gaussian rasterizer hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|cpu]
                    [--workload config3|room50k|tum3m] [--gaussians P]

One "step" = one GaussianRasterizer forward + one backward with a dense random dL/dcolor over one
synthetic view.  N>1 (under torchrun) = N independent replicas, one per GPU ("replicas only": a
single render does not shard); value is the whole-job aggregate.
--impl reference times the UNMODIFIED reference CUDA extension (baseline/_ref) on the same GPU -- the
reference has no CPU implementation of this path; --impl cpu times the C oracle on the host cores.
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import scenes  # noqa: E402

METRIC = "fwd+bwd renders/sec @1M Gaussians/1200x680"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "cpu"])
    ap.add_argument("--workload", default="config3",
                    choices=["config3", "replica50k", "splatam816k", "tum3m", "room50k", "room3m", "config1"])
    ap.add_argument("--gaussians", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mapping", action="store_true")
    ap.add_argument("--eager", action="store_true", help="ours: time the synchronous eager operator instead of graph replay")
    return ap.parse_args()


def make_scene(args):
    if args.workload == "config3":
        return scenes.config3(P=args.gaussians or 1_000_000), "config3: synthetic isotropic Gaussians, Replica intrinsics 1200x680, seed 2"
    if args.workload == "replica50k":      # BASELINE config[1] geometry: ~50k Gaussians that fill a 1200x680 view
        return scenes.view_filling(seed=12, P=args.gaussians or 50_000, cover=True), \
            "replica50k: 50k isotropic Gaussians covering the view (splat size = pixel spacing), Replica intrinsics 1200x680, seed 12"
    if args.workload == "splatam816k":     # SplaTAM's own first-frame map: one Gaussian per pixel (splatam.py:196-203)
        return scenes.view_filling(seed=12, P=args.gaussians or None), \
            "splatam816k: one isotropic Gaussian per pixel (sigma 1 px, opacity ~0.5), Replica intrinsics 1200x680, seed 12"
    if args.workload == "tum3m":           # BASELINE config[4] geometry: 3M anisotropic Gaussians, all in view
        return scenes.view_filling(seed=15, P=args.gaussians or 3_000_000, cam=scenes.TUM_FR1, anisotropic=True), \
            "tum3m: 3M anisotropic Gaussians filling the view, TUM fr1 intrinsics 640x480, seed 15"
    if args.workload == "room50k":         # round-1 workloads (97 % of the Gaussians outside the frustum), kept for continuity
        return scenes.room(P=args.gaussians or 50_000), "room: ~50k isotropic Gaussians, Replica intrinsics 1200x680, seed 1"
    if args.workload == "room3m":
        return scenes.room(seed=9, P=args.gaussians or 3_000_000, cam=scenes.TUM_FR1, anisotropic=True), \
            "room3m: anisotropic Gaussians, TUM fr1 intrinsics 640x480, seed 9 (~10 % in view)"
    return scenes.config1(), "config1: 256 Gaussians 64x64"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line).  NVML is
    polled from a thread every 20 ms (the regions here last 0.1-0.5 s, too short for `nvidia-smi -lms`, whose
    process start alone can outlast them); nvidia-smi is the fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.sm, self.mx, self.reasons = index, [], None, set()
        self.proc, self.thread, self.stop_flag, self.nvml, self.rows = None, None, threading.Event(), None, []
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
                r = int(get(self.handle))
                for name, b in bits.items():
                    if r & b:
                        self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def start(self):
        self.stop_flag.clear()
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def pause(self):
        """End of one timed region (sampling resumes with the next start())."""
        if self.nvml is not None and self.thread is not None:
            self.stop_flag.set()
            self.thread.join()
            self.thread = None

    def stop(self):
        if self.nvml is not None:
            self.pause()
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml, 20 ms poll"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 100"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def reference_extension():
    """The UNMODIFIED reference CUDA extension installed under baseline/_ref, or None."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "diff_gaussian_rasterization")):
        return None
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        import diff_gaussian_rasterization as ref
        return ref
    except Exception:
        return None


def get_ops(impl):
    if impl == "ours":
        import splatam_b200 as S
        return S.GaussianRasterizer, S.GaussianRasterizationSettings
    ref = reference_extension()
    if ref is None:
        return None, None
    return ref.GaussianRasterizer, ref.GaussianRasterizationSettings


def gpu_step_fn(scene, dev, Rast, Settings, graph=False):
    """Returns (step, inputs, eager_step, info).  graph=True (ours only): the operator's sync-free mode
    (max_rendered = 1.25 x num_rendered) with forward+backward captured once into a CUDA graph and replayed --
    every replay does the full projection/sort/blend/backward work on the HBM-resident inputs, but the step no
    longer depends on host speed (no num_rendered read-back, 1 launch instead of ~25)."""
    rs = scene.settings(Settings, dev)
    rast = Rast(rs)
    inp = scene.inputs(dev, requires_grad=True)
    g = torch.Generator().manual_seed(3)
    dL = torch.randn(3, scene.h, scene.w, generator=g).to(dev)
    leaves = [inp[k] for k in ["means3D", "means2D", "colors_precomp", "opacities", "scales", "rotations"]]

    def eager():
        for t in leaves:
            t.grad = None
        color, radii, depth = rast(**inp)
        color.backward(dL)
        return color
    if not graph:
        return eager, inp, eager, {"mode": "eager (num_rendered read back every forward)"}
    color = eager()
    torch.cuda.synchronize(dev)
    R = int(color.grad_fn.state.num_rendered)
    cap = int(R * 1.25) + 4096
    rast_sf = Rast(rs, max_rendered=cap)
    # fresh leaves for the captured path: their AccumulateGrad nodes must be born on the capture side stream
    # (the eager leaves above were first used on the default stream, which a capture may not touch)
    inp_g = scene.inputs(dev, requires_grad=True)
    leaves_g = [inp_g[k] for k in ["means3D", "means2D", "colors_precomp", "opacities", "scales", "rotations"]]

    def sync_free():
        for t in leaves_g:
            t.grad = None
        c, _, _ = rast_sf(**inp_g)
        c.backward(dL)
        return c
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            sync_free()
    torch.cuda.current_stream(dev).wait_stream(side)
    for t in leaves_g:
        t.grad = None
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        c, _, _ = rast_sf(**inp_g)
        c.backward(dL)
    info = {"mode": "sync-free operator (max_rendered=%d = 1.25 x num_rendered) replayed from a CUDA graph" % cap,
            "max_rendered": cap, "check": rast_sf.last_counts,
            "keepalive": (inp_g, dL, c, cg)}     # the graph's static inputs/outputs must outlive this function
    return cg.replay, inp, eager, info


def e2e_host_buffers(scene):
    host = {k: v.clone().pin_memory() for k, v in dict(means3D=scene.means3D, colors_precomp=scene.colors,
                                                      opacities=scene.opacities, scales=scene.scales,
                                                      rotations=scene.rotations).items()}
    g = torch.Generator().manual_seed(3)
    host_dL = torch.randn(3, scene.h, scene.w, generator=g).pin_memory()
    h2d = sum(t.numel() * 4 for t in host.values()) + host_dL.numel() * 4
    d2h = 4 * scene.h * scene.w * 4
    return host, host_dL, h2d, d2h


def e2e_step_fn(scene, dev, Rast, Settings, max_rendered=None):
    """Public-API call with HOST buffers.  Every step copies the five input tensors and dL/dcolor from pinned host
    memory into one of two device buffer sets (H2D stream), renders forward+backward from that set, and reads the
    colour and depth images back to pinned host memory (D2H stream, overlapping the backward).  The two buffer
    sets let step i+1's upload overlap step i's compute, as an input pipeline would; every step's copies are
    inside the timed region.  Returns (step, h2d, d2h, begin, finish, note)."""
    rs = scene.settings(Settings, dev)
    rast = Rast(rs) if max_rendered is None else Rast(rs, max_rendered=max_rendered)   # ours: sync-free operator
    host, host_dL, h2d, d2h = e2e_host_buffers(scene)
    out_color = torch.empty(3, scene.h, scene.w).pin_memory()
    out_depth = torch.empty(1, scene.h, scene.w).pin_memory()
    sets = [dict({k: torch.empty_like(v, device=dev) for k, v in host.items()}, dL=torch.empty_like(host_dL, device=dev))
            for _ in range(2)]
    up, down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    free = [torch.cuda.Event() for _ in range(2)]
    state = {"i": 0}

    def step():
        cur = torch.cuda.current_stream(dev)
        j = state["i"] & 1
        state["i"] += 1
        buf = sets[j]
        with torch.cuda.stream(up):
            up.wait_event(free[j])                      # compute of step i-2 has finished reading this set
            for k, v in host.items():
                buf[k].copy_(v, non_blocking=True)
            buf["dL"].copy_(host_dL, non_blocking=True)
            ready = torch.cuda.Event(); ready.record(up)
        cur.wait_event(ready)
        inp = {k: buf[k].detach().requires_grad_(True) for k in host}
        inp["means2D"] = torch.zeros_like(inp["means3D"], requires_grad=True)
        color, radii, depth = rast(**inp)
        fwd_done = torch.cuda.Event(); fwd_done.record(cur)
        with torch.cuda.stream(down):
            down.wait_event(fwd_done)
            out_color.copy_(color.detach(), non_blocking=True)
            out_depth.copy_(depth.detach(), non_blocking=True)
            d2h_done = torch.cuda.Event(); d2h_done.record(down)
        color.backward(buf["dL"])
        cur.wait_event(d2h_done)        # the step ends when both the gradients and the host images are complete
        free[j].record(cur)
    note = ("per step: 5 input tensors + dL/dcolor copied from pinned host memory (double-buffered upload stream), "
            "colour+depth images read back to pinned host memory; eager operator calls")
    return step, h2d, d2h, None, None, note


def e2e_graph_step_fn(scene, dev, Rast, Settings, max_rendered):
    """The same end-to-end step (pinned-host inputs -> H2D -> forward+backward -> D2H of the images) through the
    operator's sync-free mode, with the WHOLE step -- the six H2D copies, the operator's kernels and the two D2H
    copies -- captured into a CUDA graph, one graph per buffer set.  The two graphs are replayed alternately on
    two streams, so step i+1's upload overlaps step i's compute exactly as in the eager harness, but the step
    costs the host one graph launch instead of ~40 enqueues.  Every step's copies are inside the timed region."""
    rs = scene.settings(Settings, dev)
    rast = Rast(rs, max_rendered=max_rendered)
    host, host_dL, h2d, d2h = e2e_host_buffers(scene)
    lanes = []
    for j in range(2):
        st, down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        buf = dict({k: torch.empty_like(v, device=dev) for k, v in host.items()}, dL=torch.empty_like(host_dL, device=dev))
        out_color = torch.empty(3, scene.h, scene.w).pin_memory()
        out_depth = torch.empty(1, scene.h, scene.w).pin_memory()
        keep = {}

        def body(buf=buf, down=down, out_color=out_color, out_depth=out_depth, keep=keep):
            cur = torch.cuda.current_stream(dev)
            for k, v in host.items():
                buf[k].copy_(v, non_blocking=True)
            buf["dL"].copy_(host_dL, non_blocking=True)
            inp = {k: buf[k].detach().requires_grad_(True) for k in host}
            inp["means2D"] = torch.zeros_like(inp["means3D"], requires_grad=True)
            color, radii, depth = rast(**inp)
            down.wait_stream(cur)
            with torch.cuda.stream(down):
                out_color.copy_(color.detach(), non_blocking=True)
                out_depth.copy_(depth.detach(), non_blocking=True)
            color.backward(buf["dL"])
            cur.wait_stream(down)
            keep.update(inp=inp, color=color, depth=depth, radii=radii)
        st.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(st):
            for _ in range(2):
                body()
        torch.cuda.synchronize(dev)
        keep.clear()
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg, stream=st):
            body()
        lanes.append(dict(stream=st, graph=cg, keep=keep, buf=buf, out=(out_color, out_depth)))
    state = {"i": 0}

    def begin():        # both lanes start after everything already on the current stream (the start event)
        cur = torch.cuda.current_stream(dev)
        for ln in lanes:
            ln["stream"].wait_stream(cur)

    def step():
        ln = lanes[state["i"] & 1]
        state["i"] += 1
        with torch.cuda.stream(ln["stream"]):
            ln["graph"].replay()

    def finish():       # the current stream (the stop event) waits for both lanes
        cur = torch.cuda.current_stream(dev)
        for ln in lanes:
            cur.wait_stream(ln["stream"])
    note = ("per step: 5 input tensors + dL/dcolor copied from pinned host memory, colour+depth images read back to "
            "pinned host memory; the whole step (copies + sync-free operator forward+backward) is one CUDA graph "
            "per buffer set, two sets replayed alternately on two streams so uploads overlap compute")
    return step, h2d, d2h, begin, finish, note, lanes


def timed(step, steps, warmup, dev, dist_on, sampler=None, begin=None, finish=None):
    """W untimed steps, then exactly K steps between CUDA events on the current stream, bracketed by barrier +
    synchronize; max over ranks.  begin/finish: hooks that fork / join side streams inside the timed region."""
    import torch.distributed as dist
    for _ in range(warmup):
        step()
    if finish:
        finish()
    torch.cuda.synchronize(dev)
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize(dev)
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if begin:
        begin()
    for _ in range(steps):
        step()
    if finish:
        finish()
    e1.record()
    torch.cuda.synchronize(dev)
    if sampler:
        sampler.pause()
    if dist_on:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if dist_on:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def mapping_problem(dev, Settings, P, nframes=8):
    """The mapping workload: a view-filling SplaTAM-style map (scenes.view_filling: one Gaussian per pixel of a
    1200x680 frame + back-projections of later keyframes, every Gaussian in view, R ~ 2.3 P) and `nframes` keyframes
    at slightly different poses."""
    sc = scenes.view_filling(seed=12, P=P)
    cam = sc.settings(Settings, dev)
    g = torch.Generator().manual_seed(0)
    gauss = dict(means3D=sc.means3D, rgb_colors=sc.colors, unnorm_rotations=sc.rotations,
                 logit_opacities=torch.logit(sc.opacities.clamp(0.02, 0.98)), log_scales=torch.log(sc.scales[:, :1]))
    gauss = {k: v.to(dev).contiguous() for k, v in gauss.items()}
    rots = torch.zeros(1, 4, nframes); rots[:, 0] = 1.0
    rots[:, 1:] = 0.002 * torch.randn(1, 3, nframes, generator=g)
    trans = 0.01 * torch.randn(1, 3, nframes, generator=g)
    vv, uu = torch.meshgrid(torch.arange(sc.h, dtype=torch.float32), torch.arange(sc.w, dtype=torch.float32), indexing="ij")
    depth = scenes._surface_depth(uu, vv)[None]
    frames = [dict(id=t, cam=cam, w2c=torch.eye(4, device=dev), im=torch.rand(3, sc.h, sc.w, generator=g).to(dev),
                   depth=(depth * (1.0 + 0.01 * torch.randn(1, sc.h, sc.w, generator=g))).to(dev)) for t in range(nframes)]
    return sc, gauss, rots.to(dev), trans.to(dev), frames


def mapping_bench(dev, world, dist_on, impl, steps=50, warmup=5, P=1_000_000):
    """Secondary BASELINE metric: mapping keyframe-iterations/sec.  One iteration = SplaTAM's
    get_loss(mapping=True) (2 raster fwd + 2 raster bwd + glue + L1/SSIM) + one Adam step.
    ours: a K-rank step renders K keyframes (one per GPU), all-reduces the packed per-Gaussian gradient bucket over
    NCCL and applies the same fused Adam update on every rank.
    reference: the STOCK mapping inner loop of R/scripts/splatam.py:828-885 -- the unmodified get_loss (:214-347),
    transform_to_frame / rendervars (R/utils/slam_helpers.py), calc_ssim (R/utils/slam_external.py) and
    initialize_optimizer (:160-166), imported from baseline/_ref/SplaTAM, over the unmodified reference extension;
    single GPU (the reference has no multi-GPU path), so its value is per process."""
    Rast, Settings = get_ops(impl)
    sc, gauss, rots, trans, frames = mapping_problem(dev, Settings, P)
    note = ("SplaTAM get_loss(mapping=True) + Adam per keyframe (RGB and depth/silhouette renders, L1+SSIM, masked depth "
            "L1) on a view-filling map (scenes.view_filling: every Gaussian in view)")
    if impl == "ours":
        from splatam_b200 import mapping as M
        mapper = M.ShardedMapper(gauss, rots, trans, seed=11, fused=True)
        mapper.enable_graph(frames)       # loss fwd+bwd of a keyframe as one CUDA graph over the sync-free rasterizer
        ms = timed(lambda: mapper.step(frames), steps, warmup, dev, dist_on) / steps
        n_r, overflow = mapper.check_capacity()
        assert not overflow, "sync-free capacity overflowed: the mapping numbers would be invalid"
        visible = int((mapper.g.seen_f > 0).sum().item())
        out = dict(value=world * 1000.0 / ms, keyframes_per_step=world, num_rendered=int(n_r), visible=visible,
                   allreduce_bytes=int(mapper.g.bucket.numel() * 4) if world > 1 else 0, impl_note=note +
                   "; fused glue + fused two-set render + fused loss + fused Adam, step replayed from a CUDA graph; "
                   "NCCL all-reduce of the packed gradient bucket when n_gpus > 1")
        out.update(mapper.timing_breakdown(frames) if hasattr(mapper, "timing_breakdown") else {})
    else:
        import refsrc
        if not refsrc.available():
            return {"unavailable": "reference Python (baseline/_ref/SplaTAM) not installed"}
        R = refsrc.load(reference_extension())
        params = {k: torch.nn.Parameter(v.clone().contiguous()) for k, v in dict(gauss, cam_unnorm_rots=rots, cam_trans=trans).items()}
        lrs = dict(means3D=0.0001, rgb_colors=0.0025, unnorm_rotations=0.001, logit_opacities=0.05, log_scales=0.001,
                   cam_unnorm_rots=0.0, cam_trans=0.0)                 # R/configs/replica/splatam.py:92-100
        opt = R.splatam.initialize_optimizer(params, lrs, tracking=False)
        variables = dict(max_2D_radius=torch.zeros(sc.P, device=dev), means2D_gradient_accum=torch.zeros(sc.P, device=dev),
                         denom=torch.zeros(sc.P, device=dev))
        rng = np.random.RandomState(11)
        stats = {}

        def step():
            fr = frames[rng.randint(0, len(frames))]
            data = dict(cam=fr["cam"], im=fr["im"], depth=fr["depth"], id=fr["id"], intrinsics=None, w2c=fr["w2c"],
                        iter_gt_w2c_list=None)
            loss, var, losses = R.splatam.get_loss(params, data, variables, fr["id"], dict(im=0.5, depth=1.0), False, 0.5,
                                                   True, False, mapping=True)
            loss.backward()
            with torch.no_grad():
                opt.step()
                opt.zero_grad(set_to_none=True)
            stats["seen"] = var["seen"]
        ms = timed(step, steps, warmup, dev, dist_on) / steps
        out = dict(value=1000.0 / ms, keyframes_per_step=1, visible=int(stats["seen"].sum().item()), allreduce_bytes=0,
                   impl_note=note + "; stock get_loss / initialize_optimizer of R/scripts/splatam.py imported unmodified, "
                   "reference extension, eager PyTorch; per process (the reference maps on one GPU)")
    return dict(dict(metric="mapping keyframe-iters/sec", unit="keyframe-iters/s", ms_per_step=ms, steps=steps, warmup=warmup,
                     gaussians=P, width=sc.w, height=sc.h, workload=sc.name), **out)


def tracking_bench(dev, impl, iters=40, P=50_000):
    """BASELINE config[1]: the tracking-only inner loop (camera-only Adam, silhouette-masked L1 sums, 40 iterations per
    frame: R/configs/replica/splatam.py:59-80) on a ~50k-Gaussian map that covers a 1200x680 view.
    ours: mapping.track_frame over the fused path (fused glue with the pose gradient, two-set render, fused masked L1).
    reference: the STOCK loop of R/scripts/splatam.py:690-738 -- unmodified get_loss(tracking=True) +
    initialize_optimizer(tracking=True) over the unmodified reference extension.  Single GPU (tracking does not shard)."""
    Rast, Settings = get_ops(impl)
    sc = scenes.view_filling(seed=12, P=P, cover=True, opacity=(0.85, 0.95))
    cam = sc.settings(Settings, dev)
    gauss = dict(means3D=sc.means3D, rgb_colors=sc.colors, unnorm_rotations=sc.rotations,
                 logit_opacities=torch.logit(sc.opacities.clamp(0.02, 0.98)), log_scales=torch.log(sc.scales[:, :1]))
    gauss = {k: v.to(dev).contiguous() for k, v in gauss.items()}
    from splatam_b200 import slam, mapping as M
    rots, trans = slam.look_trajectory(2, dev)
    render = None if impl == "ours" else (lambda settings, **rv: Rast(raster_settings=settings)(**rv))
    target = slam.render_frame(gauss, rots, trans, 1, cam, render)
    target["depth"] = torch.where(target["sil"] > 0.9, target["depth"] / target["sil"].clamp(min=1e-6), torch.zeros_like(target["depth"]))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def fresh_pose():
        r = torch.zeros(1, 4, 2, device=dev); r[:, 0] = 1.0
        return r, torch.zeros(1, 3, 2, device=dev)
    if impl == "ours":
        def run():
            r, t = fresh_pose()
            p = dict({k: v.detach() for k, v in gauss.items()}, cam_unnorm_rots=r, cam_trans=t)
            # graph=True: a tracking-only run over a map that stays put re-uses one captured iteration frame after frame
            return M.track_frame(p, target, num_iters=iters, fused=True, graph=True)
    else:
        import refsrc
        if not refsrc.available():
            return {"unavailable": "reference Python (baseline/_ref/SplaTAM) not installed"}
        R = refsrc.load(reference_extension())
        lrs = dict(means3D=0.0, rgb_colors=0.0, unnorm_rotations=0.0, logit_opacities=0.0, log_scales=0.0,
                   cam_unnorm_rots=0.0004, cam_trans=0.002)               # R/configs/replica/splatam.py:71-79
        data = dict(cam=cam, im=target["im"], depth=target["depth"], id=1, intrinsics=None, w2c=torch.eye(4, device=dev),
                    iter_gt_w2c_list=None)

        def run():
            r, t = fresh_pose()
            params = {k: torch.nn.Parameter(v.clone().contiguous()) for k, v in dict(gauss, cam_unnorm_rots=r, cam_trans=t).items()}
            variables = dict(max_2D_radius=torch.zeros(sc.P, device=dev), means2D_gradient_accum=torch.zeros(sc.P, device=dev),
                             denom=torch.zeros(sc.P, device=dev))
            opt = R.splatam.initialize_optimizer(params, lrs, tracking=True)
            best, losses = float(1e20), []
            for _ in range(iters):                                        # splatam.py:690-712
                loss, variables, _ = R.splatam.get_loss(params, data, variables, 1, dict(im=0.5, depth=1.0), True, 0.99, True,
                                                        False, tracking=True)
                loss.backward()
                opt.step()
                opt.zero_grad(set_to_none=True)
                with torch.no_grad():
                    if loss < best:
                        best = loss
                        cand = (params["cam_unnorm_rots"][..., 1].detach().clone(), params["cam_trans"][..., 1].detach().clone())
                losses.append(float(loss))
            return losses
    run()                                                                # warm-up frame
    torch.cuda.synchronize(dev)
    e0.record()
    losses = run()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / iters
    return dict(metric="tracking iters/sec", value=1000.0 / ms, unit="iters/s", ms_per_iter=ms, iters=iters, gaussians=sc.P,
                width=sc.w, height=sc.h, workload=sc.name, first_loss=losses[0], last_loss=losses[-1],
                note="config 2: tracking-only inner loop on a 50k-Gaussian view-covering map; " +
                     ("fused path (mapping.track_frame)" if impl == "ours" else
                      "stock get_loss(tracking=True) + initialize_optimizer(tracking=True), reference extension"))


def cpu_oracle_run(scene, budget_s=25.0):
    """Times the C oracle (fwd render OpenMP over tiles, backward single-threaded double accumulation)
    on a bounded sample: the same view with the first P_s Gaussians, P_s chosen so one fwd+bwd stays
    within ~budget_s; reports renders/sec of that sample and the sample description."""
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    P_s = min(scene.P, 1_000_000)
    sub = scenes.Scene(scene.name + "_cpu", scene.w, scene.h, scene.fx, scene.fy, scene.cx, scene.cy,
                       scene.means3D[:P_s], scene.colors[:P_s], scene.opacities[:P_s], scene.scales[:P_s],
                       scene.rotations[:P_s])
    g = torch.Generator().manual_seed(3)
    dL = torch.randn(3, scene.h, scene.w, generator=g).numpy()
    t0 = time.perf_counter()
    o = sub.oracle()
    o.render()
    o.backward(dL)
    dt = time.perf_counter() - t0
    return dict(value=1.0 / dt, unit="renders/s", cores=cores, kind="port",
                sample=f"1 fwd+bwd render of the first {P_s} of {scene.P} Gaussians at {scene.w}x{scene.h} "
                       f"(R={o.R}); C oracle, OpenMP forward + sequential backward, {dt:.2f} s"), P_s, o.R


def cpu_bruteforce_run(scene, P_s, max_elems=16_000_000):
    """BASELINE.json's "brute-force PyTorch alpha-composite on the host cores": the dense [P_s, pixels] float32
    composite of oracle/bruteforce_torch.py with torch.autograd as the backward, walked over the full image in row
    bands (pixels are independent; the Gaussian gradients accumulate across bands), on the first P_s Gaussians of
    the workload.  Reports renders/s of that sample and (Gaussian, pixel) pair evaluations per second."""
    from oracle import bruteforce_torch as BF
    torch.set_num_threads(os.cpu_count() or 1)
    names = ["means3D", "colors", "opacities", "scales", "rotations"]
    vals = [scene.means3D[:P_s], scene.colors[:P_s], scene.opacities[:P_s], scene.scales[:P_s], scene.rotations[:P_s]]
    leaves = [v.clone().float().requires_grad_(True) for v in vals]
    g = torch.Generator().manual_seed(3)
    dL = torch.randn(3, scene.h, scene.w, generator=g)
    band = max(1, min(scene.h, max_elems // (P_s * scene.w)))
    t0 = time.perf_counter()
    for r0 in range(0, scene.h, band):
        r1 = min(scene.h, r0 + band)
        out = BF.render(*leaves, width=scene.w, height=scene.h, tanfovx=scene.w / (2 * scene.fx),
                        tanfovy=scene.h / (2 * scene.fy), bg=scene.bg, viewmatrix=scene.view[0], projmatrix=scene.proj[0],
                        dtype=torch.float32, rows=(r0, r1))
        (out["color"] * dL[:, r0:r1]).sum().backward()
    dt = time.perf_counter() - t0
    return dict(gaussians=P_s, renders_per_s=1.0 / dt, seconds=dt, pair_evals_per_s=P_s * scene.w * scene.h / dt,
                cores=os.cpu_count() or 1, rows_per_band=band)


def main():
    args = parse()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ and args.impl != "cpu":
        # convenience: `python bench.py --gpus N` re-launches itself one rank per GPU (the driver uses torchrun)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 400)] + sys.argv
        sys.exit(subprocess.call(cmd))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    scene, wl_desc = make_scene(args)
    base = dict(metric=METRIC, unit="renders/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic")

    if args.impl == "cpu":
        if rank != 0:
            return
        cb, P_s, R_s = cpu_oracle_run(scene)
        # the brute-force PyTorch composite is O(P x pixels): timed on small prefixes of the workload, never extrapolated
        # BASELINE.md B2 names P in {256, 4096, 65536}: the first two are timed; 65536 Gaussians x 816k pixels is 5.3e10 pair
        # evaluations (hours on these host cores) and is reported as not run rather than extrapolated
        cb["bruteforce_torch"] = [cpu_bruteforce_run(scene, n) for n in (256, 4096) if n <= scene.P]
        cb["bruteforce_torch_not_run"] = {"gaussians": 65536, "pair_evals": 65536 * scene.w * scene.h,
                                          "why": "beyond the time budget of one bench invocation; no extrapolation reported"}
        line = dict(base, impl="cpu", value=cb["value"], ms_per_step=1000.0 / cb["value"], n_gpus=0,
                    config=dict(workload=wl_desc, sample=cb["sample"]), cpu_baseline=cb,
                    e2e=dict(value=cb["value"], unit="renders/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    gpu_launches=0)
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback on the product path)"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist_on = world > 1
    if dist_on:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    Rast, Settings = get_ops(args.impl)
    if Rast is None:
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref (reference CUDA extension) not built"}))
        return

    step, inp, eager_step, mode = gpu_step_fn(scene, dev, Rast, Settings, graph=(args.impl == "ours" and not args.eager))
    color = eager_step()
    torch.cuda.synchronize(dev)
    R = int(color.grad_fn.num_rendered) if args.impl == "reference" else int(color.grad_fn.state.num_rendered)

    sampler = ClockSampler(local)
    ms = timed(step, args.steps, args.warmup, dev, dist_on, sampler=sampler)
    if "check" in mode:
        n_r, overflow = mode["check"]()
        assert not overflow and n_r == R, f"sync-free capacity check failed: num_rendered={n_r} overflow={overflow} expected={R}"
    ms_per_step = ms / args.steps
    value = world * 1000.0 / ms_per_step

    # end-to-end through the public API with host buffers
    e_steps = max(args.steps // 2, 4)
    if "max_rendered" in mode:
        estep, h2d, d2h, begin, finish, enote, lanes = e2e_graph_step_fn(scene, dev, Rast, Settings, mode["max_rendered"])
    else:
        estep, h2d, d2h, begin, finish, enote = e2e_step_fn(scene, dev, Rast, Settings)
    # median of three timed regions: the end-to-end leg depends on the host (graph launches, PCIe), which varies more
    # from box to box than the device-resident metric
    e_runs = sorted(timed(estep, e_steps, 4, dev, dist_on, sampler=sampler, begin=begin, finish=finish) / e_steps
                    for _ in range(3))
    ems = e_runs[1]
    clocks = sampler.stop()
    eager_ms = None
    if args.impl == "ours" and not args.eager:
        # what UNMODIFIED SplaTAM gets: Renderer(raster_settings=cam)(**rendervar) -> the synchronous operator with
        # the num_rendered read-back every forward (no capacity argument, no CUDA graph)
        eager_ms = timed(eager_step, max(args.steps // 2, 4), 3, dev, dist_on) / max(args.steps // 2, 4)
    if "max_rendered" in mode:
        n_r, overflow = Rast.last_counts()
        assert not overflow and n_r == R, f"e2e sync-free capacity check failed: {n_r} {overflow} expected {R}"
    e2e = dict(value=world * 1000.0 / ems, unit="renders/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
               ms_per_step=ems, steps=e_steps, note=enote,
               # the copies are real: sustained host<->device rate they imply per GPU (PCIe Gen5 x16 ~ 55 GB/s per direction)
               h2d_gbs=h2d / (ems * 1e-3) / 1e9, d2h_gbs=d2h / (ems * 1e-3) / 1e9)

    b_algo = scenes.algorithmic_bytes(scene.P, R, scene.w, scene.h)
    peak, peak_src = load_peaks()
    e2e["runs_ms_per_step"] = [round(x, 4) for x in e_runs]
    line = dict(base, value=value, ms_per_step=ms_per_step, e2e=e2e, clocks=clocks)
    if eager_ms is not None:
        line["eager_drop_in"] = dict(value=world * 1000.0 / eager_ms, unit="renders/s", ms_per_step=eager_ms,
                                     note="the synchronous operator exactly as unmodified SplaTAM calls it (num_rendered "
                                          "read back every forward, ~25 launches, no CUDA graph); host-speed dependent")
    line["config"] = dict(workload=wl_desc, gaussians=scene.P, width=scene.w, height=scene.h, num_rendered=R,
                          parallelism=f"replicas x{world}", mode=mode["mode"],
                          l2="per-step working set (inputs 56 B/Gaussian + geometry/binning/record workspaces, "
                             "> 250 MB at 1M Gaussians) exceeds the 126 MB L2; no explicit flush",
                          algorithmic_bytes_per_render=b_algo)
    if args.impl == "reference":
        line.update(impl="reference",
                    cpu_baseline=dict(value=value, unit="renders/s", cores=0, kind="reference",
                                      sample="the reference is itself a CUDA extension: this arm runs the unmodified "
                                             "diff_gaussian_rasterization._C from baseline/_ref on the same B200 "
                                             "(no CPU implementation of this path exists in the reference)"),
                    roofline=dict(bound="hbm", achieved=b_algo / (ms_per_step * 1e-3) / 1e9, peak=peak, unit="GB/s",
                                  frac=b_algo / (ms_per_step * 1e-3) / 1e9 / peak, traffic=None,
                                  scope="whole fwd+bwd render, B_algo of SURVEY.md 8(d)", peak_source=peak_src),
                    gpu_launches=0)
    else:
        # per-stage device timing pass (CUDA events around every stage launch, same K steps)
        from splatam_b200 import _lib
        lib = _lib.load()
        lib.sb_profile_begin()
        for _ in range(args.steps):
            eager_step()
        ms_arr, calls = (ctypes.c_float * 10)(), (ctypes.c_int * 10)()
        lib.sb_profile_end(ms_arr, calls)
        stages = {lib.sb_stage_name(i).decode(): (ms_arr[i] / max(calls[i], 1)) for i in range(10)}
        top = max(stages, key=stages.get)
        HW = scene.w * scene.h
        V = int((inp["means3D"].grad.abs().sum(1) != 0).sum().item()) if inp["means3D"].grad is not None else scene.P
        stage_bytes = {  # algorithmic bytes per launch
            "project": 44 * scene.P + 4 * scene.P,
            "depth_sort": 4 * 16 * scene.P, "depth_scan": 8 * scene.P,
            "emit_instances": 12 * R, "tile_sort": 2 * 24 * R, "ranges_records": 32 * R + 12 * R,
            "blend_forward": 32 * R + 24 * HW,
            "accum_zero": 36 * scene.P,
            "blend_backward": 28 * R + 20 * HW + 36 * V,
            "geometry_backward": 56 * scene.P + 36 * scene.P + 68 * scene.P,
        }
        ach = stage_bytes[top] / (stages[top] * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed
        # ncu --set full capture of this exact workload (profiles/r02_traffic.json, written by tools/ncu_traffic.py from
        # the .ncu-rep); null for any other workload
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if args.workload == "config3" and scene.P == 1_000_000 and os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("ours", {}).get(top, {}).get("dram_bytes")
        line["roofline"] = dict(bound="hbm", kernel=top, achieved=ach, peak=peak, unit="GB/s", frac=ach / peak,
                                traffic=traffic, peak_source=peak_src, kernel_ms=stages[top],
                                algorithmic_bytes_per_launch=stage_bytes[top],
                                render=dict(achieved=b_algo / (ms_per_step * 1e-3) / 1e9,
                                            frac=b_algo / (ms_per_step * 1e-3) / 1e9 / peak,
                                            note="whole fwd+bwd render, B_algo of SURVEY.md 8(d)"),
                                stage_ms={k: round(v, 4) for k, v in stages.items()})
        own = 15 if args.eager else 16
        line["gpu_launches"] = own * args.steps
        line["config"]["gpu_launches_note"] = (
            "%d hand-written kernels per step: project, depth sort (radix_hist + 4 radix_onesweep passes), emit_instances, "
            "tile sort (radix_hist + 2 passes), ranges_records, blend_forward, blend_backward, geometry_backward%s; the CUB "
            "inclusive scan and 3 memsets are not counted" % (own, "" if args.eager else ", finalize_count of the sync-free mode"))
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], _, _ = cpu_oracle_run(scene)
    if not args.no_mapping:
        try:
            line["mapping"] = mapping_bench(dev, world, dist_on, args.impl)
        except Exception as e:  # never lose the headline line
            line["mapping"] = {"error": repr(e)[:200]}
        try:
            line["tracking"] = tracking_bench(dev, args.impl)
        except Exception as e:
            line["tracking"] = {"error": repr(e)[:200]}
    if rank == 0:
        print(json.dumps(line))
    if dist_on:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
