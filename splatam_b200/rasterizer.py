"""Host-side mirror of the reference operator API.

Same names, kwargs, return tuple and error behaviour as
``diff_gaussian_rasterization`` (reference: X/diff_gaussian_rasterization/__init__.py:17-196), but
every stage runs in the hand-written sm_100a kernels behind ``libsplatam_b200.so``; torch is only
used for device memory, the current stream and autograd bookkeeping.
"""
import ctypes
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _lib


class GaussianRasterizationSettings(NamedTuple):
    # field-for-field the reference's NamedTuple (__init__.py:134-145)
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool


def _ptr(t):
    """Device pointer, or NULL for the reference's "not provided" empty tensor (__init__.py:173-183)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32c(t, name):
    if t is None or t.numel() == 0:
        return t
    if not t.is_cuda:
        raise _lib.SplatamB200Error(f"{name} must be a CUDA tensor (there is no CPU fallback)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class _SettingsPack:
    """sb_settings plus the tensors that keep its device pointers alive."""

    def __init__(self, rs: GaussianRasterizationSettings, device):
        self.bg = _f32c(rs.bg.to(device), "bg")
        self.view = _f32c(rs.viewmatrix.to(device), "viewmatrix")
        self.proj = _f32c(rs.projmatrix.to(device), "projmatrix")
        self.campos = _f32c(rs.campos.to(device), "campos") if rs.campos is not None else None
        self.c = _lib.SbSettings(
            int(rs.image_height), int(rs.image_width), float(rs.tanfovx), float(rs.tanfovy),
            self.bg.data_ptr(), float(rs.scale_modifier), self.view.data_ptr(), self.proj.data_ptr(),
            int(rs.sh_degree), _ptr(self.campos), int(bool(rs.prefiltered)))


_PACK_CACHE = {}


def _settings_pack(rs, device):
    """sb_settings for a GaussianRasterizationSettings tuple, cached: SplaTAM builds the camera once per
    resolution (R/utils/recon_helpers.py:4-27) and passes the same tuple to every call.  The cache entry
    keeps the tuple alive (so ids stay unique) and is invalidated if a tensor is modified in place."""
    key = (id(rs), str(device))
    ver = (rs.bg._version, rs.viewmatrix._version, rs.projmatrix._version,
           rs.campos._version if isinstance(rs.campos, torch.Tensor) else -1)
    hit = _PACK_CACHE.get(key)
    if hit is not None and hit[0] is rs and hit[1] == ver:
        return hit[2]
    if len(_PACK_CACHE) > 64:
        _PACK_CACHE.clear()
    pack = _SettingsPack(rs, device)
    _PACK_CACHE[key] = (rs, ver, pack)
    return pack


_SIZE_CACHE = {}
_R_HINT = {}
_LAST_ASYNC_STATE = {}      # device index -> state of the most recent sync-free forward on that device
_LAST_SYNC_R = [0]          # num_rendered of the most recent synchronous forward


def _sizes(lib, P, W, H, sets):
    key = (P, W, H, sets)
    hit = _SIZE_CACHE.get(key)
    if hit is None:
        n = ctypes.c_size_t(0)
        _lib.check(lib.sb_geometry_workspace_bytes(P, ctypes.byref(n)), "sb_geometry_workspace_bytes")
        geom = n.value
        _lib.check(lib.sb_image_workspace_bytes(W, H, ctypes.byref(n)), "sb_image_workspace_bytes")
        img = n.value
        _lib.check(lib.sb_backward_workspace_bytes_ex(P, sets, ctypes.byref(n)), "sb_backward_workspace_bytes_ex")
        if len(_SIZE_CACHE) > 256:
            _SIZE_CACHE.clear()
        hit = _SIZE_CACHE[key] = (geom, img, n.value)
    return hit


class _on_device:
    """torch.cuda.device(dev) only when dev is not already current (the context manager costs ~10 us)."""

    def __init__(self, device):
        self.ctx = None if torch.cuda.current_device() == device.index else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _State:
    """Scratch kept alive between forward and backward (the reference saves its three byte buffers
    on the autograd ctx, __init__.py:82-84)."""
    __slots__ = ("geom", "binning", "image", "num_rendered", "settings", "P")

    def header(self):
        """int32 view [num_rendered, num_visible, overflow flag] of the device-side counters (no synchronisation)."""
        return self.geom[:12].view(torch.int32)

    def counts(self):
        """(num_rendered, overflowed) of a sync-free forward; synchronises the current stream."""
        lib = _lib.load()
        r, o = ctypes.c_int(0), ctypes.c_int(0)
        _lib.check(lib.sb_read_counts(self.geom.data_ptr(), self.geom.numel(), self.P, ctypes.byref(r), ctypes.byref(o),
                                      _stream(self.geom.device)), "sb_read_counts")
        return r.value, bool(o.value)


def _forward_impl(means3D, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings,
                  colors2=None, capacity=None):
    lib = _lib.load()
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        # rasterize_points.cu:56-58
        raise _lib.SplatamB200Error("means3D must have dimensions (num_points, 3)")
    device = means3D.device
    if not means3D.is_cuda:
        raise _lib.SplatamB200Error("means3D must be a CUDA tensor (there is no CPU fallback)")
    P = means3D.shape[0]
    H, W = int(raster_settings.image_height), int(raster_settings.image_width)
    means3D = _f32c(means3D, "means3D")
    colors_precomp = _f32c(colors_precomp, "colors_precomp")
    opacities = _f32c(opacities, "opacities")
    scales = _f32c(scales, "scales")
    rotations = _f32c(rotations, "rotations")
    cov3Ds_precomp = _f32c(cov3Ds_precomp, "cov3D_precomp")
    colors2 = _f32c(colors2, "colors_extra")
    sets = 2 if colors2 is not None else 1

    with _on_device(device):
        pack = _settings_pack(raster_settings, device)
        st = _stream(device)
        color = torch.empty((3, H, W), dtype=torch.float32, device=device)
        depth = torch.empty((1, H, W), dtype=torch.float32, device=device)
        radii = torch.empty((P,), dtype=torch.int32, device=device)
        state = _State()
        state.settings = pack
        state.P = P
        n = ctypes.c_size_t(0)
        geom_bytes, img_bytes, _ = _sizes(lib, P, W, H, sets)
        state.geom = _ws(geom_bytes, device)
        state.image = _ws(img_bytes, device)
        R = ctypes.c_int(0)
        color2 = torch.empty((3, H, W), dtype=torch.float32, device=device) if sets == 2 else None
        if capacity is not None:
            # sync-free mode: fixed-capacity binning workspace, num_rendered stays on the device (CUDA-graph safe)
            capacity = int(capacity)
            _lib.check(lib.sb_binning_workspace_bytes_ex(capacity, W, H, sets, ctypes.byref(n)),
                       "sb_binning_workspace_bytes_ex")
            state.binning = _ws(n.value, device)
            _lib.check(lib.sb_forward_async(
                ctypes.byref(pack.c), P, _ptr(means3D), _ptr(opacities), _ptr(scales), _ptr(rotations),
                _ptr(cov3Ds_precomp), _ptr(colors_precomp), _ptr(colors2), _ptr(radii), state.geom.data_ptr(),
                state.geom.numel(), state.binning.data_ptr(), state.binning.numel(), capacity, state.image.data_ptr(),
                state.image.numel(), color.data_ptr(), _ptr(color2), depth.data_ptr(), st), "sb_forward_async")
            state.num_rendered = capacity
            _LAST_ASYNC_STATE[device.index] = state
            saved = (means3D, colors_precomp, scales, rotations, cov3Ds_precomp)
            if sets == 2:
                return color, color2, radii, depth, state, saved + (colors2,)
            return color, radii, depth, state, saved
        # binning workspace sized from the last num_rendered seen for this problem shape (+12 % slack): the
        # whole forward is then ONE library call and the GPU only waits for the num_rendered read-back itself
        hint_key = (P, W, H, sets, device.index)
        hint = _R_HINT.get(hint_key, 0)
        if hint > 0:
            _lib.check(lib.sb_binning_workspace_bytes_ex(int(hint * 1.125) + 4096, W, H, sets, ctypes.byref(n)),
                       "sb_binning_workspace_bytes_ex")
            state.binning = _ws(n.value, device)
            bin_ptr, bin_bytes = state.binning.data_ptr(), state.binning.numel()
        else:
            state.binning, bin_ptr, bin_bytes = None, None, 0
        rc = lib.sb_forward(ctypes.byref(pack.c), P, _ptr(means3D), _ptr(opacities), _ptr(scales), _ptr(rotations),
                            _ptr(cov3Ds_precomp), _ptr(colors_precomp), _ptr(colors2), _ptr(radii),
                            state.geom.data_ptr(), state.geom.numel(), bin_ptr, bin_bytes, state.image.data_ptr(),
                            state.image.numel(), color.data_ptr(), _ptr(color2), depth.data_ptr(), ctypes.byref(R), st)
        if rc == 5:      # SB_ERR_BINNING_TOO_SMALL: stage 1 is done, the guess was too small -> size exactly and finish
            _lib.check(lib.sb_binning_workspace_bytes_ex(R.value, W, H, sets, ctypes.byref(n)),
                       "sb_binning_workspace_bytes_ex")
            state.binning = _ws(n.value, device)
            _lib.check(lib.sb_forward_render_ex(
                ctypes.byref(pack.c), P, R.value, _ptr(colors_precomp), _ptr(colors2), state.geom.data_ptr(),
                state.geom.numel(), state.binning.data_ptr(), state.binning.numel(), state.image.data_ptr(),
                state.image.numel(), color.data_ptr(), _ptr(color2), depth.data_ptr(), st), "sb_forward_render_ex")
        else:
            _lib.check(rc, "sb_forward")
        if state.binning is None:
            state.binning = _ws(16, device)
        state.num_rendered = R.value
        _LAST_SYNC_R[0] = R.value
        _R_HINT[hint_key] = R.value
        if len(_R_HINT) > 256:
            _R_HINT.clear()
    saved = (means3D, colors_precomp, scales, rotations, cov3Ds_precomp)
    if sets == 2:
        return color, color2, radii, depth, state, saved + (colors2,)
    return color, radii, depth, state, saved


def _backward_impl(state, saved, radii, grad_out_color, grad_out_color2=None):
    lib = _lib.load()
    means3D, colors_precomp, scales, rotations, cov3Ds_precomp = saved[:5]
    colors2 = saved[5] if len(saved) > 5 else None
    sets = 2 if colors2 is not None else 1
    device = means3D.device
    P = means3D.shape[0]
    pack = state.settings
    grad_out_color = _f32c(grad_out_color, "grad_out_color")
    with _on_device(device):
        e = lambda *shape: torch.empty(shape, dtype=torch.float32, device=device)
        g_means3D, g_means2D, g_colors, g_opac = e(P, 3), e(P, 3), e(P, 3), e(P, 1)
        have_sr = scales is not None and scales.numel() != 0
        g_scales = e(P, 3) if have_sr else torch.zeros((P, 3), dtype=torch.float32, device=device)
        g_rot = e(P, 4) if have_sr else torch.zeros((P, 4), dtype=torch.float32, device=device)
        g_cov3D = None if have_sr else e(P, 6)     # only the cov3D_precomp branch has a consumer for it
        g_colors2 = e(P, 3) if sets == 2 else None
        grad_out_color2 = _f32c(grad_out_color2, "grad_out_color2") if sets == 2 else None
        W_, H_ = int(pack.c.image_width), int(pack.c.image_height)
        bwd_ws = _ws(_sizes(lib, P, W_, H_, sets)[2], device)
        _lib.check(lib.sb_backward_ex(
            ctypes.byref(pack.c), P, state.num_rendered, _ptr(means3D), _ptr(colors_precomp), _ptr(scales),
            _ptr(rotations), _ptr(cov3Ds_precomp), _ptr(radii),
            state.geom.data_ptr(), state.geom.numel(), state.binning.data_ptr(), state.binning.numel(),
            state.image.data_ptr(), state.image.numel(), bwd_ws.data_ptr(), bwd_ws.numel(),
            _ptr(grad_out_color), _ptr(grad_out_color2), _ptr(g_means3D), _ptr(g_means2D), _ptr(g_colors),
            _ptr(g_colors2), _ptr(g_opac), _ptr(g_scales) if have_sr else None, _ptr(g_rot) if have_sr else None,
            _ptr(g_cov3D), _stream(device)), "sb_backward_ex")
    if sets == 2:
        return g_means3D, g_means2D, g_colors, g_opac, g_scales, g_rot, g_cov3D, g_colors2
    return g_means3D, g_means2D, g_colors, g_opac, g_scales, g_rot, g_cov3D


class _RasterizeGaussians(torch.autograd.Function):
    """autograd node with the reference's argument order and gradient tuple (__init__.py:40-132)."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings, capacity=None):
        if P_is_zero(means3D):
            H, W = int(raster_settings.image_height), int(raster_settings.image_width)
            dev = means3D.device
            ctx.empty = True
            ctx.dev = dev
            ctx.shapes = (means3D.shape, means2D.shape, colors_precomp.shape, opacities.shape, scales.shape,
                          rotations.shape)
            # rasterize_points.cu:67-75,81: P == 0 returns the zero-filled images without launching
            return (torch.zeros((3, H, W), device=dev), torch.zeros((0,), dtype=torch.int32, device=dev),
                    torch.zeros((1, H, W), device=dev))
        use_sh = sh is not None and sh.numel() != 0
        ctx.sh = None
        if use_sh:     # SH colours -> rgb, then the colours-precomp path (forward.cu:241-247)
            sh_c = _f32c(sh, "shs")
            m_c = _f32c(means3D, "means3D")
            P, M = sh_c.shape[0], sh_c.shape[1]
            rgb = torch.empty((P, 3), dtype=torch.float32, device=m_c.device)
            clamped = torch.empty((P,), dtype=torch.uint8, device=m_c.device)
            campos = _f32c(raster_settings.campos.to(m_c.device), "campos")
            with torch.cuda.device(m_c.device):
                _lib.check(_lib.load().sb_sh_forward(P, int(raster_settings.sh_degree), M, m_c.data_ptr(),
                                                     campos.data_ptr(), sh_c.data_ptr(), rgb.data_ptr(),
                                                     clamped.data_ptr(), _stream(m_c.device)), "sb_sh_forward")
            colors_precomp = rgb
            ctx.sh = (sh_c, clamped, campos, M, int(raster_settings.sh_degree), sh.shape)
        color, radii, depth, state, saved = _forward_impl(
            means3D, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings, capacity=capacity)
        ctx.empty = False
        ctx.state = state
        ctx.opac_shape = opacities.shape
        ctx.save_for_backward(radii, *[t if t is not None else torch.empty(0) for t in saved])
        ctx.mark_non_differentiable(radii, depth)
        return color, radii, depth

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii, _grad_depth):
        # depth carries no gradient (the reference drops it, __init__.py:88)
        if ctx.empty:
            z = [torch.zeros(s, device=ctx.dev) for s in ctx.shapes]
            return z[0], z[1], None, z[2], z[3], z[4], z[5], None, None, None
        radii, *saved = ctx.saved_tensors
        g_means3D, g_means2D, g_colors, g_opac, g_scales, g_rot, g_cov3D = _backward_impl(
            ctx.state, saved, radii, grad_out_color)
        needs_cov = saved[4].numel() != 0
        if ctx.sh is not None:
            sh_c, clamped, campos, M, deg, sh_shape = ctx.sh
            means3D = saved[0]
            g_sh = torch.empty((means3D.shape[0], M, 3), dtype=torch.float32, device=means3D.device)
            with torch.cuda.device(means3D.device):
                _lib.check(_lib.load().sb_sh_backward(means3D.shape[0], deg, M, means3D.data_ptr(), campos.data_ptr(),
                                                      sh_c.data_ptr(), clamped.data_ptr(), g_colors.data_ptr(),
                                                      radii.data_ptr(), g_sh.data_ptr(), g_means3D.data_ptr(),
                                                      _stream(means3D.device)),
                           "sb_sh_backward")
            return (g_means3D, g_means2D, g_sh.reshape(sh_shape), None, g_opac.reshape(ctx.opac_shape),
                    g_scales if not needs_cov else None, g_rot if not needs_cov else None,
                    g_cov3D if needs_cov else None, None, None)
        return (g_means3D, g_means2D, None, g_colors, g_opac.reshape(ctx.opac_shape),
                g_scales if not needs_cov else None, g_rot if not needs_cov else None,
                g_cov3D if needs_cov else None, None, None)


class _RasterizeGaussiansFused(torch.autograd.Function):
    """Two colour sets over one geometry in one pass: returns (color, color_extra, radii,
    depth).  Equivalent to two _RasterizeGaussians calls that share every input except colors_precomp;
    the means2D gradient carries the first set's share only (SplaTAM's densification statistic)."""

    @staticmethod
    def forward(ctx, means3D, means2D, colors_precomp, colors_extra, opacities, scales, rotations, raster_settings,
                capacity=None):
        if P_is_zero(means3D):
            raise _lib.SplatamB200Error("fused render needs at least one Gaussian")
        empty = torch.empty(0)
        color, color2, radii, depth, state, saved = _forward_impl(
            means3D, colors_precomp, opacities, scales, rotations, empty, raster_settings, colors2=colors_extra,
            capacity=capacity)
        ctx.state = state
        ctx.opac_shape = opacities.shape
        ctx.save_for_backward(radii, *[t if t is not None else torch.empty(0) for t in saved])
        ctx.mark_non_differentiable(radii, depth)
        return color, color2, radii, depth

    @staticmethod
    def backward(ctx, grad_color, grad_color2, _grad_radii, _grad_depth):
        radii, *saved = ctx.saved_tensors
        if grad_color is None:
            grad_color = torch.zeros_like(grad_color2)
        if grad_color2 is None:
            grad_color2 = torch.zeros_like(grad_color)
        g_means3D, g_means2D, g_colors, g_opac, g_scales, g_rot, _g_cov, g_colors2 = _backward_impl(
            ctx.state, tuple(saved), radii, grad_color, grad_color2)
        return (g_means3D, g_means2D, g_colors, g_colors2, g_opac.reshape(ctx.opac_shape), g_scales, g_rot, None, None)


def P_is_zero(means3D):
    return means3D.shape[0] == 0


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings, capacity=None):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings, capacity)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings, max_rendered=None):
        """`max_rendered` (extension): fixed capacity, in tile instances, of the binning workspace.  When given,
        the forward never synchronises with the host (num_rendered stays on the device) and forward + backward
        can be captured into a CUDA graph; `last_counts()` reports the true count and whether it overflowed."""
        super().__init__()
        self.raster_settings = raster_settings
        self.max_rendered = max_rendered

    @staticmethod
    def last_state(device=None):
        """Workspace state of the most recent sync-free forward on `device` (default: the current device)."""
        idx = torch.cuda.current_device() if device is None else torch.device(device).index
        return _LAST_ASYNC_STATE.get(idx)

    @staticmethod
    def last_counts(device=None):
        """(num_rendered, overflowed) of the most recent sync-free forward on `device` (default: the current
        device); synchronises.  On overflow the images of that forward are NaN."""
        st = GaussianRasterizer.last_state(device)
        return None if st is None else st.counts()

    def markVisible(self, positions):
        """Boolean mask ``view_z > 0.2`` per point (reference __init__.py:152-161)."""
        lib = _lib.load()
        with torch.no_grad():
            rs = self.raster_settings
            pos = _f32c(positions, "positions")
            P = pos.shape[0]
            present = torch.zeros((P,), dtype=torch.bool, device=pos.device)
            if P:
                view = _f32c(rs.viewmatrix.to(pos.device), "viewmatrix")
                proj = _f32c(rs.projmatrix.to(pos.device), "projmatrix")
                with torch.cuda.device(pos.device):
                    _lib.check(lib.sb_mark_visible(P, pos.data_ptr(), view.data_ptr(), proj.data_ptr(),
                                                   present.data_ptr(), _stream(pos.device)), "sb_mark_visible")
        return present

    def forward_fused(self, means3D, means2D, opacities, colors_precomp, colors_extra, scales, rotations):
        """One pass, two colour sets (e.g. SplaTAM's RGB and [depth, 1, depth^2]); see _RasterizeGaussiansFused."""
        return _RasterizeGaussiansFused.apply(means3D, means2D, colors_precomp, colors_extra, opacities, scales,
                                              rotations, self.raster_settings, self.max_rendered)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        raster_settings = self.raster_settings
        # the two argument checks of the reference, same exception type and text (__init__.py:167-171)
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, raster_settings, self.max_rendered)
