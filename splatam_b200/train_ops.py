"""Host-side wrappers of the fused training ops in libsplatam_b200.so (csrc/train_ops.cu):
``FusedAdam`` over a flat parameter buffer and ``image_loss`` = 0.8*L1 + 0.2*(1-SSIM) with a fused
backward.  CUDA only; no CPU fallback."""
import ctypes

import torch

from . import _lib


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class FusedAdam:
    """torch.optim.Adam semantics (betas, eps, per-segment lr; no amsgrad / weight decay) over ONE flat
    fp32 buffer and its flat gradient; one kernel launch per step."""

    def __init__(self, flat, flat_grad, seg_sizes, seg_lrs, betas=(0.9, 0.999), eps=1e-15):
        if not flat.is_cuda:
            raise _lib.SplatamB200Error("FusedAdam needs CUDA tensors (there is no CPU fallback)")
        self.flat, self.flat_grad = flat, flat_grad
        self.m, self.v = torch.zeros_like(flat), torch.zeros_like(flat)
        ends, acc = [], 0
        for n in seg_sizes:
            acc += int(n)
            ends.append(acc)
        assert acc == flat.numel() and len(ends) == len(seg_lrs) <= 16
        self.n = len(ends)
        self.seg_end = (ctypes.c_uint32 * self.n)(*ends)
        self.seg_lr = (ctypes.c_float * self.n)(*[float(x) for x in seg_lrs])
        self.betas, self.eps, self.t = betas, eps, 0

    def step(self):
        self.t += 1
        lib = _lib.load()
        with torch.cuda.device(self.flat.device):
            _lib.check(lib.sb_adam_step(self.flat.data_ptr(), self.flat_grad.data_ptr(), self.m.data_ptr(),
                                        self.v.data_ptr(), self.flat.numel(), self.seg_end, self.seg_lr, self.n, self.t,
                                        self.betas[0], self.betas[1], self.eps, _stream(self.flat.device)), "sb_adam_step")


class _ImageLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, w_l1, w_ssim):
        lib = _lib.load()
        x, y = x.contiguous().float(), y.contiguous().float()
        C, H, W = x.shape
        dev = x.device
        work = torch.empty(lib.sb_image_loss_workspace_floats(C, H, W), dtype=torch.float32, device=dev)
        sums = torch.empty(2, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.sb_image_loss_forward(x.data_ptr(), y.data_ptr(), C, H, W, work.data_ptr(), sums.data_ptr(),
                                                 _stream(dev)), "sb_image_loss_forward")
        n = float(C * H * W)
        ctx.save_for_backward(x, y, work)
        ctx.w = (float(w_l1), float(w_ssim))
        return (w_l1 * sums[1] / n + w_ssim * (1.0 - sums[0] / n)).float()

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        x, y, work = ctx.saved_tensors
        C, H, W = x.shape
        gx = torch.empty_like(x)
        go = grad_out.contiguous().float().reshape(1)
        with torch.cuda.device(x.device):
            _lib.check(lib.sb_image_loss_backward(x.data_ptr(), y.data_ptr(), C, H, W, work.data_ptr(), go.data_ptr(),
                                                  ctx.w[1], ctx.w[0], gx.data_ptr(), _stream(x.device)),
                       "sb_image_loss_backward")
        return gx, None, None, None


def image_loss(im, gt, w_l1=0.8, w_ssim=0.2):
    """0.8 * mean|im-gt| + 0.2 * (1 - SSIM(im, gt)) -- SplaTAM's mapping RGB loss (R/scripts/splatam.py:290)."""
    if not im.is_cuda:
        raise _lib.SplatamB200Error("image_loss needs CUDA tensors (there is no CPU fallback)")
    return _ImageLoss.apply(im, gt, w_l1, w_ssim)
