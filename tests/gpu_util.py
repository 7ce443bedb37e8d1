"""Helpers for the -m gpu parity tests: wrap torch CUDA tensors as phs_tensor and call the C-ABI."""
import ctypes

import torch


class Caller:
    def __init__(self, lib):
        self.L = lib
        self.h = lib.load()
        self.keep = []

    def T(self, t, c_off=0, C=None):
        """phs_tensor view of an NHWC torch tensor [N,H,W,ld] (optionally a channel slice)."""
        assert t.is_cuda and t.is_contiguous() and t.dim() == 4
        dt = {torch.float32: self.L.PHS_F32, torch.bfloat16: self.L.PHS_BF16}[t.dtype]
        N, H, W, ld = t.shape
        C = ld - c_off if C is None else C
        d = self.L.phs_tensor(t.data_ptr() + c_off * t.element_size(), N, H, W, C, ld, dt)
        self.keep.append((d, t))
        return ctypes.byref(d)

    def __call__(self, name, *args):
        st = torch.cuda.current_stream().cuda_stream
        args = [a.data_ptr() if torch.is_tensor(a) else a for a in args]
        rc = getattr(self.h, name)(*args, st)
        self.L.check(rc, name)

    def rc(self, name, *args):
        st = torch.cuda.current_stream().cuda_stream
        args = [a.data_ptr() if torch.is_tensor(a) else a for a in args]
        return getattr(self.h, name)(*args, st)


def cu(t, dtype=None):
    t = t.detach().contiguous().cuda()
    return t.to(dtype) if dtype is not None else t
