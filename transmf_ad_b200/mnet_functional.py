"""Autograd glue of the MiSePyNet / Mnet baseline (reference models/MiSePyNet.py) over csrc/mnet_ops.cu: fp32 NCDHW tensors.

``conv_bn_relu`` is one autograd node per Conv3d + BatchNorm3d + ReLU triple (the only pattern the reference uses):
conv -> per-block partial statistics -> ``tmf_bn_finalize`` (shared with the sNet path) -> BN + ReLU apply; the backward is
reduce -> ``tmf_bn_bwd_finalize`` -> apply -> weight gradient (+ input gradient).  Saved: the conv input, the pre-BN conv
output and 4C coefficients.  No CPU / library fallback.
"""
from __future__ import annotations

import torch

from . import _lib as L
from .functional import _f32c, grad_out


class ConvBNReLUFunction(torch.autograd.Function):
    """cfg = (kind, kernel, stride, training, eps, momentum); kind 'line': Conv3d(Cin, 8, (1,1,k)) along the last axis of
    x (N,Cin,A,B,L); kind '2d': Conv3d(Cin, Cout, (kh,kw,1), stride) on x (N,Cin,X,Y,1)."""

    @staticmethod
    def forward(ctx, x, w, b, gamma, beta, rmean, rvar, nbt, cfg):
        kind, kern, stride, training, eps, momentum = cfg
        x = _f32c(x)
        N, Cin = x.shape[0], x.shape[1]
        Cout = w.shape[0]
        dev = x.device
        wc, bc = _f32c(w), _f32c(b)
        if kind == "line":
            A, Bd, Ln = x.shape[2], x.shape[3], x.shape[4]
            k = kern
            z = torch.empty((N, Cout, A, Bd, Ln - k + 1), dtype=torch.float32, device=dev)
            L.call("tmf_line_conv_fwd", L.ptr(x), L.ptr(wc), L.ptr(bc), L.ptr(z), N, Cin, A * Bd, Ln, k)
        else:
            X, Y = x.shape[2], x.shape[3]
            kh, kw = kern
            Xo, Yo = (X - kh) // stride + 1, (Y - kw) // stride + 1
            z = torch.empty((N, Cout, Xo, Yo, 1), dtype=torch.float32, device=dev)
            L.call("tmf_conv2d_fwd", L.ptr(x), L.ptr(wc), L.ptr(bc), L.ptr(z), N, Cin, X, Y, Cout, kh, kw, stride)
        S = z.numel() // (N * Cout)
        coef = torch.empty(4 * Cout, dtype=torch.float32, device=dev)
        rows = None
        if training:
            rows = L.stat_buffers(1, Cout, dev)[0]
            L.call("tmf_nchw_bn_stats", L.ptr(z), L.ptr(rows), N, Cout, S)
        L.call("tmf_bn_finalize", 1, L.ptrs([rows]) if training else L.ptrs(None), L.ptrs([_f32c(gamma)]), L.ptrs([_f32c(beta)]),
               L.ptrs([rmean]), L.ptrs([rvar]), L.ptrs([nbt]), L.ptrs([coef]), Cout, N * S, float(momentum), float(eps), int(training))
        out = torch.empty_like(z)
        L.call("tmf_nchw_bn_relu_fwd", L.ptr(z), L.ptr(coef), L.ptr(out), N, Cout, S)
        ctx.save_for_backward(x, wc, z, coef)
        ctx.cfg = cfg
        ctx.refs = (w, b, gamma, beta)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, wc, z, coef = ctx.saved_tensors
        kind, kern, stride, training, eps, momentum = ctx.cfg
        w, b, gamma, beta = ctx.refs
        N, Cin, Cout = x.shape[0], x.shape[1], wc.shape[0]
        dev = x.device
        S = z.numel() // (N * Cout)
        dout = _f32c(dout)
        rows = L.stat_buffers(1, Cout, dev)[0]
        L.call("tmf_nchw_bn_relu_bwd_reduce", L.ptr(dout), L.ptr(z), L.ptr(coef), L.ptr(rows), N, Cout, S)
        dgamma, dbeta, dbias = grad_out(gamma), grad_out(beta), grad_out(b)
        bcoef = torch.empty(2 * Cout, dtype=torch.float32, device=dev)
        L.call("tmf_bn_bwd_finalize", 1, L.ptrs([rows]), L.ptrs([coef]), L.ptrs([dgamma]), L.ptrs([dbeta]), L.ptrs([dbias]),
               L.ptrs([bcoef]), Cout, N * S, int(training))
        dz = torch.empty_like(z)
        L.call("tmf_nchw_bn_relu_bwd_apply", L.ptr(dout), L.ptr(z), L.ptr(coef), L.ptr(bcoef), L.ptr(dz), N, Cout, S)
        dw = grad_out(w)
        dx = None
        if kind == "line":
            A, Bd, Ln = x.shape[2], x.shape[3], x.shape[4]
            k = kern
            nws = int(L.load().tmf_line_conv_wgrad_workspace_bytes(Cin, k))
            ws = torch.empty(nws, dtype=torch.uint8, device=dev)
            L.call("tmf_line_conv_wgrad", L.ptr(dz), L.ptr(x), L.ptr(dw), L.ptr(None), N, Cin, A * Bd, Ln, k, L.ptr(ws), nws)
            if ctx.needs_input_grad[0]:
                dx = torch.empty_like(x)
                L.call("tmf_line_conv_dgrad", L.ptr(dz), L.ptr(wc), L.ptr(dx), N, Cin, A * Bd, Ln, k)
        else:
            X, Y = x.shape[2], x.shape[3]
            kh, kw = kern
            L.call("tmf_conv2d_wgrad", L.ptr(dz), L.ptr(x), L.ptr(dw), L.ptr(None), N, Cin, X, Y, Cout, kh, kw, stride)
            if ctx.needs_input_grad[0]:
                dx = torch.empty_like(x)
                L.call("tmf_conv2d_dgrad", L.ptr(dz), L.ptr(wc), L.ptr(dx), N, Cin, X, Y, Cout, kh, kw, stride)
        return dx, dw, dbias, dgamma, dbeta, None, None, None, None


def conv_bn_relu(x, conv, bn, training):
    """``conv``: nn.Conv3d with kernel (1,1,k) [stride 1] or (kh,kw,1) [stride (s,s,s)], no padding; ``bn``: nn.BatchNorm3d."""
    kd, kh, kw = conv.kernel_size
    if bn.momentum is None or not bn.track_running_stats or not bn.affine or conv.bias is None:
        raise NotImplementedError("Mnet on the B200 kernels needs biased convs and affine BatchNorm3d with running statistics")
    if any(p != 0 for p in conv.padding) or any(d != 1 for d in conv.dilation) or conv.groups != 1:
        raise NotImplementedError("Mnet conv kernels take unpadded, undilated, ungrouped convolutions (what the reference uses)")
    if kw == 1:                                            # (kh,kw,1) spatial convolution, incl. the 1x1x1 one
        if x.shape[-1] != 1 or len(set(conv.stride)) != 1:
            raise NotImplementedError("spatial convolution expects a unit last axis and an isotropic stride")
        cfg = ("2d", (kd, kh), conv.stride[0], bool(training), bn.eps, bn.momentum)
    elif kd == 1 and kh == 1:
        if conv.stride != (1, 1, 1) or conv.out_channels != 8 or conv.in_channels > 8:
            raise NotImplementedError("slice convolution must be Conv3d(<=8, 8, (1,1,k)), stride 1")
        cfg = ("line", kw, 1, bool(training), bn.eps, bn.momentum)
    else:
        raise NotImplementedError(f"unsupported Mnet kernel size {conv.kernel_size}")
    return ConvBNReLUFunction.apply(x, conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                    bn.num_batches_tracked, cfg)


class MaxPool2dFunction(torch.autograd.Function):
    """nn.MaxPool3d((ph,pw,1)) (stride = kernel, floor mode) on (N,C,X,Y,1)."""

    @staticmethod
    def forward(ctx, x, ph, pw):
        x = _f32c(x)
        N, C, X, Y = x.shape[:4]
        y = torch.empty((N, C, X // ph, Y // pw, 1), dtype=torch.float32, device=x.device)
        idx = torch.empty(y.shape, dtype=torch.uint8, device=x.device)
        L.call("tmf_maxpool2d_fwd", L.ptr(x), L.ptr(y), L.ptr(idx), N * C, X, Y, ph, pw)
        ctx.save_for_backward(idx)
        ctx.cfg = (N, C, X, Y, ph, pw)
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        N, C, X, Y, ph, pw = ctx.cfg
        dx = torch.empty((N, C, X, Y, 1), dtype=torch.float32, device=dy.device)
        L.call("tmf_maxpool2d_bwd", L.ptr(_f32c(dy)), L.ptr(idx), L.ptr(dx), N * C, X, Y, ph, pw)
        return dx, None, None


def maxpool_xy(x, pool):
    ph, pw, pz = pool.kernel_size
    if pz != 1 or x.shape[-1] != 1 or pool.padding not in (0, (0, 0, 0)) or pool.stride not in (pool.kernel_size, None):
        raise NotImplementedError("Mnet pooling kernel takes MaxPool3d((ph,pw,1)) with stride = kernel and no padding")
    return MaxPool2dFunction.apply(x, ph, pw)
