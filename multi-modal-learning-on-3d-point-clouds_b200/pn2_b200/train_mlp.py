"""Training-mode shared MLP: n x (1x1 conv -> BatchNorm (batch statistics) -> ReLU) as ONE autograd Function over
channel-last row matrices, on the kernels of csrc/train_mlp.cu.

The reference runs this chain as torch.nn.Conv2d / Conv1d + BatchNorm + F.relu per layer (model/pointnet_util.py:105-107,
162-165, 218-220): cuDNN fprop / dgrad / wgrad plus separate BatchNorm and ReLU passes in both directions.  Here every
layer is one forward launch (the previous layer's normalisation and ReLU are applied while the operand is loaded, the
batch statistics are reduced in the same pass) and three backward launches (BatchNorm reductions, input gradient, weight
gradient); only the pre-BatchNorm activations are kept for the backward.  Semantics are torch's: biased variance for the
normalisation, unbiased for the running estimate, eps / momentum of the module, running statistics updated in place.
"""
import torch
from torch.autograd import Function

from . import _lib
from ._lib import ptr


def _stream(t):
    return _lib.stream_ptr(t.device)


class _FusedMlpTrain(Function):
    @staticmethod
    def forward(ctx, x0, meta, *tensors):
        """x0 (R, C0) fp32 rows; meta = [(eps, momentum or None, track), ...] per layer; tensors = (W (cout, cin), b, gamma,
        beta, running_mean or None, running_var or None) per layer -> a_L (R, C_L).  The running statistics are updated in
        place by the per-channel kernel (momentum, unbiased variance), like torch.nn.BatchNorm in training mode."""
        L = len(tensors) // 6
        R = x0.shape[0]
        dev = x0.device
        x0 = _lib.check_f32(x0, "x0")
        zs, scales, shifts, mrs = [], [], [], []
        in_scale = in_shift = None
        x = x0
        st = _stream(x0)
        with torch.cuda.device(dev):
            for l in range(L):
                W, b, gamma, beta, rmean, rvar = tensors[6 * l:6 * l + 6]
                eps, momentum = meta[l]
                cout, cin = W.shape
                Wc, bc = W.detach().contiguous(), b.detach().contiguous()
                gc, bec = gamma.detach().contiguous(), beta.detach().contiguous()
                z = torch.empty((R, cout), dtype=torch.float32, device=dev)
                stats = torch.zeros((2, cout), dtype=torch.float64, device=dev)
                _lib.call("pn2_train_linear_fwd", R, cin, cout, ptr(x), ptr(in_scale), ptr(in_shift), ptr(Wc), ptr(bc), ptr(z), ptr(stats), st)
                scale = torch.empty(cout, dtype=torch.float32, device=dev)
                shift = torch.empty(cout, dtype=torch.float32, device=dev)
                mr = torch.empty((2, cout), dtype=torch.float64, device=dev)
                _lib.call("pn2_train_bn_finalize", R, cout, ptr(stats), ptr(gc), ptr(bec), float(eps), float(momentum), ptr(scale), ptr(shift),
                          ptr(mr), ptr(rmean), ptr(rvar), st)
                zs.append(z); scales.append(scale); shifts.append(shift); mrs.append(mr)
                x, in_scale, in_shift = z, scale, shift
            a = torch.empty_like(zs[-1])
            _lib.call("pn2_train_bn_relu", R, a.shape[1], ptr(zs[-1]), ptr(scales[-1]), ptr(shifts[-1]), ptr(a), st)
        ctx.L = L
        params = []
        for l in range(L):
            params += list(tensors[6 * l:6 * l + 4])
        ctx.save_for_backward(x0, *zs, *scales, *shifts, *mrs, *params)
        return a

    @staticmethod
    def backward(ctx, g_a):
        L = ctx.L
        saved = ctx.saved_tensors
        x0 = saved[0]
        zs, scales, shifts, mrs = saved[1:1 + L], saved[1 + L:1 + 2 * L], saved[1 + 2 * L:1 + 3 * L], saved[1 + 3 * L:1 + 4 * L]
        params = saved[1 + 4 * L:]
        R = x0.shape[0]
        dev = x0.device
        g = g_a.contiguous()
        grads = [None] * (6 * L)
        st = _stream(x0)
        with torch.cuda.device(dev):
            for l in range(L - 1, -1, -1):
                W, b, gamma, beta = params[4 * l:4 * l + 4]
                cout, cin = W.shape
                z, scale, shift, mr = zs[l], scales[l], shifts[l], mrs[l]
                sums = torch.zeros((2, cout), dtype=torch.float64, device=dev)
                _lib.call("pn2_train_bn_bwd_reduce", R, cout, ptr(g), ptr(z), ptr(scale), ptr(shift), ptr(sums), st)
                coef = torch.empty((3, cout), dtype=torch.float32, device=dev)
                dgamma = torch.empty(cout, dtype=torch.float32, device=dev)
                dbeta = torch.empty(cout, dtype=torch.float32, device=dev)
                gc = gamma.detach().contiguous()
                _lib.call("pn2_train_bn_bwd_coeffs", R, cout, ptr(sums), ptr(mr), ptr(gc), ptr(coef), ptr(dgamma), ptr(dbeta), st)
                need_in = l > 0 or ctx.needs_input_grad[0]
                g_in = torch.empty((R, cin), dtype=torch.float32, device=dev) if need_in else None
                dW = torch.zeros((cout, cin), dtype=torch.float32, device=dev)
                x_in = x0 if l == 0 else zs[l - 1]
                in_scale = None if l == 0 else scales[l - 1]
                in_shift = None if l == 0 else shifts[l - 1]
                Wt = W.detach().t().contiguous()
                _lib.call("pn2_train_linear_bwd", R, cin, cout, ptr(x_in), ptr(in_scale), ptr(in_shift), ptr(Wt), ptr(g), ptr(z),
                          ptr(scale), ptr(shift), ptr(coef[0]), ptr(coef[1]), ptr(coef[2]), ptr(g_in), ptr(dW), st)
                grads[6 * l] = dW.view_as(W)
                grads[6 * l + 1] = torch.zeros_like(b)   # a bias in front of a batch-statistics BatchNorm has zero gradient
                grads[6 * l + 2] = dgamma
                grads[6 * l + 3] = dbeta
                g = g_in
        return (g, None, *grads)


def fused_mlp_train(x0, convs, bns):
    """x0 (R, C0) rows -> (R, C_L) = the conv -> BatchNorm(train) -> ReLU chain; updates the BatchNorm running statistics
    (momentum, unbiased variance, num_batches_tracked) exactly like torch.nn.BatchNorm in training mode."""
    tensors, meta = [], []
    for conv, bn in zip(convs, bns):
        track = bn.track_running_stats and bn.running_mean is not None
        if track:
            bn.num_batches_tracked += 1
        momentum = bn.momentum if bn.momentum is not None else (1.0 / float(bn.num_batches_tracked) if track else 0.0)
        tensors += [conv.weight.reshape(conv.out_channels, conv.in_channels), conv.bias, bn.weight, bn.bias,
                    bn.running_mean if track else None, bn.running_var if track else None]
        meta.append((float(bn.eps), float(momentum)))
    return _FusedMlpTrain.apply(x0, meta, *tensors)


def fusable_training(x0, convs, bns):
    """The fused training chain serves CUDA fp32 stacks of 1x1 convolutions with bias, each followed by an affine BatchNorm."""
    if not (x0.is_cuda and x0.dtype == torch.float32):
        return False
    for conv, bn in zip(convs, bns):
        if conv.bias is None or bn is None or not bn.affine or conv.weight.dtype != torch.float32 or not conv.weight.is_cuda:
            return False
        if tuple(conv.kernel_size) != (1,) * len(conv.kernel_size):
            return False
    return True
