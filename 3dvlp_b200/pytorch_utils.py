"""SharedMLP and friends, checkpoint-compatible with the reference.

The reference (/root/reference/lib/pointnet2/pytorch_utils.py) builds its shared MLPs as nested
``nn.Sequential`` containers; the names of those containers ARE the checkpoint format, e.g.

    sa1.mlp_module.layer0.conv.weight
    sa1.mlp_module.layer0.bn.bn.{weight,bias,running_mean,running_var,num_batches_tracked}

(SharedMLP :11-36 -> ``layer{i}`` = Conv2d :157-188 -> ``conv`` [no bias under bn], ``bn`` [a wrapper
Sequential holding ``bn``, :39-58], ``activation``).  This module reproduces that tree, the
constructor signatures and the initialisation (:45-46,96: kaiming-normal conv weights, BN weight 1 /
bias 0) so ``load_state_dict`` works both ways.  It only declares parameters; the arithmetic of
the hot path is done by the fused sm_100a kernels (see pointnet2_modules.py).
"""
import torch.nn as nn

_CONV = {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}
_NORM = {1: nn.BatchNorm1d, 2: nn.BatchNorm2d, 3: nn.BatchNorm3d}


class _NormWrap(nn.Sequential):
    """A Sequential with a single child called ``bn`` (weight 1, bias 0)."""

    ndim = 2

    def __init__(self, in_size, name=""):
        super().__init__()
        norm = _NORM[self.ndim](in_size)
        nn.init.constant_(norm.weight, 1.0)
        nn.init.constant_(norm.bias, 0)
        self.add_module(name + "bn", norm)


class BatchNorm1d(_NormWrap):
    ndim = 1

    def __init__(self, in_size, *, name=""):
        super().__init__(in_size, name)


class BatchNorm2d(_NormWrap):
    ndim = 2


class BatchNorm3d(_NormWrap):
    ndim = 3


_WRAP = {1: BatchNorm1d, 2: BatchNorm2d, 3: BatchNorm3d}


class _ConvBlock(nn.Sequential):
    """conv (+ bn) (+ activation), or the pre-activation order when ``preact``."""

    ndim = 2

    def __init__(self, in_size, out_size, *, kernel_size=None, stride=None, padding=None,
                 activation=nn.ReLU(inplace=True), bn=False, init=nn.init.kaiming_normal_, bias=True,
                 preact=False, name=""):
        super().__init__()
        d = self.ndim
        ks = kernel_size if kernel_size is not None else ((1,) * d if d > 1 else 1)
        st = stride if stride is not None else ((1,) * d if d > 1 else 1)
        pd = padding if padding is not None else ((0,) * d if d > 1 else 0)
        use_bias = bias and not bn
        conv = _CONV[d](in_size, out_size, kernel_size=ks, stride=st, padding=pd, bias=use_bias)
        init(conv.weight)
        if use_bias:
            nn.init.constant_(conv.bias, 0)
        tail = []
        if bn:
            tail.append((name + "bn", _WRAP[d](in_size if preact else out_size)))
        if activation is not None:
            tail.append((name + "activation", activation))
        order = tail + [(name + "conv", conv)] if preact else [(name + "conv", conv)] + tail
        for key, mod in order:
            self.add_module(key, mod)


class Conv1d(_ConvBlock):
    ndim = 1


class Conv2d(_ConvBlock):
    ndim = 2


class Conv3d(_ConvBlock):
    ndim = 3


class SharedMLP(nn.Sequential):
    """Stack of 1x1 Conv2d blocks over (B, C, npoint, nsample) tensors: args = channel list."""

    def __init__(self, args, *, bn=False, activation=nn.ReLU(inplace=True), preact=False, first=False, name=""):
        super().__init__()
        for i, (cin, cout) in enumerate(zip(args[:-1], args[1:])):
            plain = (not first) or (not preact) or (i != 0)
            self.add_module("{}layer{}".format(name, i),
                            Conv2d(cin, cout, bn=plain and bn, activation=activation if plain else None,
                                   preact=preact))


class FC(nn.Sequential):
    def __init__(self, in_size, out_size, *, activation=nn.ReLU(inplace=True), bn=False, init=None, preact=False,
                 name=""):
        super().__init__()
        fc = nn.Linear(in_size, out_size, bias=not bn)
        if init is not None:
            init(fc.weight)
        if not bn:
            nn.init.constant_(fc.bias, 0)
        tail = []
        if bn:
            tail.append((name + "bn", BatchNorm1d(in_size if preact else out_size)))
        if activation is not None:
            tail.append((name + "activation", activation))
        order = tail + [(name + "fc", fc)] if preact else [(name + "fc", fc)] + tail
        for key, mod in order:
            self.add_module(key, mod)


def set_bn_momentum_default(bn_momentum):
    def fn(m):
        if isinstance(m, tuple(_NORM.values())):
            m.momentum = bn_momentum
    return fn


class BNMomentumScheduler(object):
    """Re-schedules BatchNorm momentum per epoch (pytorch_utils.py:271-296; the reference solver
    imports it, lib/joint/solver_3dvlp.py:21)."""

    def __init__(self, model, bn_lambda, last_epoch=-1, setter=set_bn_momentum_default):
        if not isinstance(model, nn.Module):
            raise RuntimeError("Class '{}' is not a PyTorch nn Module".format(type(model).__name__))
        self.model, self.setter, self.lmbd = model, setter, bn_lambda
        self.step(last_epoch + 1)
        self.last_epoch = last_epoch

    def step(self, epoch=None):
        self.last_epoch = self.last_epoch + 1 if epoch is None else epoch
        self.model.apply(self.setter(self.lmbd(self.last_epoch)))
