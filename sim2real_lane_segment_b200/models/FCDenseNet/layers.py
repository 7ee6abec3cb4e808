"""Parameter containers that reproduce the reference's module tree (rightLaneNetwork/models/FCDenseNet/layers.py)
so that ``state_dict()`` / ``load_state_dict()`` / checkpoints are interchangeable with the reference.

These modules hold tensors only; the arithmetic of DenseLayer (layers.py:5-12), DenseBlock (:18-40),
TransitionDown (:43-52), TransitionUp (:58-69), Bottleneck (:72-76) and center_crop (:82-86) is executed by the
fused CUDA schedule in csrc/ (see DESIGN.md), never layer by layer in Python — calling ``forward`` on a piece raises.
"""
import math

import torch
import torch.nn as nn


class _NoStandaloneForward(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError(f"{type(self).__name__} is a parameter container of the fused CUDA network; "
                           "call the FCDenseNetFeatureExtractor / FCDenseNetClassifier instead")


class _Conv(_NoStandaloneForward):
    """weight/bias with nn.Conv2d's (or nn.ConvTranspose2d's) shapes and default initialisation"""

    def __init__(self, cin, cout, k, transposed=False):
        super().__init__()
        shape = (cin, cout, k, k) if transposed else (cout, cin, k, k)
        self.weight = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.empty(cout))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        fan_in = shape[1] * k * k
        bound = 1 / math.sqrt(fan_in)
        nn.init.uniform_(self.bias, -bound, bound)


class _Norm(_NoStandaloneForward):
    """nn.BatchNorm2d's tensors (eps 1e-5, momentum 0.1, affine, tracked running statistics)"""

    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class DenseLayer(_NoStandaloneForward):
    def __init__(self, in_channels, growth_rate):
        super().__init__()
        self.norm = _Norm(in_channels)
        self.conv = _Conv(in_channels, growth_rate, 3)


class DenseBlock(_NoStandaloneForward):
    def __init__(self, in_channels, growth_rate, n_layers, upsample=False):
        super().__init__()
        self.upsample = upsample
        self.layers = nn.ModuleList([DenseLayer(in_channels + i * growth_rate, growth_rate) for i in range(n_layers)])


class TransitionDown(_NoStandaloneForward):
    def __init__(self, in_channels):
        super().__init__()
        self.norm = _Norm(in_channels)
        self.conv = _Conv(in_channels, in_channels, 1)


class TransitionUp(_NoStandaloneForward):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.convTrans = _Conv(in_channels, out_channels, 3, transposed=True)


class Bottleneck(_NoStandaloneForward):
    def __init__(self, in_channels, growth_rate, n_layers):
        super().__init__()
        self.bottleneck = DenseBlock(in_channels, growth_rate, n_layers, upsample=True)
