"""TEST INFRASTRUCTURE ONLY — functional fp32 restatement of the reference FC-DenseNet (Tiramisu).

Follows /root/reference/rightLaneNetwork/models/FCDenseNet/:
  layers.py:5-12    DenseLayer      BN -> ReLU -> Conv3x3(Cin->growth, pad 1, bias) -> Dropout2d(0.2)
  layers.py:18-40   DenseBlock      x = cat([x, out]); upsample=True returns only the new features
  layers.py:43-52   TransitionDown  BN -> ReLU -> Conv1x1(C->C) -> Dropout2d(0.2) -> MaxPool2d(2)
  layers.py:58-69   TransitionUp    ConvTranspose2d(k3,s2,p0) -> center_crop -> cat([up, skip])
  layers.py:82-86   center_crop
  tiramisu.py:21-109 FCDenseNetFeatureExtractor (forward :89-106, F.normalize at :105)
  tiramisu.py:112-125 FCDenseNetClassifier (1x1 conv, /T with T=0.05, softmax over channels)
The network is expressed as pure functions of a state_dict with the reference's key names, so the
same tensors can be fed to the CUDA path and to this oracle.
"""
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class NetConfig:
    in_channels: int = 3
    down_blocks: Tuple[int, ...] = (5, 5, 5, 5, 5)
    up_blocks: Tuple[int, ...] = (5, 5, 5, 5, 5)
    bottleneck_layers: int = 5
    growth_rate: int = 16
    first_conv: int = 48

    @property
    def feature_channels(self) -> int:
        # tiramisu.py:76-87: last up block keeps its input (prev new features + level-0 skip) and adds its own
        c = self.first_conv
        skips = []
        for n in self.down_blocks:
            c += self.growth_rate * n
            skips.append(c)
        prev = self.growth_rate * self.bottleneck_layers
        for i in range(len(self.up_blocks) - 1):
            prev = self.growth_rate * self.up_blocks[i]
        return prev + skips[0] + self.growth_rate * self.up_blocks[-1]


FCDENSENET67 = NetConfig()
FCDENSENET57 = NetConfig(down_blocks=(4,) * 5, up_blocks=(4,) * 5, bottleneck_layers=4, growth_rate=12)

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
DROP_P = 0.2
TEMPERATURE = 0.05


class _Ctx:
    """Carries mode flags and collects side results of one forward pass."""

    def __init__(self, sd, train, drop_masks, update_running):
        self.sd = sd
        self.train = train
        self.drop_masks = drop_masks
        self.update_running = update_running
        self.new_running: Dict[str, torch.Tensor] = {}
        self.taps: Dict[str, torch.Tensor] = {}


def _bn_relu(ctx: _Ctx, x, prefix):
    w, b = ctx.sd[prefix + ".weight"], ctx.sd[prefix + ".bias"]
    rm, rv = ctx.sd[prefix + ".running_mean"], ctx.sd[prefix + ".running_var"]
    if ctx.train:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if ctx.update_running:
            n = x.numel() / x.shape[1]
            with torch.no_grad():
                ctx.new_running[prefix + ".running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean
                ctx.new_running[prefix + ".running_var"] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var * n / max(n - 1, 1)
    else:
        mean, var = rm, rv
    xh = (x - mean[None, :, None, None]) * torch.rsqrt(var + BN_EPS)[None, :, None, None]
    return torch.relu(xh * w[None, :, None, None] + b[None, :, None, None])


def _dropout2d(ctx: _Ctx, x, name):
    """Dropout2d given an explicit keep mask [B, C] of 0/1 (train mode); identity if no mask supplied."""
    if not ctx.train or ctx.drop_masks is None or name not in ctx.drop_masks:
        return x
    keep = ctx.drop_masks[name].to(x.dtype)
    return x * keep[:, :, None, None] / (1.0 - DROP_P)


def _dense_layer(ctx, x, prefix):
    y = _bn_relu(ctx, x, prefix + ".norm")
    y = F.conv2d(y, ctx.sd[prefix + ".conv.weight"], ctx.sd[prefix + ".conv.bias"], padding=1)
    return _dropout2d(ctx, y, prefix + ".drop")


def _dense_block(ctx, x, prefix, n_layers, only_new):
    new = []
    for i in range(n_layers):
        out = _dense_layer(ctx, x, f"{prefix}.layers.{i}")
        x = torch.cat([x, out], 1)
        new.append(out)
    ctx.taps[prefix] = x  # full concat in the reference's channel order (test introspection only)
    return torch.cat(new, 1) if only_new else x


def _transition_down(ctx, x, prefix):
    y = _bn_relu(ctx, x, prefix + ".norm")
    y = F.conv2d(y, ctx.sd[prefix + ".conv.weight"], ctx.sd[prefix + ".conv.bias"])
    y = _dropout2d(ctx, y, prefix + ".drop")
    return F.max_pool2d(y, 2)


def _transition_up(ctx, x, skip, prefix):
    y = F.conv_transpose2d(x, ctx.sd[prefix + ".convTrans.weight"], ctx.sd[prefix + ".convTrans.bias"], stride=2)
    h, w = skip.shape[2], skip.shape[3]
    x0, y0 = (y.shape[3] - w) // 2, (y.shape[2] - h) // 2
    y = y[:, :, y0:y0 + h, x0:x0 + w]
    return torch.cat([y, skip], 1)


def feature_extractor(sd, x, cfg: NetConfig = FCDENSENET67, train=False, drop_masks=None, update_running=False,
                      prefix="featureExtractor.", normalize=True, taps=None):
    """tiramisu.py:89-106.  Returns (features, dict of updated running stats)."""
    ctx = _Ctx(sd, train, drop_masks, update_running)
    if taps is not None:
        ctx.taps = taps
    p = prefix
    out = F.conv2d(x, sd[p + "firstconv.weight"], sd[p + "firstconv.bias"], padding=1)
    skips = []
    for i, n in enumerate(cfg.down_blocks):
        out = _dense_block(ctx, out, f"{p}denseBlocksDown.{i}", n, only_new=False)
        skips.append(out)
        out = _transition_down(ctx, out, f"{p}transDownBlocks.{i}")
    out = _dense_block(ctx, out, f"{p}bottleneck.bottleneck", cfg.bottleneck_layers, only_new=True)
    for i, n in enumerate(cfg.up_blocks):
        out = _transition_up(ctx, out, skips.pop(), f"{p}transUpBlocks.{i}")
        out = _dense_block(ctx, out, f"{p}denseBlocksUp.{i}", n, only_new=(i != len(cfg.up_blocks) - 1))
    if normalize:
        out = F.normalize(out)  # L2 over channels, eps 1e-12
    return out, ctx.new_running


def classifier(sd, feats, use_softmax=True, prefix="classifier.", temperature=TEMPERATURE):
    """tiramisu.py:120-125"""
    z = F.conv2d(feats, sd[prefix + "finalConv.weight"], sd[prefix + "finalConv.bias"])
    z = z / temperature
    return torch.softmax(z, dim=1) if use_softmax else z


def forward(sd, x, cfg: NetConfig = FCDENSENET67, train=False, drop_masks=None, update_running=False):
    """TrainingBase.forward (trainingModules/TrainingBase.py:54-57): probabilities [B, C, H, W]."""
    feats, new_running = feature_extractor(sd, x, cfg, train, drop_masks, update_running)
    return classifier(sd, feats), new_running


# ------------------------------------------------------------------------------------------------
# parameter bookkeeping (names and shapes exactly as the reference's state_dict)
# ------------------------------------------------------------------------------------------------
def state_dict_spec(cfg: NetConfig = FCDENSENET67, num_cls: int = 4):
    """Ordered list of (name, shape, kind) with kind in {'param', 'buffer'} in the reference's
    registration order (TrainingBase.__init__: featureExtractor then classifier)."""
    spec = []
    g = cfg.growth_rate

    def bn(prefix, c):
        spec.extend([(prefix + ".weight", (c,), "param"), (prefix + ".bias", (c,), "param"),
                     (prefix + ".running_mean", (c,), "buffer"), (prefix + ".running_var", (c,), "buffer"),
                     (prefix + ".num_batches_tracked", (), "buffer")])

    def conv(prefix, co, ci, k):
        spec.extend([(prefix + ".weight", (co, ci, k, k), "param"), (prefix + ".bias", (co,), "param")])

    def dense_block(prefix, c, n):
        for i in range(n):
            bn(f"{prefix}.layers.{i}.norm", c + i * g)
            conv(f"{prefix}.layers.{i}.conv", g, c + i * g, 3)

    p = "featureExtractor."
    conv(p + "firstconv", cfg.first_conv, cfg.in_channels, 3)
    c = cfg.first_conv
    skips = []
    down, tdown = [], []
    for i, n in enumerate(cfg.down_blocks):
        down.append((f"{p}denseBlocksDown.{i}", c, n))
        c += g * n
        skips.insert(0, c)
        tdown.append((f"{p}transDownBlocks.{i}", c))
    # registration order in FCDenseNetFeatureExtractor.__init__: firstconv, denseBlocksDown, transDownBlocks,
    # bottleneck, transUpBlocks, denseBlocksUp (tiramisu.py:33-87)
    for a in down:
        dense_block(*a)
    for pref, cc in tdown:
        bn(pref + ".norm", cc)
        conv(pref + ".conv", cc, cc, 1)
    dense_block(p + "bottleneck.bottleneck", c, cfg.bottleneck_layers)
    prev = g * cfg.bottleneck_layers
    ups, dups = [], []
    for i, n in enumerate(cfg.up_blocks):
        ups.append((f"{p}transUpBlocks.{i}.convTrans", prev))
        cur = prev + skips[i]
        dups.append((f"{p}denseBlocksUp.{i}", cur, n))
        prev = g * n
        last_c = cur + g * n
    for pref, cc in ups:
        spec.extend([(pref + ".weight", (cc, cc, 3, 3), "param"), (pref + ".bias", (cc,), "param")])
    for a in dups:
        dense_block(*a)
    conv("classifier.finalConv", num_cls, last_c, 1)
    return spec


def init_state_dict(cfg: NetConfig = FCDENSENET67, num_cls: int = 4, seed: int = 42, dtype=torch.float32):
    """Deterministic random init with the *shapes* of the reference (not torch's default init stream):
    kaiming-uniform-like bounds so activations stay O(1).  Used where /root/reference is unavailable."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape, kind in state_dict_spec(cfg, num_cls):
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.zeros((), dtype=torch.int64)
        elif name.endswith("running_mean"):
            sd[name] = torch.zeros(shape, dtype=dtype)
        elif name.endswith("running_var"):
            sd[name] = torch.ones(shape, dtype=dtype)
        elif ".norm.weight" in name:
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=gen, dtype=dtype)
        elif ".norm.bias" in name:
            sd[name] = 0.1 * torch.randn(shape, generator=gen, dtype=dtype)
        elif name.endswith(".weight"):
            if "convTrans" in name:
                fan_in = shape[1] * shape[2] * shape[3]
            else:
                fan_in = shape[1] * shape[2] * shape[3]
            bound = (1.0 / fan_in) ** 0.5
            sd[name] = (torch.rand(shape, generator=gen, dtype=dtype) * 2 - 1) * bound * (3.0 ** 0.5)
        else:  # conv bias
            sd[name] = (torch.rand(shape, generator=gen, dtype=dtype) * 2 - 1) * 0.05
    return sd
