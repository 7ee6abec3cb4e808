"""Drop-in mirror of rightLaneNetwork/models/FCDenseNet/tiramisu.py: same public names, constructor signatures,
attribute names and state_dict keys; the arithmetic runs in librightlane (sm_100a) through ``runtime.Runtime``.

  GradReverse / grad_reverse            tiramisu.py:7-18
  FCDenseNetFeatureExtractor            tiramisu.py:21-109   (forward :89-106)
  FCDenseNetClassifier                  tiramisu.py:112-125
  FCDenseNet67Base / FCDenseNet67Classifier   tiramisu.py:178-194
Only the FC-DenseNet67 configuration (the one TrainingBase builds, TrainingBase.py:32-33) has a CUDA schedule.
"""
import torch
import torch.nn as nn

from ...runtime import FeatureToken, Runtime, _ProbsFn
from .layers import Bottleneck, DenseBlock, TransitionDown, TransitionUp, _Conv


def grad_reverse(x):
    """identity forward, negated gradient (tiramisu.py:7-18); on a feature token it flips the sign the fused head
    backward applies to d(features)"""
    if isinstance(x, FeatureToken):
        return x.reversed()
    return _GradReverse.apply(x)


class _GradReverse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.neg()


GradReverse = _GradReverse


class FCDenseNetFeatureExtractor(nn.Module):
    def __init__(self, in_channels=3, down_blocks=(5, 5, 5, 5, 5), up_blocks=(5, 5, 5, 5, 5), bottleneck_layers=5,
                 growth_rate=16, out_chans_first_conv=48):
        super().__init__()
        cfg = (in_channels, tuple(down_blocks), tuple(up_blocks), bottleneck_layers, growth_rate, out_chans_first_conv)
        if cfg != (3, (5,) * 5, (5,) * 5, 5, 16, 48):
            raise NotImplementedError("librightlane implements the FC-DenseNet67 configuration only (see DESIGN.md)")
        self.down_blocks, self.up_blocks = down_blocks, up_blocks
        self.firstconv = _Conv(in_channels, out_chans_first_conv, 3)
        cur = out_chans_first_conv
        skips = []
        self.denseBlocksDown = nn.ModuleList([])
        self.transDownBlocks = nn.ModuleList([])
        for n in down_blocks:
            self.denseBlocksDown.append(DenseBlock(cur, growth_rate, n))
            cur += growth_rate * n
            skips.insert(0, cur)
            self.transDownBlocks.append(TransitionDown(cur))
        self.bottleneck = Bottleneck(cur, growth_rate, bottleneck_layers)
        prev = growth_rate * bottleneck_layers
        self.transUpBlocks = nn.ModuleList([])
        self.denseBlocksUp = nn.ModuleList([])
        for i, n in enumerate(up_blocks):
            self.transUpBlocks.append(TransitionUp(prev, prev))
            cur = prev + skips[i]
            self.denseBlocksUp.append(DenseBlock(cur, growth_rate, n, upsample=(i != len(up_blocks) - 1)))
            prev = growth_rate * n
            cur += prev
        self.featureChannels = cur
        object.__setattr__(self, "_rt", Runtime(self))

    def getFeatureChannels(self):
        return self.featureChannels

    def runtime(self):
        return self._rt

    def forward(self, x):
        """returns a FeatureToken standing for F.normalize(features) (tiramisu.py:105); feed it to the classifier"""
        return self._rt.run_extractor(x)

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._rt.params = None          # storage moved (.cuda(), .to()): re-flatten lazily
        self._rt.engines.clear()
        return out

    def _load_from_state_dict(self, *a, **k):
        super()._load_from_state_dict(*a, **k)
        self._rt.mark_weights_changed()


class FCDenseNetClassifier(nn.Module):
    def __init__(self, in_channels, n_classes, temperature=0.05, kernel_size=1):
        super().__init__()
        if kernel_size != 1 or in_channels != 288 or abs(temperature - 0.05) > 1e-12 or not (2 <= n_classes <= 4):
            raise NotImplementedError("fused head: 1x1 classifier over 288 features, T=0.05, 2..4 classes")
        self.finalConv = _Conv(in_channels, n_classes, 1)
        self.softmax = nn.Softmax(dim=1)
        self.T = temperature
        self.n_classes = n_classes

    def forward(self, x, useSoftmax=True):
        if not isinstance(x, FeatureToken):
            raise TypeError("FCDenseNetClassifier consumes the FeatureToken returned by FCDenseNetFeatureExtractor")
        if not useSoftmax:
            raise NotImplementedError("useSoftmax=False is not used by any reference call site")
        rt = x.rt
        if rt.classifier_ref is None or rt.classifier_ref() is not self:
            raise RuntimeError("attach the classifier before the extractor forward: extractor.runtime().attach_classifier(cls)")
        if x.train and torch.is_grad_enabled():
            return _ProbsFn.apply(rt._anchor, x)
        return rt.run_head(x)[0]

    def _load_from_state_dict(self, *a, **k):
        super()._load_from_state_dict(*a, **k)
        self.__dict__.get("_rt_notify", lambda: None)()


class FCDenseNet(nn.Module):
    def __init__(self, in_channels=3, down_blocks=(5, 5, 5, 5, 5), up_blocks=(5, 5, 5, 5, 5), bottleneck_layers=5,
                 growth_rate=16, out_chans_first_conv=48, n_classes=12, kernel_size=1):
        super().__init__()
        self.featureExtractor = FCDenseNetFeatureExtractor(in_channels, down_blocks, up_blocks, bottleneck_layers,
                                                           growth_rate, out_chans_first_conv)
        self.classifier = FCDenseNetClassifier(self.featureExtractor.getFeatureChannels(), n_classes,
                                               kernel_size=kernel_size)
        bind(self.featureExtractor, self.classifier)

    def forward(self, x):
        return self.classifier(self.featureExtractor(x))


def bind(extractor, classifier):
    """tie a classifier to an extractor's runtime (TrainingBase does this for its two sub-modules)"""
    rt = extractor.runtime()
    rt.attach_classifier(classifier)
    classifier.__dict__["_rt_notify"] = rt.mark_weights_changed
    return rt


def FCDenseNet67(n_classes):
    return FCDenseNet(n_classes=n_classes)


def FCDenseNet67Base():
    return FCDenseNetFeatureExtractor()


def FCDenseNet67Classifier(n_classes):
    return FCDenseNetClassifier(288, n_classes)


def _unsupported(name):
    def f(*a, **k):
        raise NotImplementedError(f"{name}: only FC-DenseNet67 has a CUDA schedule in this release (DESIGN.md, scope)")
    return f


FCDenseNet57, FCDenseNet103 = _unsupported("FCDenseNet57"), _unsupported("FCDenseNet103")
FCDenseNet57Base, FCDenseNet57Classifier = _unsupported("FCDenseNet57Base"), _unsupported("FCDenseNet57Classifier")
