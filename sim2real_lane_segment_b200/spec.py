"""Names, shapes and flat-arena offsets of every tensor in the reference's state_dict
(TrainingBase: featureExtractor = FCDenseNet67Base(), classifier = FCDenseNet67Classifier(num_cls);
reference trainingModules/TrainingBase.py:32-33, models/FCDenseNet/tiramisu.py:21-125).

The C side (csrc/rl_plan.cu) lays parameters out in exactly this order; ``check_against_plan`` asserts it.
"""
from collections import OrderedDict

GROWTH, FIRST, NLAYERS, UP = 16, 48, 5, 80
SKIPS = (128, 208, 288, 368, 448)          # channels after each down block
FEATURES = 288


def _bn(prefix, c):
    return [(f"{prefix}.weight", (c,), "param"), (f"{prefix}.bias", (c,), "param"),
            (f"{prefix}.running_mean", (c,), "bn"), (f"{prefix}.running_var", (c,), "bn"),
            (f"{prefix}.num_batches_tracked", (), "count")]


def _dense_block(prefix, cin):
    out = []
    for j in range(NLAYERS):
        c = cin + GROWTH * j
        out += _bn(f"{prefix}.layers.{j}.norm", c)
        out += [(f"{prefix}.layers.{j}.conv.weight", (GROWTH, c, 3, 3), "param"),
                (f"{prefix}.layers.{j}.conv.bias", (GROWTH,), "param")]
    return out


def entries(num_cls=4):
    """[(name, shape, kind)] in nn.Module registration order; kind in {'param', 'bn', 'count'}"""
    fe = "featureExtractor."
    e = [(fe + "firstconv.weight", (FIRST, 3, 3, 3), "param"), (fe + "firstconv.bias", (FIRST,), "param")]
    cin = FIRST
    for i, cs in enumerate(SKIPS):
        e += _dense_block(f"{fe}denseBlocksDown.{i}", cin)
        cin = cs
    for i, cs in enumerate(SKIPS):
        e += _bn(f"{fe}transDownBlocks.{i}.norm", cs)
        e += [(f"{fe}transDownBlocks.{i}.conv.weight", (cs, cs, 1, 1), "param"),
              (f"{fe}transDownBlocks.{i}.conv.bias", (cs,), "param")]
    e += _dense_block(f"{fe}bottleneck.bottleneck", SKIPS[-1])
    for i in range(5):
        e += [(f"{fe}transUpBlocks.{i}.convTrans.weight", (UP, UP, 3, 3), "param"),
              (f"{fe}transUpBlocks.{i}.convTrans.bias", (UP,), "param")]
    for i in range(5):
        e += _dense_block(f"{fe}denseBlocksUp.{i}", UP + SKIPS[4 - i])
    e += [("classifier.finalConv.weight", (num_cls, FEATURES, 1, 1), "param"),
          ("classifier.finalConv.bias", (num_cls,), "param")]
    return e


def _numel(shape):
    n = 1
    for s in shape:
        n *= s
    return n


def layout(num_cls=4):
    """(params, bn): OrderedDict name -> (offset, shape) into the two flat fp32 arenas; counts: list of names"""
    params, bn, counts = OrderedDict(), OrderedDict(), []
    po = bo = 0
    for name, shape, kind in entries(num_cls):
        if kind == "param":
            params[name] = (po, shape)
            po += _numel(shape)
        elif kind == "bn":
            bn[name] = (bo, shape)
            bo += _numel(shape)
        else:
            counts.append(name)
    return params, bn, counts, po, bo
