"""shared helpers for the GPU parity tests (CUDA path vs oracle on identical seeded inputs)"""
import torch

from sim2real_lane_segment_b200 import spec


def flatten_state_dict(sd, num_cls, device):
    params_l, bn_l, _, npar, nbn = spec.layout(num_cls)
    params = torch.empty(npar, dtype=torch.float32)
    bn = torch.empty(nbn, dtype=torch.float32)
    for name, (off, shape) in params_l.items():
        params[off:off + sd[name].numel()] = sd[name].detach().reshape(-1).float()
    for name, (off, shape) in bn_l.items():
        bn[off:off + sd[name].numel()] = sd[name].detach().reshape(-1).float()
    return params.to(device), bn.to(device)


def unflatten_params(flat, num_cls):
    params_l, _, _, _, _ = spec.layout(num_cls)
    out = {}
    for name, (off, shape) in params_l.items():
        n = 1
        for s in shape:
            n *= s
        out[name] = flat[off:off + n].view(shape)
    return out


def ref_to_phys(t, cskip):
    """reference up-block channel order [up80 | skip | new] -> physical arena order [skip | up80 | new]"""
    return torch.cat([t[:, 80:80 + cskip], t[:, :80], t[:, 80 + cskip:]], 1)


def rel_err(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def layer_names():
    """state_dict prefixes of the layers in the C plan's forward order (csrc/rl_plan.cu)"""
    names = ["featureExtractor.firstconv"]
    for l in range(5):
        names += [f"featureExtractor.denseBlocksDown.{l}.layers.{j}" for j in range(5)]
        names.append(f"featureExtractor.transDownBlocks.{l}")
    names += [f"featureExtractor.bottleneck.bottleneck.layers.{j}" for j in range(5)]
    for i in range(5):
        names.append(f"featureExtractor.transUpBlocks.{i}")
        names += [f"featureExtractor.denseBlocksUp.{i}.layers.{j}" for j in range(5)]
    return names


def install_drop_masks(eng, masks, p_drop=0.2):
    """write oracle keep masks (dict '<layer prefix>.drop' -> [B, C] of 0/1) into the plan's dropout scale blocks"""
    for idx, name in enumerate(layer_names()):
        key = name + ".drop"
        if key in masks:
            eng.layer_mask_view(idx).copy_(masks[key].to("cuda") / (1.0 - p_drop))
