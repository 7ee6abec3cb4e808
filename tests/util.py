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
