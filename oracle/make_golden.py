"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.pt by running the *reference's own model code*
(/root/reference/rightLaneNetwork/models/FCDenseNet, imported unchanged) in this build container.

    python -m oracle.make_golden

Weights are not stored (13.8 MB): they are regenerated from ``oracle.fcdensenet.init_state_dict(seed)``
and loaded into the reference modules with ``load_state_dict``; the fixture keeps inputs, outputs, loss,
gradient norms and a few complete small gradients.  /root/reference does not exist on the GPU box, so
tests only read the committed fixtures there.
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/rightLaneNetwork"

from oracle import fcdensenet as net  # noqa: E402
from oracle import steps  # noqa: E402


class _MaskMul(torch.nn.Module):
    """stand-in for nn.Dropout2d with an injected keep mask (reference layers.py:12,51 use p=0.2)"""

    def __init__(self, keep):
        super().__init__()
        self.keep = keep

    def forward(self, x):
        return x * self.keep[:, :, None, None].to(x.dtype) / (1.0 - net.DROP_P) if self.training else x


def reference_modules(sd, num_cls):
    sys.path.insert(0, REF)
    from models.FCDenseNet.tiramisu import FCDenseNet67Base, FCDenseNet67Classifier
    fe, cl = FCDenseNet67Base(), FCDenseNet67Classifier(num_cls)
    fe.load_state_dict({k[len("featureExtractor."):]: v for k, v in sd.items() if k.startswith("featureExtractor.")})
    cl.load_state_dict({k[len("classifier."):]: v for k, v in sd.items() if k.startswith("classifier.")})
    return fe, cl


def make_drop_masks(B, seed, cfg=net.FCDENSENET67, num_cls=4):
    g = torch.Generator().manual_seed(seed)
    masks = {}
    for name, shape, _ in net.state_dict_spec(cfg, num_cls):
        if name.endswith(".conv.weight"):
            key = name[:-len(".conv.weight")] + ".drop"
            masks[key] = (torch.rand(B, shape[0], generator=g) >= net.DROP_P).float()
    return masks


def install_masks(fe, masks):
    for name, mod in list(fe.named_modules()):
        if isinstance(mod, torch.nn.Dropout2d):
            parent = fe.get_submodule(name.rsplit(".", 1)[0])
            setattr(parent, "drop", _MaskMul(masks["featureExtractor." + name]))


def run_case(tag, B, H, W, num_cls, seed, with_masks):
    sd = net.init_state_dict(num_cls=num_cls, seed=seed)
    x, y = steps.synthetic_batch(B, H, W, num_cls, seed=seed)
    fe, cl = reference_modules(sd, num_cls)
    out = {"B": B, "H": H, "W": W, "num_cls": num_cls, "seed": seed, "x": x, "y": y.to(torch.uint8)}
    # eval forward
    fe.eval(); cl.eval()
    with torch.no_grad():
        probs = cl(fe(x))
    out["eval_probs"] = probs
    out["eval_mask"] = probs.argmax(1).to(torch.uint8)
    out["eval_loss"] = F.cross_entropy(probs, y)
    # train forward/backward (SimpleTrain.py:11-25)
    masks = make_drop_masks(B, seed + 1, num_cls=num_cls) if with_masks else None
    fe.train(); cl.train()
    if with_masks:
        install_masks(fe, masks)
        out["drop_seed"] = seed + 1
    else:
        for m in fe.modules():
            if isinstance(m, torch.nn.Dropout2d):
                m.p = 0.0
    probs = cl(fe(x))
    w = steps.get_class_weight(y, num_cls)
    loss = F.cross_entropy(probs, y, weight=w)
    loss.backward()
    out["train_probs"] = probs.detach()
    out["train_loss"] = loss.detach()
    out["class_weight"] = w
    grads = {"featureExtractor." + k: p.grad for k, p in fe.named_parameters()}
    grads.update({"classifier." + k: p.grad for k, p in cl.named_parameters()})
    out["grad_norms"] = {k: float(v.norm()) for k, v in grads.items()}
    keep = ["featureExtractor.firstconv.weight", "featureExtractor.firstconv.bias", "classifier.finalConv.weight",
            "classifier.finalConv.bias", "featureExtractor.denseBlocksDown.0.layers.0.norm.weight",
            "featureExtractor.denseBlocksDown.0.layers.0.conv.weight", "featureExtractor.denseBlocksUp.4.layers.4.conv.weight",
            "featureExtractor.transDownBlocks.0.norm.bias", "featureExtractor.transUpBlocks.0.convTrans.bias",
            "featureExtractor.bottleneck.bottleneck.layers.2.conv.weight"]
    out["grads"] = {k: grads[k].clone() for k in keep}
    out["running_after"] = {
        "featureExtractor.denseBlocksDown.0.layers.0.norm.running_mean":
            fe.denseBlocksDown[0].layers[0].norm.running_mean.clone(),
        "featureExtractor.denseBlocksUp.4.layers.4.norm.running_var":
            fe.denseBlocksUp[4].layers[4].norm.running_var.clone()}
    path = os.path.join(ROOT, "tests", "golden", f"{tag}.pt")
    torch.save(out, path)
    print(tag, "loss", float(loss), "eval_loss", float(out["eval_loss"]), os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    torch.manual_seed(0)
    run_case("ref_b2_40x56_c4", 2, 40, 56, 4, 7, with_masks=False)
    run_case("ref_b2_40x56_c4_drop", 2, 40, 56, 4, 11, with_masks=True)
    run_case("ref_b2_64x96_c2", 2, 64, 96, 2, 13, with_masks=False)
    run_case("ref_b1_120x160_c4", 1, 120, 160, 4, 42, with_masks=False)
