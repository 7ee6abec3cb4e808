"""TEST INFRASTRUCTURE ONLY — restatement of the reference's training / evaluation step and metrics.

Follows /root/reference/rightLaneNetwork/trainingModules/:
  TrainingBase.py:12-23   getClassWeight    w[c] = 1 / count_c  (absent class -> inf)
  SimpleTrain.py:11-25    training_step     cross_entropy(softmax-probabilities, y, weight) ; argmax ; accuracy*100
  SimpleTrain.py:27-30    configure_optimizers  AdamW(lr, weight_decay=decay) + CosineAnnealingLR(T_max=25, eta_min=lr/lrRatio)
  TrainingBase.py:79-95   evaluate_batch    unweighted CE on probabilities; accuracy, dice_score, iou (each * batch size)
  TrainingBase.py:97-110  summarize_evaluation_results
  MMETrainingModule.py:10-11 adentropy
and rightLaneNetwork/test.py:88-101 (confusion_matrix with num_classes=4).
pytorch_lightning==1.2.1 (requirements.txt:9) is not installable here; its four metric functions are
restated from that release's published semantics — all are functions of the integer
confusion matrix CM[t, p] = #pixels with label t and prediction p.
"""
import math

import torch
import torch.nn.functional as F

from . import fcdensenet as net


def get_class_weight(targets, max_classes=None):
    elements, counts = torch.unique(targets, sorted=True, return_counts=True)
    if max_classes:
        assert max_classes > int(elements.max()), f"Found more label classes than given maxClasses={max_classes}"
    else:
        max_classes = int(elements.max()) + 1
    per_class = torch.zeros(max_classes, dtype=torch.float)
    per_class[elements.long().cpu()] = counts.float().cpu()
    return torch.reciprocal(per_class)


# ---------------------------------------------------------------------------------------------- metrics
def confusion_matrix(pred, target, num_classes):
    """rows = label, cols = prediction (matches the print-out at test.py:113)"""
    idx = target.reshape(-1).long() * num_classes + pred.reshape(-1).long()
    return torch.bincount(idx, minlength=num_classes * num_classes).reshape(num_classes, num_classes)


def accuracy_from_cm(cm):
    return cm.diag().sum().float() / cm.sum().float()


def iou_from_cm(cm, k):
    """PL-1.2.1 iou(pred, target): num_classes = max(pred.max(), target.max()) + 1 = k (data dependent);
    absent_score 0.0, mean over the k classes."""
    c = cm[:k, :k]
    inter = c.diag().float()
    union = c.sum(0).float() + c.sum(1).float() - inter
    scores = torch.where(union == 0, torch.zeros_like(inter), inter / union.clamp(min=1))
    return scores.mean()


def dice_from_cm(cm):
    """PL-1.2.1 dice_score(probs, target, bg=False, nan_score=0, no_fg_score=0): classes 1..C-1."""
    C = cm.shape[0]
    scores = torch.zeros(C - 1)
    for i in range(1, C):
        if cm[i, :].sum() == 0:       # class absent from the target
            continue
        tp = cm[i, i]
        fp = cm[:, i].sum() - tp
        fn = cm[i, :].sum() - tp
        denom = (2 * tp + fp + fn).float()
        scores[i - 1] = (2 * tp).float() / denom if denom != 0 else 0.0
    return scores.mean()


def data_dependent_k(pred, target):
    return int(max(int(pred.max()), int(target.max()))) + 1


def accuracy(pred, target):
    return (pred == target).sum().float() / target.numel()


def iou(pred, target):
    k = data_dependent_k(pred, target)
    return iou_from_cm(confusion_matrix(pred, target, k), k)


def dice_score(probs, target):
    C = probs.shape[1]
    return dice_from_cm(confusion_matrix(probs.argmax(1), target, C))


# ---------------------------------------------------------------------------------------------- steps
def training_step(sd, x, y, num_cls, cfg=net.FCDENSENET67, drop_masks=None, update_running=True):
    """Returns (loss, train_acc_percent, labels_hat, new_running_stats); differentiable w.r.t. sd tensors."""
    probs, new_running = net.forward(sd, x, cfg, train=True, drop_masks=drop_masks, update_running=update_running)
    w = get_class_weight(y, num_cls).to(probs.device, probs.dtype)
    loss = F.cross_entropy(probs, y, weight=w)
    labels_hat = probs.argmax(1)
    return loss, accuracy(labels_hat, y) * 100, labels_hat, new_running


def evaluate_batch(sd, x, y, cfg=net.FCDENSENET67):
    with torch.no_grad():
        probs, _ = net.forward(sd, x, cfg, train=False)
        loss = F.cross_entropy(probs, y)
        labels_hat = probs.argmax(1)
    wgt = x.shape[0]
    return {"loss": loss * wgt, "acc": accuracy(labels_hat, y) * wgt, "dice": dice_score(probs, y) * wgt,
            "iou": iou(labels_hat, y) * wgt, "weight": wgt, "labels_hat": labels_hat, "probs": probs}


def summarize_evaluation_results(outputs):
    tw = sum(o["weight"] for o in outputs)
    return {"loss": sum(o["loss"] for o in outputs) / tw, "acc": sum(o["acc"] for o in outputs) / tw * 100.0,
            "dice": sum(o["dice"] for o in outputs) / tw, "iou": sum(o["iou"] for o in outputs) / tw * 100.0}


class _GradReverse(torch.autograd.Function):
    """models/FCDenseNet/tiramisu.py:7-18: identity forward, negated gradient"""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.neg()


def mme_unlabelled_loss(sd, x_unl, cfg=net.FCDENSENET67, drop_masks=None):
    """MMETrainingModule.training_step, optimizer_idx 0 (MMETrainingModule.py:28-32): entropy maximisation on the
    unlabelled stream with the gradient reversed between extractor and classifier"""
    feats, new_running = net.feature_extractor(sd, x_unl, cfg, train=True, drop_masks=drop_masks, update_running=True)
    probs = net.classifier(sd, _GradReverse.apply(feats))
    return adentropy(probs, lamda=0.1), new_running


def adentropy(probs, lamda=1.0):
    return lamda * torch.mean(torch.sum(probs * torch.log(probs + 1e-5), 1))


def cosine_lr(base_lr, eta_min, t, t_max=25):
    """closed form of torch CosineAnnealingLR at epoch t"""
    return eta_min + (base_lr - eta_min) * (1 + math.cos(math.pi * t / t_max)) / 2


class Trainer:
    """Runs SimpleTrainModule's optimisation loop on CPU: fwd + weighted CE + bwd + AdamW."""

    def __init__(self, sd, num_cls=4, lr=1e-3, decay=1e-4, cfg=net.FCDENSENET67):
        self.cfg, self.num_cls = cfg, num_cls
        self.sd = {k: v.clone() for k, v in sd.items()}
        self.param_names = [n for n, _, kind in net.state_dict_spec(cfg, num_cls) if kind == "param"]
        for n in self.param_names:
            self.sd[n].requires_grad_(True)
        self.opt = torch.optim.AdamW([self.sd[n] for n in self.param_names], lr=lr, weight_decay=decay)

    def step(self, x, y, drop_masks=None):
        self.opt.zero_grad(set_to_none=True)
        loss, acc, labels_hat, new_running = training_step(self.sd, x, y, self.num_cls, self.cfg, drop_masks)
        loss.backward()
        self.opt.step()
        with torch.no_grad():
            for k, v in new_running.items():
                self.sd[k].copy_(v)
            for k in self.sd:
                if k.endswith("num_batches_tracked"):
                    self.sd[k] += 1
        return float(loss), float(acc), labels_hat

    def grads(self):
        return {n: self.sd[n].grad for n in self.param_names}


# ---------------------------------------------------------------------------------------------- data
def synthetic_batch(B, H, W, num_cls=4, seed=42, binary=False):
    """x ~ N(0,1) (post-Normalize images), y piecewise-constant: class 1 inside a random
    trapezoid in the lower half (right lane), optional class-2 strip and class-3 blobs."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, H, W, generator=g)
    y = torch.zeros(B, H, W, dtype=torch.int64)
    ys = torch.arange(H).view(H, 1).float()
    xs = torch.arange(W).view(1, W).float()
    for b in range(B):
        r = torch.rand(8, generator=g)
        top = H * (0.45 + 0.1 * float(r[0]))
        cx = W * (0.35 + 0.3 * float(r[1]))
        t = ((ys - top) / max(H - top, 1)).clamp(0, 1)
        half = W * (0.03 + 0.22 * t)
        right = (ys >= top) & (xs >= cx) & (xs < cx + 2 * half)
        y[b][right] = 1
        if not binary and num_cls > 2:
            left = (ys >= top) & (xs < cx) & (xs >= cx - 2 * half)
            y[b][left] = 2
            if num_cls > 3 and float(r[2]) > 0.3:
                oy, ox = H * (0.5 + 0.4 * float(r[3])), W * float(r[4])
                blob = ((ys - oy) ** 2 + (xs - ox) ** 2) < (0.06 * H) ** 2
                y[b][blob] = 3
    return x, y
