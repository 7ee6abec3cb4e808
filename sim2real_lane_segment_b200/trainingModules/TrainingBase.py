"""Mirror of rightLaneNetwork/trainingModules/TrainingBase.py (same names, signatures, logged keys, state_dict).

Differences that are invisible to callers:
  * the network runs in librightlane (sm_100a CUDA); ``forward(x)`` returns the same fp32 probabilities [B,C,H,W];
  * evaluate_batch computes loss / accuracy / dice / IoU from one on-device 4x4 confusion matrix (all four PL-1.2.1
    metrics are functions of it, TrainingBase.py:88-93) instead of four passes over the predictions;
  * getClassWeight stays available with the reference's semantics (TrainingBase.py:12-23) but the training step uses
    the on-device label histogram (no device->host synchronisation per step).
"""
from argparse import ArgumentParser

import torch

try:  # the real Lightning if present, else the protocol shim (pytorch_lightning 1.2.1 is not installable here)
    from pytorch_lightning import LightningModule
except Exception:  # pragma: no cover - depends on the image
    from ..pl_shim import LightningModule

from ..models.FCDenseNet.tiramisu import FCDenseNet67Base, FCDenseNet67Classifier, bind


def getClassWeight(targets, maxClasses: int = None):
    counts = torch.bincount(targets.reshape(-1).long().cpu())
    if maxClasses:
        assert maxClasses > counts.numel() - 1, f"Found more label classes than given maxClasses={maxClasses}"
    else:
        maxClasses = counts.numel()
    countPerClass = torch.zeros(maxClasses, dtype=torch.float)
    countPerClass[:counts.numel()] = counts.float()
    return torch.reciprocal(countPerClass)


def metrics_from_confusion(cm):
    """PL-1.2.1 accuracy / dice_score(bg=False) / iou from CM[label, pred] (int64 [C,C] device tensor); returns
    fp32 device scalars — exact integer ratios evaluated in fp32, as the PL functions do."""
    C = cm.shape[0]
    cmf = cm.float()
    diag = cmf.diag()
    rows, cols = cmf.sum(1), cmf.sum(0)          # label counts, prediction counts
    acc = diag.sum() / cmf.sum()
    # dice over classes 1..C-1: 2tp / (2tp + fp + fn); class absent from the target -> 0
    denom = rows + cols
    dice_c = torch.where((rows > 0) & (denom > 0), 2 * diag / denom.clamp(min=1), torch.zeros_like(diag))
    dice = dice_c[1:].mean()
    # iou over the first K = max(pred.max(), target.max()) + 1 classes; union == 0 -> 0
    present = (rows + cols) > 0
    idx = torch.arange(C, device=cm.device)
    K = (idx * present).max() + 1
    union = rows + cols - diag
    iou_c = torch.where(union > 0, diag / union.clamp(min=1), torch.zeros_like(diag))
    iou = (iou_c * (idx < K)).sum() / K.float()
    return acc, dice, iou


class TrainingBase(LightningModule):
    def __init__(self, lr=1e-3, decay=1e-4, lrRatio=1e3, num_cls=2):
        super().__init__()
        self.save_hyperparameters('lr', 'decay', 'lrRatio')

        # Create network parts
        self.featureExtractor = FCDenseNet67Base()
        self.classifier = FCDenseNet67Classifier(n_classes=num_cls)
        bind(self.featureExtractor, self.classifier)

        # Training parameters
        self.lr = lr
        self.decay = decay
        self.lrRatio = lrRatio

        self.num_cls = num_cls

    @staticmethod
    def add_model_specific_args(parent_parser):
        parser = ArgumentParser(parents=[parent_parser], add_help=False)
        argGroup = parser.add_argument_group('TrainingModule', 'Parameters defining network training')
        argGroup.add_argument('-lr', '--learningRate', type=float, default=1e-3, help="Starting learning rate")
        argGroup.add_argument('--decay', type=float, default=1e-4, help="L2 weight decay value")
        argGroup.add_argument('--lrRatio', type=float, default=1000,
                              help="Ratio of maximum and minimum of learning rate for cosine LR scheduler")
        return parser

    def runtime(self):
        return self.featureExtractor.runtime()

    def prepare(self):
        """flatten parameters into the arenas now (otherwise done lazily by the first forward)"""
        self.runtime().flatten()
        return self

    def forward(self, x):
        x = self.featureExtractor(x)
        x = self.classifier(x)
        return x

    def validation_step(self, batch, batch_idx):
        return self.evaluate_batch(batch)

    def validation_epoch_end(self, outputs):
        logs = self.summarize_evaluation_results(outputs)
        self.log('val_loss', logs['loss'])
        self.log('val_acc', logs['acc'], prog_bar=True, logger=True)
        self.log('val_dice', logs['dice'])
        self.log('val_iou', logs['iou'], prog_bar=True, logger=True)

    def test_step(self, batch, batch_idx):
        return self.evaluate_batch(batch)

    def test_epoch_end(self, outputs):
        logs = self.summarize_evaluation_results(outputs)
        self.log('test_loss', logs['loss'])
        self.log('test_acc', logs['acc'])
        self.log('test_dice', logs['dice'])
        self.log('test_iou', logs['iou'])

    EVAL_CHUNK_PIXELS = 32 * 480 * 640   # largest slice of a batch that goes through one plan (activation arenas ~12 GB)

    def evaluate_batch(self, batch, return_predictions=False, return_probs=None):
        """TrainingBase.evaluate_batch (TrainingBase.py:79-95) / the test.py loop body (test.py:93-101).  Large batches
        (test.py's 256 x 3 x 480 x 640) run through the plan in slices along the batch axis; loss sums and the confusion
        matrix are integer / linear in the slices, so the metrics of the whole batch are exactly those of one pass."""
        x, y = batch
        rt = self.runtime()
        B = x.shape[0]
        if return_probs is None:
            return_probs = return_predictions       # masks + confusion matrix only: return_predictions=True, return_probs=False
        per = max(1, self.EVAL_CHUNK_PIXELS // (x.shape[2] * x.shape[3]))
        n_slices = (B + per - 1) // per
        per = (B + n_slices - 1) // n_slices          # equal slices: one plan shape (plus at most one tail shape)
        loss_sums, cm = None, None
        probs_l, mask_l = [], []
        with torch.no_grad():
            for b0 in range(0, B, per):
                xs, ys = x[b0:b0 + per], y[b0:b0 + per].contiguous()
                tok = self.featureExtractor(xs)
                probs, mask = rt.run_head(tok, y=ys, weighted=False, want_probs=return_probs, want_mask=return_predictions)
                hist, cmi, ls, bad = tok.engine.scalars()
                C = self.num_cls
                loss_sums = ls.clone() if loss_sums is None else loss_sums + ls
                cmi = cmi[:C * C].view(C, C).long()
                cm = cmi.clone() if cm is None else cm + cmi
                if return_predictions:
                    probs_l.append(probs)
                    mask_l.append(mask)
        loss = (loss_sums[0] / loss_sums[1]).float()
        acc, dice, iou = metrics_from_confusion(cm)
        weight = B
        out = {'loss': loss * weight, 'acc': acc * weight, 'dice': dice * weight, 'iou': iou * weight, 'weight': weight}
        if return_predictions:
            out.update(labels_hat=mask_l[0] if len(mask_l) == 1 else torch.cat(mask_l), confusion_matrix=cm)
            if return_probs:
                out.update(probs=probs_l[0] if len(probs_l) == 1 else torch.cat(probs_l))
        return out

    @staticmethod
    def summarize_evaluation_results(outputs):
        total_weight = sum(x['weight'] for x in outputs)
        loss = sum([x['loss'] for x in outputs]) / total_weight
        acc = sum([x['acc'] for x in outputs]) / total_weight * 100.0
        dice = sum([x['dice'] for x in outputs]) / total_weight
        iou = sum([x['iou'] for x in outputs]) / total_weight * 100.0
        return {'loss': loss, 'acc': acc, 'dice': dice, 'iou': iou}
