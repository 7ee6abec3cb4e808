"""Mirror of rightLaneNetwork/trainingModules/MMETrainingModule.py (minimax-entropy domain adaptation, two optimizers)."""
import torch
from torch.optim.lr_scheduler import CosineAnnealingLR

from ..models.FCDenseNet.tiramisu import grad_reverse
from ..optim import FusedAdamW, FusedSGD
from ..runtime import _FusedLossFn
from .TrainingBase import TrainingBase, getClassWeight  # noqa: F401


def adentropy(output, lamda=1.0):
    return lamda * torch.mean(torch.sum(output * (torch.log(output + 1e-5)), 1))


class MMETrainingModule(TrainingBase):
    def configure_optimizers(self):
        self.prepare()
        optimizerF = FusedAdamW(self.parameters(), lr=self.lr, weight_decay=self.decay)
        optimizerG = FusedSGD([
            {'params': self.featureExtractor.parameters(), 'lr': self.lr / 3},
            {'params': self.classifier.parameters(), 'lr': self.lr}
        ], lr=self.lr, weight_decay=self.decay, momentum=0.9, nesterov=True)
        lr_schedulerF = CosineAnnealingLR(optimizerF, T_max=25, eta_min=self.lr * 1e-3)
        lr_schedulerG = CosineAnnealingLR(optimizerG, T_max=25, eta_min=self.lr * 1e-3)
        return [optimizerG, optimizerF], [lr_schedulerG, lr_schedulerF]

    def training_step(self, batch, batch_idx, optimizer_idx=0):
        x_labelled, x_unlabelled, labels, _ = batch

        if optimizer_idx == 0:  # We are unlabelled optimizer -> maximize entropy
            outputs = self.featureExtractor(x_unlabelled)
            outputs = grad_reverse(outputs)
            outputs = self.classifier(outputs)       # probabilities; backward enters the head with d(loss)/d(probs)
            loss = adentropy(outputs, lamda=0.1)
        if optimizer_idx == 1:  # We are labelled optimizer -> minimize entropy
            tok = self.featureExtractor(x_labelled)
            loss, _ = _FusedLossFn.apply(self.runtime()._anchor, tok, labels.contiguous(), True)

        return loss
