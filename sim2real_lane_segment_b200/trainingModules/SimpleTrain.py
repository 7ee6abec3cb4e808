"""Mirror of rightLaneNetwork/trainingModules/SimpleTrain.py (SimpleTrainModule — the module train.py builds for
``--trainType sim`` and ``st``, train.py:46-53)."""
import torch
from torch.optim.lr_scheduler import CosineAnnealingLR

from ..optim import FusedAdamW
from ..runtime import _FusedLossFn
from .TrainingBase import TrainingBase, getClassWeight  # noqa: F401  (re-exported like the reference)


class SimpleTrainModule(TrainingBase):
    def training_step(self, batch, batch_idx):
        x, y = batch
        rt = self.runtime()

        # Forward propagation, loss calculation: cross_entropy(probabilities, y, weight=1/count) (SimpleTrain.py:15-16)
        tok = self.featureExtractor(x)
        loss, cm = _FusedLossFn.apply(rt._anchor, tok, y.contiguous(), True)

        # Accuracy calculation (SimpleTrain.py:19-20) from the on-device confusion counts
        train_acc = cm.diag().sum().float() / cm.sum().float() * 100

        self.log('tr_loss', loss)
        self.log('tr_acc', train_acc, prog_bar=True)
        return loss

    def configure_optimizers(self):
        self.prepare()
        optimizer = FusedAdamW(self.parameters(), lr=self.lr, weight_decay=self.decay)
        scheduler = CosineAnnealingLR(optimizer, 25, eta_min=self.lr / self.lrRatio)
        return [optimizer], [scheduler]
