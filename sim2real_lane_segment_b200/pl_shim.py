"""Minimal stand-in for the slice of pytorch_lightning==1.2.1 the reference's training modules touch
(reference requirements.txt:9; trainingModules/TrainingBase.py:4-5,26-29,64-77; train.py:39-67).

pytorch_lightning 1.2.x is not installable in this image, so ``trainingModules`` imports these names when the real
package is absent.  Only protocol surface is provided: LightningModule (save_hyperparameters, log, device, hparams,
load_from_checkpoint), seed_everything and a tiny single-process Trainer good enough to drive fit()/test() loops in
tests and examples.  The CUDA path does not depend on it.
"""
import argparse
import random
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn


def seed_everything(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    return seed


class LightningModule(nn.Module):
    def __init__(self):
        super().__init__()
        self.hparams = SimpleNamespace()
        self.logged = {}
        self.trainer = None

    def save_hyperparameters(self, *names):
        import inspect
        frame = inspect.currentframe().f_back
        for n in names:
            if n in frame.f_locals:
                setattr(self.hparams, n, frame.f_locals[n])

    def log(self, name, value, **kwargs):
        self.logged[name] = value

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, **kwargs):
        ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu")
        hp = dict(ckpt.get("hyper_parameters", {}))
        hp.update(kwargs)
        model = cls(**hp)
        model.load_state_dict(ckpt["state_dict"] if "state_dict" in ckpt else ckpt)
        return model

    def _load_model_state(self, ckpt, **kwargs):
        self.load_state_dict(ckpt["state_dict"])
        return self


class Trainer:
    """single-process fit/test loop honouring the hooks the reference modules define"""

    def __init__(self, max_epochs=1, gpus=0, limit_train_batches=None, **_):
        self.max_epochs, self.gpus, self.limit_train_batches = max_epochs, gpus, limit_train_batches

    @staticmethod
    def add_argparse_args(parser):
        g = parser.add_argument_group("Trainer")
        g.add_argument("--max_epochs", type=int, default=1)
        g.add_argument("--gpus", type=int, default=0)
        return parser

    @classmethod
    def from_argparse_args(cls, args, **kw):
        return cls(max_epochs=getattr(args, "max_epochs", 1), gpus=getattr(args, "gpus", 0), **kw)

    def fit(self, model, train_dataloader=None, val_dataloaders=None, datamodule=None):
        if datamodule is not None:
            datamodule.setup("fit") if hasattr(datamodule, "setup") else None
            train_dataloader = datamodule.train_dataloader()
            val_dataloaders = datamodule.val_dataloader() if hasattr(datamodule, "val_dataloader") else None
        model.trainer = self
        self.model = model
        self.current_epoch, self.global_step = 0, 0
        opts, scheds = model.configure_optimizers()
        dev = model.device
        for epoch in range(self.max_epochs):
            model.train()
            for bi, batch in enumerate(train_dataloader):
                if self.limit_train_batches is not None and bi >= self.limit_train_batches:
                    break
                batch = [b.to(dev) if torch.is_tensor(b) else b for b in batch]
                for oi, opt in enumerate(opts):
                    opt.zero_grad()
                    loss = model.training_step(batch, bi, oi) if len(opts) > 1 else model.training_step(batch, bi)
                    loss.backward()
                    opt.step()
                self.global_step += 1
            for s in scheds:
                s.step()
            self.current_epoch = epoch + 1
            if val_dataloaders is not None:
                model.eval()
                outs = []
                with torch.no_grad():
                    for bi, batch in enumerate(val_dataloaders):
                        batch = [b.to(dev) if torch.is_tensor(b) else b for b in batch]
                        outs.append(model.validation_step(batch, bi))
                model.validation_epoch_end(outs)
        return model

    def save_checkpoint(self, filepath, model=None):
        """checkpoint file with the keys pytorch-lightning 1.2.1 writes and `load_from_checkpoint` reads (test.py:28-30,
        makeDemoVideo.py:52-54): state_dict (434 entries, the reference's names), hyper_parameters, epoch, global_step"""
        model = model if model is not None else getattr(self, "model", None)
        if model is None:
            raise ValueError("save_checkpoint: no model (call fit() first or pass model=)")
        if hasattr(model, "runtime"):
            torch.cuda.synchronize() if torch.cuda.is_available() else None
        ckpt = {
            "epoch": getattr(self, "current_epoch", 0),
            "global_step": getattr(self, "global_step", 0),
            "pytorch-lightning_version": "1.2.1",
            "state_dict": {k: v.detach().cpu().clone() for k, v in model.state_dict().items()},
            "hyper_parameters": dict(vars(model.hparams)) if hasattr(model, "hparams") else {},
        }
        torch.save(ckpt, filepath)
        return filepath

    def test(self, model, test_dataloaders=None):
        model.eval()
        outs = []
        dev = model.device
        with torch.no_grad():
            for bi, batch in enumerate(test_dataloaders):
                batch = [b.to(dev) if torch.is_tensor(b) else b for b in batch]
                outs.append(model.test_step(batch, bi))
        model.test_epoch_end(outs)
        return [dict(model.logged)]
