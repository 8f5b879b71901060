"""Keras callbacks the reference's trainers pass to ``fit``.

/root/reference: Super_resolution/code/train_adaptive_unet.py:604-620 (EarlyStopping,
ModelCheckpoint, BackupAndRestore, TensorBoard); Segmenation/code/train_adaptive_unet.py:415-448
(plus ReduceLROnPlateau).  Same constructor arguments and epoch-level behaviour.
"""
from __future__ import annotations

import json
import math
import os
from pathlib import Path

import numpy as np


class Callback:
    model = None

    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None): pass
    def on_train_end(self, logs=None): pass
    def on_epoch_begin(self, epoch, logs=None): pass
    def on_epoch_end(self, epoch, logs=None): pass
    def initial_epoch(self, requested):
        return requested


def _better(mode, monitor):
    if mode == "max" or (mode == "auto" and any(k in monitor for k in ("acc", "psnr", "dice", "iou", "auc"))):
        return lambda a, b, d=0.0: a > b + d, -math.inf
    return lambda a, b, d=0.0: a < b - d, math.inf


class EarlyStopping(Callback):
    """keras 3 semantics (the reference pins keras==3.3.3): ``wait`` counts epochs without an improvement of more than
    ``min_delta``; training stops when it reaches ``patience`` (never after the very first epoch); with
    ``restore_best_weights`` the best epoch's weights are put back at the END of training -- also when ``fit`` ran out of
    epochs without stopping early."""

    def __init__(self, monitor="val_loss", min_delta=0.0, patience=0, verbose=0, mode="auto", restore_best_weights=False,
                 **kwargs):
        self.monitor, self.min_delta, self.patience, self.verbose = monitor, abs(min_delta), patience, verbose
        self.restore_best_weights = restore_best_weights
        self._cmp, self._worst = _better(mode, monitor)

    def on_train_begin(self, logs=None):
        self.best, self.wait, self.best_weights, self.stopped_epoch, self.best_epoch = self._worst, 0, None, 0, 0

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if cur is None:
            return
        if self.restore_best_weights and self.best_weights is None:      # nothing recorded yet: these are the best so far
            self.best_weights, self.best_epoch = self.model.get_weights(), epoch
        if self._cmp(cur, self.best, self.min_delta):
            self.best, self.wait, self.best_epoch = cur, 0, epoch
            if self.restore_best_weights:
                self.best_weights = self.model.get_weights()
            return
        self.wait += 1
        if self.wait >= self.patience and epoch > 0:
            self.stopped_epoch = epoch
            self.model.stop_training = True

    def on_train_end(self, logs=None):
        if self.stopped_epoch and self.verbose:
            print(f"Epoch {self.stopped_epoch + 1}: early stopping")
        if self.restore_best_weights and self.best_weights is not None:
            if self.verbose:
                print(f"Restoring model weights from the end of the best epoch: {self.best_epoch + 1}.")
            self.model.set_weights(self.best_weights)


class ModelCheckpoint(Callback):
    def __init__(self, filepath, monitor="val_loss", verbose=0, save_best_only=False, save_weights_only=False,
                 mode="auto", **kwargs):
        self.filepath, self.monitor, self.verbose, self.save_best_only = str(filepath), monitor, verbose, save_best_only
        self._cmp, self.best = _better(mode, monitor)

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        path = self.filepath.format(epoch=epoch + 1, **logs)
        if self.save_best_only:
            cur = logs.get(self.monitor)
            if cur is None:
                return
            if not self._cmp(cur, self.best):
                if self.verbose:
                    print(f"\nEpoch {epoch + 1}: {self.monitor} did not improve from {self.best:.5f}")
                return
            if self.verbose:
                print(f"\nEpoch {epoch + 1}: {self.monitor} improved from {self.best:.5f} to {cur:.5f}, saving model to {path}")
            self.best = cur
        Path(path).parent.mkdir(parents=True, exist_ok=True)
        self.model.save(path)


class BackupAndRestore(Callback):
    """Epoch-granular resume of an interrupted ``fit`` (weights, optimizer moments, epoch index)."""

    def __init__(self, backup_dir, **kwargs):
        self.dir = Path(backup_dir)

    def initial_epoch(self, requested):
        meta = self.dir / "state.json"
        if not meta.exists():
            return requested
        import torch
        st = json.loads(meta.read_text())
        self.model.load_weights(self.dir / "weights.keras")
        opt = self.model.optimizer
        opt.ensure_state(self.model)
        blob = torch.load(self.dir / "optimizer.pt", map_location=self.model.P.device)
        opt.restore((blob["m"], blob["v"], blob["step"]))
        opt.iterations = int(st["iterations"])
        return max(requested, int(st["epoch"]) + 1)

    def on_epoch_end(self, epoch, logs=None):
        import torch
        self.dir.mkdir(parents=True, exist_ok=True)
        self.model.save(self.dir / "weights.keras")
        opt = self.model.optimizer
        m, v, step = opt.snapshot()
        torch.save({"m": m, "v": v, "step": step}, self.dir / "optimizer.pt")
        (self.dir / "state.json").write_text(json.dumps({"epoch": epoch, "iterations": opt.iterations}))

    def on_train_end(self, logs=None):
        for f in ("state.json", "optimizer.pt", "weights.keras"):
            p = self.dir / f
            if p.exists():
                p.unlink()


class TensorBoard(Callback):
    """Epoch scalars under ``log_dir/train`` and ``log_dir/validation`` (profile_batch ignored)."""

    def __init__(self, log_dir="logs", update_freq="epoch", **kwargs):
        self.log_dir = str(log_dir)
        self._writers = None

    def on_train_begin(self, logs=None):
        try:
            from torch.utils.tensorboard import SummaryWriter
            self._writers = {"train": SummaryWriter(os.path.join(self.log_dir, "train")),
                             "validation": SummaryWriter(os.path.join(self.log_dir, "validation"))}
        except Exception:  # tensorboard not importable: scalars are still in History / stdout
            self._writers = None

    def on_epoch_end(self, epoch, logs=None):
        if not self._writers:
            return
        for k, v in (logs or {}).items():
            if k.startswith("val_"):
                self._writers["validation"].add_scalar("epoch_" + k[4:], v, epoch)
            else:
                self._writers["train"].add_scalar("epoch_" + k, v, epoch)

    def on_train_end(self, logs=None):
        if self._writers:
            for w in self._writers.values():
                w.flush(); w.close()


class ReduceLROnPlateau(Callback):
    def __init__(self, monitor="val_loss", factor=0.1, patience=10, verbose=0, mode="auto", min_delta=1e-4, cooldown=0,
                 min_lr=0.0, **kwargs):
        self.monitor, self.factor, self.patience, self.verbose = monitor, factor, patience, verbose
        self.min_delta, self.cooldown, self.min_lr = min_delta, cooldown, min_lr
        self._cmp, self._worst = _better(mode, monitor)

    def on_train_begin(self, logs=None):
        self.best, self.wait, self.cool = self._worst, 0, 0

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if cur is None:
            return
        if self.cool > 0:
            self.cool -= 1
            self.wait = 0
        if self._cmp(cur, self.best, self.min_delta):
            self.best, self.wait = cur, 0
        elif self.cool <= 0:
            self.wait += 1
            if self.wait >= self.patience:
                opt = self.model.optimizer
                old = opt.current_lr()
                if old > self.min_lr and not callable(opt.learning_rate):
                    new = max(old * self.factor, self.min_lr)
                    opt.set_learning_rate(new)
                    if self.verbose:
                        print(f"\nEpoch {epoch + 1}: ReduceLROnPlateau reducing learning rate to {new}.")
                self.cool, self.wait = self.cooldown, 0


class CallbackList:
    def __init__(self, callbacks, model):
        self.callbacks = list(callbacks)
        for cb in self.callbacks:
            cb.set_model(model)

    def initial_epoch(self, requested):
        for cb in self.callbacks:
            requested = cb.initial_epoch(requested)
        return requested

    def on_train_begin(self):
        for cb in self.callbacks: cb.on_train_begin()

    def on_train_end(self):
        for cb in self.callbacks: cb.on_train_end()

    def on_epoch_begin(self, epoch):
        for cb in self.callbacks: cb.on_epoch_begin(epoch)

    def on_epoch_end(self, epoch, logs):
        for cb in self.callbacks: cb.on_epoch_end(epoch, logs)
