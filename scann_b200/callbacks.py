"""Training-shell callbacks of the reference, restated for the ``fit`` loop of ``scann_b200.model``.

The reference drives training with Keras callbacks (scann/models/scann_model.py:163-197):
``ModelCheckpoint(monitor="val_mae", save_best_only=True)``, ``EarlyStopping(monitor="val_mae", patience=200)``
and, for ``scheduler: sgdr``, its own ``SGDRC`` warm-restart schedule (scann/layers/custom_layers.py:78-179)
wrapped in a ``LearningRateScheduler``; otherwise a ``LearningRateLoggingCallback``.  Keras is not installed, so
the protocol (``set_model`` / ``on_train_begin`` / ``on_epoch_begin`` / ``on_epoch_end`` and
``model.stop_training``) and the behaviour of each callback are implemented here; the learning rate itself is
consumed by the fused Adam kernel (legacy ``decay=1e-5`` applied on top, csrc/optim.cu).
"""
from __future__ import annotations

import math
import os

import numpy as np
from typing import Callable, Dict, Optional


class Callback:
    model = None

    def set_model(self, model) -> None:
        self.model = model

    def on_train_begin(self, logs: Optional[Dict] = None) -> None:
        pass

    def on_epoch_begin(self, epoch: int, logs: Optional[Dict] = None) -> None:
        pass

    def on_epoch_end(self, epoch: int, logs: Optional[Dict] = None) -> None:
        pass

    def on_train_end(self, logs: Optional[Dict] = None) -> None:
        pass


class LearningRateScheduler(Callback):
    """keras.callbacks.LearningRateScheduler: ``lr = schedule(epoch)`` at the start of every epoch."""

    def __init__(self, schedule: Callable[[int], float], verbose: int = 0):
        self.schedule, self.verbose = schedule, verbose

    def on_epoch_begin(self, epoch, logs=None):
        lr = float(self.schedule(epoch))
        self.model.lr = lr
        if self.verbose:
            print(f"Epoch {epoch + 1}: LearningRateScheduler setting learning rate to {lr}.")

    def on_epoch_end(self, epoch, logs=None):
        if logs is not None:
            logs["lr"] = self.model._lr_now()


class LearningRateLoggingCallback(Callback):
    """Prints the optimiser's current learning rate after every epoch (custom_layers.py:68-75)."""

    def on_epoch_end(self, epoch, logs=None):
        print(f"Current learning rate: {self.model._lr_now():.8f}")


class ModelCheckpoint(Callback):
    """keras.callbacks.ModelCheckpoint for the options the reference uses: monitor (min mode for *mae / *loss),
    save_best_only, save_weights_only; writes Keras legacy HDF5 (``filepath`` ends in .h5) through h5lite."""

    def __init__(self, filepath: str, monitor: str = "val_loss", save_best_only: bool = False,
                 save_weights_only: bool = False, verbose: int = 0, mode: str = "auto"):
        self.filepath, self.monitor = filepath, monitor
        self.save_best_only, self.save_weights_only, self.verbose = save_best_only, save_weights_only, verbose
        self.maximize = mode == "max" or (mode == "auto" and ("acc" in monitor or monitor.startswith("fmeasure")))
        self.best = -math.inf if self.maximize else math.inf

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        path = self.filepath.format(epoch=epoch + 1, **logs)
        if self.save_best_only:
            cur = logs.get(self.monitor)
            if cur is None:
                print(f"Can save best model only with {self.monitor} available, skipping.")
                return
            better = cur > self.best if self.maximize else cur < self.best
            if not better:
                if self.verbose:
                    print(f"Epoch {epoch + 1}: {self.monitor} did not improve from {self.best:.5f}")
                return
            if self.verbose:
                print(f"Epoch {epoch + 1}: {self.monitor} improved from {self.best:.5f} to {cur:.5f}, saving model to {path}")
            self.best = cur
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        (self.model.save_weights if self.save_weights_only else self.model.save)(path)


class EarlyStopping(Callback):
    """keras.callbacks.EarlyStopping(monitor, patience), min mode for *mae / *loss."""

    def __init__(self, monitor: str = "val_loss", patience: int = 0, min_delta: float = 0.0, mode: str = "auto"):
        self.monitor, self.patience, self.min_delta = monitor, patience, abs(min_delta)
        self.maximize = mode == "max" or (mode == "auto" and "acc" in monitor)

    def on_train_begin(self, logs=None):
        self.wait, self.stopped_epoch = 0, 0
        self.best = -math.inf if self.maximize else math.inf

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if cur is None:
            return
        improved = cur - self.min_delta > self.best if self.maximize else cur + self.min_delta < self.best
        if improved:
            self.best, self.wait = cur, 0
            return
        self.wait += 1
        if self.wait >= self.patience:
            self.stopped_epoch = epoch
            self.model.stop_training = True


class SGDRC(Callback):
    """SGDR warm restarts with a validation trigger and a shrinking restart peak -- the schedule of the reference's
    ``SGDRC`` (custom_layers.py:78-179), used as ``LearningRateScheduler(sgdrc.lr_scheduler)`` next to the
    callback itself (scann_model.py:182-193).

    State machine.  Until ``val_mae <= trigger_val_mae`` has been seen once the rate stays at ``lr_max``.  From
    then on every ``lr_scheduler`` call (one per epoch) advances the position ``tcur`` inside a cycle of ``ti``
    epochs, ``lr = lr_min + (peak - lr_min) * (1 + cos(pi * tcur / ti)) / 2``; when a cycle ends its length is
    multiplied by ``tmult`` and the peak becomes ``next_peak``.  Whenever the validation MAE improves (after the
    trigger) ``next_peak = max(peak / lr_max_compression, lr)`` (or the current rate if compression <= 0).
    """

    def __init__(self, lr_max: float, lr_min: float, lr_max_compression: float = 5, t0: int = 10, tmult: float = 1,
                 trigger_val_mae: float = 9999, show_lr: bool = True):
        self.lr_max, self.lr_min = lr_max, lr_min
        self.lr_max_compression, self.t0, self.tmult = lr_max_compression, t0, tmult
        self.trigger_val_mae, self.show_lr = trigger_val_mae, show_lr
        self._reset()

    def _reset(self) -> None:
        self.triggered = False
        self.lr = self.lr_warmup_current = self.lr_warmup_next = self.lr_max
        self.ti, self.tcur = self.t0, 1
        self.best_val_mae = 9999

    def on_train_begin(self, logs=None):
        self._reset()

    def on_epoch_end(self, epoch, logs=None):
        val = (logs or {}).get("val_mae")
        if val is not None:
            self.triggered = self.triggered or val <= self.trigger_val_mae
            if self.triggered and val < self.best_val_mae:
                self.best_val_mae = val
                self.lr_warmup_next = (max(self.lr_warmup_current / self.lr_max_compression, self.lr)
                                       if self.lr_max_compression > 0 else self.lr)
        if self.show_lr:
            print(f"sgdr_triggered = {self.triggered}, current_lr = {self.lr:f}, "
                  f"next_warmup_lr = {self.lr_warmup_next:f}, next_warmup = {self.ti - self.tcur}")

    def lr_scheduler(self, epoch: int) -> float:
        if self.triggered:
            self.tcur += 1
            if self.tcur > self.ti:                       # cycle finished: longer cycle, new peak
                self.ti, self.tcur = int(self.tmult * self.ti), 1
                self.lr_warmup_current = self.lr_warmup_next
            # the reference's expression, operation for operation (custom_layers.py:176-178): the schedule is pinned
            # bit for bit against the reference's own class (tests/test_reference_pins.py)
            self.lr = float(self.lr_min + (self.lr_warmup_current - self.lr_min) *
                            (1 + np.cos(self.tcur / self.ti * np.pi)) / 2.0)
        return self.lr
