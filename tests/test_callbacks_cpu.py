"""Epoch-level behaviour of the callbacks the reference's trainers pass to fit (keras 3.3.3 semantics), on a stand-in model:
no GPU needed, these are host-side state machines."""
import json

import numpy as np
import pytest

from b200unet.keras.callbacks import (BackupAndRestore, CallbackList, EarlyStopping, ModelCheckpoint, ReduceLROnPlateau)


class _Opt:
    def __init__(self, lr=1e-3):
        self.learning_rate, self.iterations = lr, 0

    def current_lr(self):
        return float(self.learning_rate)

    def set_learning_rate(self, lr):
        self.learning_rate = float(lr)


class _Model:
    """Weights = [the index of the epoch that produced them]; save() records (path, weights)."""

    def __init__(self):
        self.w, self.stop_training, self.saved, self.optimizer = [np.array([-1.0])], False, [], _Opt()

    def get_weights(self):
        return [w.copy() for w in self.w]

    def set_weights(self, ws):
        self.w = [np.asarray(w).copy() for w in ws]

    def save(self, path):
        self.saved.append((str(path), float(self.w[0][0])))


def _run(callbacks, series, key="val_loss"):
    model = _Model()
    cbs = CallbackList(callbacks, model)
    cbs.on_train_begin()
    ran = 0
    for epoch, v in enumerate(series):
        model.w = [np.array([float(epoch)])]
        cbs.on_epoch_begin(epoch)
        cbs.on_epoch_end(epoch, {key: v, "loss": v})
        ran = epoch + 1
        if model.stop_training:
            break
    cbs.on_train_end()
    return model, ran


def test_early_stopping_patience_and_restore():
    series = [1.0, 0.8, 0.9, 0.85, 0.81, 0.7, 0.6]
    m, ran = _run([EarlyStopping(monitor="val_loss", patience=3, restore_best_weights=True)], series)
    assert ran == 5 and m.stop_training and m.w[0][0] == 1.0                 # three epochs without beating 0.8 -> stop, best = epoch 1
    m, ran = _run([EarlyStopping(monitor="val_loss", patience=10, restore_best_weights=True)], series)
    assert ran == len(series) and not m.stop_training and m.w[0][0] == 6.0   # ran out of epochs: best epoch's weights (the last)
    m, ran = _run([EarlyStopping(monitor="val_loss", patience=10, restore_best_weights=True)], [0.5, 0.6, 0.7])
    assert ran == 3 and m.w[0][0] == 0.0                                      # ... restored at the end even without an early stop
    m, ran = _run([EarlyStopping(monitor="val_loss", patience=10)], [0.5, 0.6, 0.7])
    assert m.w[0][0] == 2.0                                                   # restore_best_weights=False: last weights stay
    m, ran = _run([EarlyStopping(monitor="val_loss", patience=0)], [0.5, 0.6, 0.7])
    assert ran == 2                                                           # never after the very first epoch
    m, ran = _run([EarlyStopping(monitor="val_loss", patience=2, min_delta=0.05)], [1.0, 0.97, 0.96, 0.5])
    assert ran == 3                                                           # improvements below min_delta do not count
    m, ran = _run([EarlyStopping(monitor="val_dice", mode="max", patience=2, restore_best_weights=True)],
                  [0.1, 0.5, 0.4, 0.45, 0.9], key="val_dice")
    assert ran == 4 and m.w[0][0] == 1.0
    m, ran = _run([EarlyStopping(monitor="val_psnr", patience=1)], [30.0, 31.0, 30.5, 33.0], key="val_psnr")
    assert ran == 3                                                           # mode "auto": psnr is maximised
    m, ran = _run([EarlyStopping(monitor="missing", patience=1)], [1.0, 2.0, 3.0])
    assert ran == 3                                                           # monitored key absent: no decision


def test_model_checkpoint_best_only_and_templates(tmp_path, capsys):
    path = tmp_path / "ck" / "best.keras"
    m, _ = _run([ModelCheckpoint(filepath=str(path), monitor="val_loss", save_best_only=True, verbose=1)], [1.0, 0.8, 0.9, 0.7])
    assert [w for _, w in m.saved] == [0.0, 1.0, 3.0] and all(p == str(path) for p, _ in m.saved) and path.parent.is_dir()
    out = capsys.readouterr().out
    assert "Epoch 1: val_loss improved from inf to 1.00000, saving model to" in out
    assert "Epoch 3: val_loss did not improve from 0.80000" in out
    m, _ = _run([ModelCheckpoint(filepath=str(tmp_path / "e{epoch:02d}-{val_loss:.2f}.keras"))], [1.0, 0.5])
    assert [p.rsplit("/", 1)[1] for p, _ in m.saved] == ["e01-1.00.keras", "e02-0.50.keras"]     # every epoch, keras templates
    m, _ = _run([ModelCheckpoint(filepath=str(path), monitor="val_dice_coefficient", mode="max", save_best_only=True)],
                [0.2, 0.1, 0.3], key="val_dice_coefficient")
    assert [w for _, w in m.saved] == [0.0, 2.0]


def test_reduce_lr_on_plateau():
    cb = ReduceLROnPlateau(monitor="val_loss", factor=0.5, patience=2, min_lr=2e-4, cooldown=1)
    series = [1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0]
    model = _Model()
    cbs = CallbackList([cb], model)
    cbs.on_train_begin()
    lrs = []
    for epoch, v in enumerate(series):
        cbs.on_epoch_end(epoch, {"val_loss": v})
        lrs.append(model.optimizer.current_lr())
    # epoch 0 sets the best; epochs 1-2 wait -> halve; the cooldown epoch ends in the same callback and already counts as a
    # wait (keras: the counter is decremented before the plateau test), so the next halving comes two epochs later; the rate
    # never goes below min_lr
    assert lrs == [1e-3, 1e-3, 5e-4, 5e-4, 2.5e-4, 2.5e-4, 2e-4, 2e-4, 2e-4]
    cb = ReduceLROnPlateau(monitor="val_loss", factor=0.5, patience=1, min_delta=1e-4)
    model = _Model()
    cbs = CallbackList([cb], model)
    cbs.on_train_begin()
    for epoch, v in enumerate([1.0, 0.99995, 0.9]):       # an improvement below min_delta counts as a plateau
        cbs.on_epoch_end(epoch, {"val_loss": v})
    assert model.optimizer.current_lr() == 5e-4


def test_backup_and_restore_state_files(tmp_path, monkeypatch):
    """Epoch-granular resume: state.json + weights + optimizer moments after every epoch, removed when training completes;
    a new run that finds them resumes after the recorded epoch."""
    import torch

    class M(_Model):
        P = torch.zeros(1)

        def __init__(self):
            super().__init__()
            self.loaded = None
            self.optimizer.snapshot = lambda: (torch.ones(2), torch.ones(2) * 2, torch.tensor([7]))
            self.optimizer.restore = lambda snap: setattr(self, "restored", snap)
            self.optimizer.ensure_state = lambda model: None

        def save(self, path):
            super().save(path)
            open(path, "w").write("w")

        def load_weights(self, path):
            self.loaded = str(path)

    model = M()
    cb = BackupAndRestore(str(tmp_path / "bk"))
    cb.set_model(model)
    assert cb.initial_epoch(0) == 0
    model.optimizer.iterations = 30
    cb.on_epoch_end(2, {})
    st = json.loads((tmp_path / "bk" / "state.json").read_text())
    assert st == {"epoch": 2, "iterations": 30} and (tmp_path / "bk" / "optimizer.pt").exists()
    fresh = M()
    cb2 = BackupAndRestore(str(tmp_path / "bk"))
    cb2.set_model(fresh)
    assert cb2.initial_epoch(0) == 3 and fresh.loaded.endswith("weights.keras") and fresh.optimizer.iterations == 30
    assert float(fresh.restored[1][0]) == 2.0
    cb2.on_train_end()
    assert not list((tmp_path / "bk").iterdir())


def test_fit_loop_semantics_on_cpu(monkeypatch, capsys):
    """Model.fit's epoch loop with the GPU step replaced: steps per epoch (an infinite stream keeps its position across epochs,
    a finite one restarts), epoch logs = mean over the steps, validation every validation_freq epochs under val_ names,
    learning_rate in the logs, History layout, initial_epoch, EarlyStopping ending the loop, the verbose=2 lines."""
    import itertools
    import torch
    from b200unet import builders as B
    from b200unet.keras import clear_session
    from b200unet.keras.optimizers import Adam
    clear_session()
    model, _ = B.build_super_resolution_unet(0.5, depth_override=1, input_size=16)
    model.compile(optimizer=Adam(learning_rate=2e-4), loss="charbonnier")
    seen = []

    def fake_step(x, y, return_tensors=False):
        seen.append(int(x[0]))
        return {"loss": torch.tensor(float(x[0])), "psnr": torch.tensor(30.0 + len(seen))}

    monkeypatch.setattr(model, "train_on_batch", fake_step)
    val_calls = []
    monkeypatch.setattr(model, "evaluate", lambda ds, steps=None, return_dict=False, verbose=0:
                        (val_calls.append(steps) or {"loss": 0.5 / len(val_calls), "psnr": 20.0 + len(val_calls)}))
    batches = lambda: ((np.full((2,), i), np.zeros(2)) for i in itertools.count())
    hist = model.fit(batches(), epochs=3, steps_per_epoch=4, validation_data=[1], validation_steps=7, validation_freq=2, verbose=2)
    assert seen == list(range(12))                                             # the stream is not rewound between epochs
    assert hist.epoch == [0, 1, 2]
    assert hist.history["loss"] == [1.5, 5.5, 9.5]                             # mean over the epoch's steps
    assert hist.history["val_loss"] == [0.5] and hist.history["val_psnr"] == [21.0] and val_calls == [7]   # only epoch 2
    assert hist.history["learning_rate"] == [2e-4] * 3
    out = capsys.readouterr().out
    assert "Epoch 2/3" in out and " - loss: 5.5000 - psnr: " in out and "val_loss: 0.5000" in out and "learning_rate: 2.0000e-04" in out
    # a finite dataset without steps_per_epoch: one pass per epoch, restarted every epoch
    seen.clear()
    finite = [(np.full((2,), i), np.zeros(2)) for i in range(3)]
    hist = model.fit(finite, epochs=2, verbose=0)
    assert seen == [0, 1, 2, 0, 1, 2] and hist.history["loss"] == [1.0, 1.0]
    # ... and with steps_per_epoch larger than the dataset it wraps around
    seen.clear()
    model.fit(finite, epochs=1, steps_per_epoch=5, verbose=0)
    assert seen == [0, 1, 2, 0, 1]
    # initial_epoch skips epochs; EarlyStopping ends the loop
    seen.clear()
    hist = model.fit(batches(), epochs=5, initial_epoch=3, steps_per_epoch=1, verbose=0)
    assert hist.epoch == [3, 4] and len(seen) == 2
    seen.clear()
    val_calls.clear()
    monkeypatch.setattr(model, "evaluate", lambda ds, steps=None, return_dict=False, verbose=0:
                        (val_calls.append(steps) or {"loss": 1.0 + len(val_calls)}))
    monkeypatch.setattr(model, "get_weights", lambda: [np.zeros(1)])
    monkeypatch.setattr(model, "set_weights", lambda w: None)
    hist = model.fit(batches(), epochs=10, steps_per_epoch=1, validation_data=[1], verbose=0,
                     callbacks=[EarlyStopping(monitor="val_loss", patience=2, restore_best_weights=True)])
    assert hist.epoch == [0, 1, 2] and model.stop_training
    clear_session()


def test_cosine_decay_and_adam_host_bookkeeping():
    """CosineDecay as keras defines it (Segmenation/code/train_adaptive_unet.py:451-460), and Adam's host side: the
    schedule is evaluated at the 0-based count of updates already made, the device-resident hyper-parameters are pushed when
    the rate changes, (1 - beta) is formed in double precision before rounding to float32."""
    import math
    import types
    import torch
    from b200unet.keras.optimizers import Adam, CosineDecay
    sched = CosineDecay(initial_learning_rate=1e-3, decay_steps=100, alpha=0.0)
    for step in (0, 1, 25, 50, 99, 100, 250):
        s = min(step, 100)
        assert sched(step) == pytest.approx(1e-3 * 0.5 * (1.0 + math.cos(math.pi * s / 100)), rel=1e-12, abs=1e-18)
    assert sched(0) == 1e-3 and sched(100) == pytest.approx(0.0, abs=1e-18) and sched(10 ** 6) == sched(100)
    assert CosineDecay(1e-3, 10, alpha=0.1)(10) == pytest.approx(1e-4)
    opt = Adam(learning_rate=sched)
    fake = types.SimpleNamespace(P=torch.zeros(4))
    opt.ensure_state(fake)
    h = opt._state["hyper"]
    assert h.dtype == torch.float32 and h.tolist()[:4] == [np.float32(1e-3), np.float32(0.9), np.float32(0.999), np.float32(1e-7)]
    assert h[4].item() == np.float32(1 - 0.9) and h[5].item() == np.float32(1 - 0.999)     # double precision, then rounded
    assert h[5].item() != np.float32(np.float32(1) - np.float32(0.999))                      # ... not float32 arithmetic
    lrs = []
    for _ in range(3):
        opt.before_step()
        lrs.append(opt._state["hyper"][0].item())
    assert lrs == [np.float32(sched(0)), np.float32(sched(1)), np.float32(sched(2))] and opt.iterations == 3
    const = Adam(learning_rate=3e-4)
    const.ensure_state(fake)
    const.before_step()
    assert const.current_lr() == 3e-4 and const._state["hyper"][0].item() == np.float32(3e-4)
    const.set_learning_rate(1.5e-4)                       # what ReduceLROnPlateau calls
    assert const._state["hyper"][0].item() == np.float32(1.5e-4)
