"""Adam with keras semantics plus the learning-rate schedules the reference uses.

/root/reference: Super_resolution/code/train_adaptive_unet.py:490 (Adam(learning_rate)),
Segmenation/code/train_adaptive_unet.py:451-460 (CosineDecay).  beta_1 0.9, beta_2 0.999,
epsilon 1e-7, bias-corrected step size; one multi-tensor kernel updates the flat fp32
parameter buffer and writes the bf16 shadow copy the convolutions read.
"""
from __future__ import annotations

import math

import torch

from .. import ops


class CosineDecay:
    """keras.optimizers.schedules.CosineDecay(initial_learning_rate, decay_steps, alpha)."""

    def __init__(self, initial_learning_rate, decay_steps, alpha=0.0, **kwargs):
        self.initial_learning_rate, self.decay_steps, self.alpha = float(initial_learning_rate), int(decay_steps), float(alpha)

    def __call__(self, step):
        frac = min(max(step, 0), self.decay_steps) / max(self.decay_steps, 1)
        cos = 0.5 * (1.0 + math.cos(math.pi * frac))
        return self.initial_learning_rate * ((1.0 - self.alpha) * cos + self.alpha)


class Adam:
    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, **kwargs):
        self.learning_rate = learning_rate
        self.beta_1, self.beta_2, self.epsilon = float(beta_1), float(beta_2), float(epsilon)
        self.iterations = 0
        self._state = None
        # keras wraps the optimizer in a LossScaleOptimizer under the `mixed_float16` policy (dynamic scaling: initial
        # scale 2**15, doubled after 2000 finite steps in a row, halved -- and the step skipped -- on inf / nan):
        # Model.compile switches this on for that policy; the state lives on the device (ops.loss_scale_*)
        self.dynamic_loss_scale = False
        self.initial_scale, self.dynamic_growth_steps = 2.0 ** 15, 2000

    def loss_scale_state(self):
        """Device fp32[4] {scale, finite steps in a row, found_inf, skipped steps} or None."""
        return self._state.get("loss_scale") if (self._state is not None and self.dynamic_loss_scale) else None

    @property
    def loss_scale(self) -> float:
        ls = self.loss_scale_state()
        return float(ls[0]) if ls is not None else 1.0

    def current_lr(self) -> float:
        lr = self.learning_rate
        return float(lr(self.iterations)) if callable(lr) else float(lr)

    def ensure_state(self, model):
        if self._state is None:
            dev = model.P.device
            self._state = {
                "m": torch.zeros_like(model.P), "v": torch.zeros_like(model.P),
                "step": torch.zeros(1, dtype=torch.int32, device=dev),
                "hyper": torch.zeros(6, dtype=torch.float32, device=dev),
            }
            if self.dynamic_loss_scale:
                self._state["loss_scale"] = torch.tensor([self.initial_scale, 0.0, 0.0, 0.0], dtype=torch.float32, device=dev)
            self._push_hyper()

    def _push_hyper(self):
        # (1 - beta) is evaluated in Python double precision and then rounded, as keras does
        h = torch.tensor([self.current_lr(), self.beta_1, self.beta_2, self.epsilon, 1 - self.beta_1, 1 - self.beta_2],
                         dtype=torch.float32)
        self._state["hyper"].copy_(h, non_blocking=True)
        self._pushed_lr = float(h[0])

    def set_learning_rate(self, lr: float):
        self.learning_rate = float(lr)
        if self._state is not None:
            self._push_hyper()

    def before_step(self):
        """Host-side bookkeeping per step (schedules change the device-resident lr outside the graph)."""
        if self._state is not None and callable(self.learning_rate):
            lr = self.current_lr()
            if lr != self._pushed_lr:
                self._push_hyper()
        self.iterations += 1

    def advance(self):
        """Increment the device-resident step counter (once per training step, before any apply_ranges)."""
        ops.adam_advance(self._state["step"], self.loss_scale_state())

    def apply_ranges(self, model, ranges):
        """Adam over [(lo, hi), ...] of the flat buffers WITHOUT advancing the step counter (see advance)."""
        st = self._state
        shadow = model.S if model.S is not model.P else None
        for lo, hi in ranges:
            if hi > lo:
                ops.adam_step(model.P[lo:hi], model.G[lo:hi], st["m"][lo:hi], st["v"][lo:hi], st["hyper"], st["step"],
                              None if shadow is None else shadow[lo:hi], self.loss_scale_state())

    def apply(self, model, ranges=None):
        """One Adam step over the flat buffers, or over `ranges` = [(lo, hi), ...] of them (sharded optimizer)."""
        st = self._state
        ls = self.loss_scale_state()      # (Model._check_finite has set its found_inf flag: such a step changes nothing)
        ops.adam_advance(st["step"], ls)
        shadow = model.S if model.S is not model.P else None
        if ranges is None:
            ops.adam_step(model.P, model.G, st["m"], st["v"], st["hyper"], st["step"], shadow, ls)
        else:
            for lo, hi in ranges:
                if hi > lo:
                    ops.adam_step(model.P[lo:hi], model.G[lo:hi], st["m"][lo:hi], st["v"][lo:hi], st["hyper"], st["step"],
                                  None if shadow is None else shadow[lo:hi], ls)
        if ls is not None:
            ops.loss_scale_update(ls, self.dynamic_growth_steps)

    def snapshot(self):
        st = self._state
        ls = st.get("loss_scale")
        return (st["m"].clone(), st["v"].clone(), st["step"].clone()) + ((ls.clone(),) if ls is not None else ())

    def restore(self, snap):
        st = self._state
        st["m"].copy_(snap[0]); st["v"].copy_(snap[1]); st["step"].copy_(snap[2])
        if len(snap) > 3 and st.get("loss_scale") is not None:
            st["loss_scale"].copy_(snap[3])
