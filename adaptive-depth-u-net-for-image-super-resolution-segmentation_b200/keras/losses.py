"""Loss / metric objects of the reference, executed by the fused loss kernels.

  * SR: charbonnier / l1 / mse + the psnr metric
    (/root/reference/Super_resolution/code/train_adaptive_unet.py:294-348)
  * segmentation: BinaryCrossentropy, Dice, IoU and the BCE+Dice hybrids
    (/root/reference/Segmenation/code/train_adaptive_unet.py:258-318)
  * CategoricalCrossentropy on a softmax head (config C4; not in the reference --
    keras 3.3.3 semantics)

Each object is (a) a keras-style callable ``loss(y_true, y_pred) -> scalar tensor`` for
eager use on CUDA tensors and (b) a plan participant: ``launch`` runs the fused
forward+backward kernel that writes d(loss)/d(pred) into the output's gradient buffer.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from .._ffi import LOSS_CHARBONNIER, LOSS_L1, LOSS_MSE


class _PlanLoss:
    name = "loss"
    metric_names = ()
    target_dtype = torch.float32

    def make_state(self, plan):
        ov = plan.output_val
        st = {"out": torch.zeros(8, dtype=torch.float32, device=plan.dev),
              "target": torch.zeros((plan.batch,) + self._target_shape(ov), dtype=self.target_dtype, device=plan.dev)}
        self._more_state(plan, st)
        return st

    def _target_shape(self, ov):
        return (ov.h, ov.w, ov.c)

    def _more_state(self, plan, st):
        pass

    def set_target(self, st, y):
        if isinstance(y, np.ndarray):
            y = torch.from_numpy(y)
        tgt = st["target"]
        if y.dtype != tgt.dtype and not y.is_cuda:
            y = y.to(tgt.dtype)
        tgt.copy_(y.reshape(tgt.shape), non_blocking=True)

    def grad_tensor(self, plan, st):
        """The contiguous tensor `launch` writes d(loss)/d(prediction) into."""
        return plan.output_val.grad

    def logs(self, st):
        out = st["out"]
        d = {"loss": out[0]}
        for i, n in enumerate(self.metric_names):
            d[n] = out[self.metric_slots[i]]
        return d


class SRLoss(_PlanLoss):
    """charbonnier (eps 1e-3) / l1 / mse with the psnr metric."""
    metric_names = ("psnr",)
    metric_slots = (1,)

    def __init__(self, kind: str, eps: float = 1e-3):
        self.kind = {"charbonnier": LOSS_CHARBONNIER, "l1": LOSS_L1, "mse": LOSS_MSE}[kind]
        self.eps = float(eps)
        self.__name__ = {"charbonnier": "charbonnier_loss", "l1": "l1_loss", "mse": "mse_loss"}[kind]

    def _more_state(self, plan, st):
        st["ws"] = torch.zeros(2 + plan.batch, dtype=torch.float32, device=plan.dev)

    def launch(self, plan, st, grad_scale=1.0, with_grad=True):
        ov = plan.output_val
        ops.sr_loss(ov.buf, st["target"], self.kind, self.eps, grad_scale, st["out"], ov.grad if with_grad else None,
                    st["ws"])

    def __call__(self, y_true, y_pred):
        out = torch.zeros(2, device=y_pred.device)
        ws = torch.zeros(2 + y_pred.shape[0], device=y_pred.device)
        ops.sr_loss(y_pred.contiguous(), y_true.to(y_pred.device).float().contiguous(), self.kind, self.eps, 1.0, out,
                    None, ws)
        return out[0]


class PSNRMetric:
    __name__ = "psnr"

    def __call__(self, y_true, y_pred):
        out = torch.zeros(2, device=y_pred.device)
        ws = torch.zeros(2 + y_pred.shape[0], device=y_pred.device)
        ops.sr_loss(y_pred.contiguous(), y_true.to(y_pred.device).float().contiguous(), LOSS_MSE, 0.0, 1.0, out, None, ws)
        return out[1]


class BceDiceLoss(_PlanLoss):
    """bce_weight * BinaryCrossentropy + dice_weight * (1 - dice); logs dice and iou."""
    metric_names = ("dice", "iou")
    metric_slots = (2, 3)

    def __init__(self, bce_weight=1.0, dice_weight=0.0, name="bce_dice", global_dice=False, keras_metrics=False):
        self.bw, self.dw = float(bce_weight), float(dice_weight)
        self.__name__ = name
        # the baseline trainer's extra keras metrics (unet_vinillia.py:266-270): BinaryAccuracy(name="accuracy"),
        # Precision, Recall at threshold 0.5 -- one more pass over the (1-channel) prediction per step.  Precision and
        # recall are STATEFUL in keras (tp / fp / fn accumulate over the epoch): next to the batch values the step logs
        # the per-sample counter rates under "_tp" / "_fp" / "_fn", which Model combines at the end of an epoch.
        self.keras_metrics = bool(keras_metrics)
        if global_dice:      # the baseline trainer's metric: one Dice ratio over the whole batch (unet_vinillia.py:94-99),
            # logged under the function's name as keras does ("dice_coefficient" / "val_dice_coefficient", :270-273)
            self.metric_names, self.metric_slots = ("dice_coefficient", "iou"), (4, 3)
        self.extra_metric_names = ("accuracy", "precision", "recall") if self.keras_metrics else ()

    def _more_state(self, plan, st):
        st["ws"] = torch.zeros(1 + 3 * plan.batch, dtype=torch.float32, device=plan.dev)
        if self.keras_metrics:
            st["counts"] = torch.zeros(4, dtype=torch.float32, device=plan.dev)

    def launch(self, plan, st, grad_scale=1.0, with_grad=True):
        ov = plan.output_val
        ops.bce_dice_loss(ov.buf, st["target"], self.bw, self.dw, grad_scale, st["out"],
                          ov.grad if with_grad else None, st["ws"])
        if self.keras_metrics:
            ops.binary_confusion(ov.buf, st["target"], st["counts"])

    def logs(self, st):
        if not self.keras_metrics:
            return super().logs(st)
        d = {"loss": st["out"][0]}                    # keras' order: loss, then the metrics as listed at compile()
        tp, fp, fn, ok = st["counts"].unbind(0)
        n = float(st["target"].shape[0])
        d["accuracy"] = ok / float(st["target"].numel())
        d["precision"] = torch.where(tp + fp > 0, tp / (tp + fp).clamp_min(1.0), torch.zeros_like(tp))
        d["recall"] = torch.where(tp + fn > 0, tp / (tp + fn).clamp_min(1.0), torch.zeros_like(tp))
        for m, sl in zip(self.metric_names, self.metric_slots):
            d[m] = st["out"][sl]
        d["_tp"], d["_fp"], d["_fn"] = tp / n, fp / n, fn / n
        return d

    def __call__(self, y_true, y_pred):
        out = torch.zeros(8, device=y_pred.device)
        ws = torch.zeros(1 + 3 * y_pred.shape[0], device=y_pred.device)
        ops.bce_dice_loss(y_pred.contiguous(), y_true.to(y_pred.device).float().contiguous(), self.bw, self.dw, 1.0, out,
                          None, ws)
        return out[0]


def BinaryCrossentropy(global_dice=False, keras_metrics=False):
    return BceDiceLoss(1.0, 0.0, name="binary_crossentropy", global_dice=global_dice, keras_metrics=keras_metrics)


class CategoricalCrossentropy(_PlanLoss):
    """CategoricalCrossentropy(from_logits=False) on a softmax head; targets are int32 class ids
    (sparse) or one-hot maps (converted on set_target)."""
    target_dtype = torch.int32
    __name__ = "categorical_crossentropy"

    def _target_shape(self, ov):
        return (ov.h, ov.w)

    def _more_state(self, plan, st):
        st["ws"] = torch.zeros(1, dtype=torch.float32, device=plan.dev)
        sm = [op for op in plan.ops if op.kind == "softmax" and op.output is plan.output_val]
        if not sm:
            raise NotImplementedError("CategoricalCrossentropy needs a Conv2D(activation='softmax') output layer")
        st["logits"] = sm[0].inputs[0]

    def set_target(self, st, y):
        if isinstance(y, np.ndarray):
            y = torch.from_numpy(y)
        if y.dim() == 4:
            y = y.argmax(dim=-1)
        st["target"].copy_(y.to(torch.int32).reshape(st["target"].shape), non_blocking=True)

    def grad_tensor(self, plan, st):
        return st["logits"].grad

    def launch(self, plan, st, grad_scale=1.0, with_grad=True):
        ops.softmax_ce_loss(plan.output_val.buf, st["target"], grad_scale, st["out"],
                            st["logits"].grad if with_grad else None, st["ws"])


def resolve_loss(loss):
    if isinstance(loss, _PlanLoss):
        return loss
    if isinstance(loss, str):
        key = loss.lower()
        if key in ("charbonnier", "l1", "mse"):
            return SRLoss(key)
        if key in ("binary_crossentropy", "bce"):
            return BinaryCrossentropy()
        if key in ("categorical_crossentropy", "sparse_categorical_crossentropy"):
            return CategoricalCrossentropy()
    raise NotImplementedError(
        f"loss {loss!r}: only the reference's losses are available (charbonnier, l1, mse, BCE, BCE+Dice, "
        "categorical cross-entropy); arbitrary Python losses cannot be lowered to the fused kernels")


# ---- reference-named factories -------------------------------------------------------------------
def make_hybrid_ce_dice_loss(alpha: float, beta: float):
    """Segmenation/code/train_adaptive_unet.py:283-292."""
    return BceDiceLoss(alpha, beta, name="hybrid_ce_dice")


def make_bce_dice_loss(bce_weight: float, dice_weight: float):
    """Segmenation/code/train_adaptive_unet.py:295-304."""
    return BceDiceLoss(bce_weight, dice_weight, name="bce_dice")


class _SegMetric:
    def __init__(self, slot, name):
        self.slot, self.__name__ = slot, name

    def __call__(self, y_true, y_pred):
        out = torch.zeros(8, device=y_pred.device)
        ws = torch.zeros(1 + 3 * y_pred.shape[0], device=y_pred.device)
        ops.bce_dice_loss(y_pred.contiguous(), y_true.to(y_pred.device).float().contiguous(), 1.0, 0.0, 1.0, out, None, ws)
        return out[self.slot]


dice_metric = _SegMetric(2, "dice")                 # per-sample ratios averaged (Segmenation/code/train_adaptive_unet.py:258-265)
iou_metric = _SegMetric(3, "iou")
global_dice_metric = _SegMetric(4, "dice_coefficient")   # one ratio over the batch (Segmenation/code/unet_vinillia.py:94-99)
