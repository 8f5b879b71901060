"""Keras-shaped ``Model``: compile / fit / evaluate / predict / __call__ / summary / weights IO.

Mirrors the slice of ``keras.Model`` the reference's entry points use
(/root/reference: Super_resolution/code/train_adaptive_unet.py:489-494 compile,
:511-516 load_weights, :541-543 summary, :622-632 fit, :676 __call__;
Segmenation/code/train_adaptive_unet.py:503-546).  All arithmetic runs in the
library's sm_100a kernels through ``engine.Plan``; one training step
(zero-grad, forward, loss, backward, gradient all-reduce, Adam) is one CUDA graph.
"""
from __future__ import annotations

import io
import json
import math
import os
import time
import zipfile
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .. import ops
from .._ffi import B200Error
from . import layers as L
from . import losses
from .engine import Plan

_POLICY = {"name": "float32"}
_SEED = {"value": 1234}


def set_global_policy(name: str):
    """keras.mixed_precision.set_global_policy.  "mixed_bfloat16" computes/stores activations in bf16 with fp32 master
    weights.  "mixed_float16" (the reference's policy, train_adaptive_unet.py:471-477) runs the same bf16 kernels -- the
    tcgen05 path here is bf16 -- and, as keras does for that policy, wraps the optimizer in dynamic loss scaling
    (LossScaleOptimizer semantics: scale 2**15, x2 after 2000 finite steps, /2 and step skipped on inf / nan)."""
    if name not in ("float32", "mixed_bfloat16", "mixed_float16"):
        raise ValueError(f"unknown policy {name!r}")
    _POLICY["name"] = name


def global_policy() -> str:
    return _POLICY["name"]


def set_random_seed(seed: int):
    _SEED["value"] = int(seed)


def _init_array(kind, shape, rng):
    if kind in ("glorot_uniform", "glorot_uniform_T"):
        kh, kw, a, b = shape
        fan_in, fan_out = (kh * kw * a, kh * kw * b) if kind == "glorot_uniform" else (kh * kw * b, kh * kw * a)
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return rng.uniform(-lim, lim, size=shape).astype(np.float32)
    if kind == "ones":
        return np.ones(shape, np.float32)
    if kind == "zeros":
        return np.zeros(shape, np.float32)
    raise NotImplementedError(f"initializer {kind!r}")


def _to_snake_case(name: str) -> str:
    """keras.src.utils.naming.to_snake_case (the rule behind the group names of model.weights.h5)."""
    import re
    name = re.sub(r"\W+", "", name)
    name = re.sub("(.)([A-Z][a-z]+)", r"\1_\2", name)
    return re.sub("([a-z])([A-Z])", r"\1_\2", name).lower()


def format_epoch_line(steps: int, seconds: float, logs: Dict[str, float]) -> str:
    """The keras ``verbose=2`` epoch summary (``3/3 - 1s - 412ms/step - loss: 0.0123 - psnr: 31.2 ...``), kept parseable by the
    reference's log exporter (Super_resolution/code/export_log_metrics.py:29-74)."""
    ms = 1000.0 * seconds / max(steps, 1)
    per = f"{ms:.0f}ms/step" if ms >= 1 else f"{ms * 1000:.0f}us/step"
    items = " - ".join(f"{k}: {v:.4f}" if abs(v) >= 1e-3 or v == 0 else f"{k}: {v:.4e}" for k, v in logs.items())
    return f"{steps}/{steps} - {seconds:.0f}s - {per} - {items}"


class History:
    def __init__(self):
        self.epoch: List[int] = []
        self.history: Dict[str, List[float]] = {}


class PendingStep:
    """Handle of an enqueued training step (Model.train_on_batch_async)."""

    def __init__(self, event, host, names, slots):
        self._event, self._host, self._names, self._slots = event, host, names, slots

    def result(self) -> Dict[str, float]:
        self._event.synchronize()
        return {n: float(self._host[s]) for n, s in zip(self._names, self._slots)}


class Model:
    def __init__(self, inputs, outputs, name: Optional[str] = None):
        self.inputs = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        self.outputs = list(outputs) if isinstance(outputs, (list, tuple)) else [outputs]
        if len(self.inputs) != 1 or len(self.outputs) != 1:
            raise NotImplementedError("single-input single-output models only")
        self.name = name or "functional"
        seen, nodes = set(), []

        def visit(t):
            nd = t.node
            if id(nd) in seen:
                return
            seen.add(id(nd))
            for i in nd.inputs:
                visit(i)
            nodes.append(nd)

        visit(self.outputs[0])
        self._nodes = sorted(nodes, key=lambda n: n.seq)
        self.layers: List[L.Layer] = []
        for nd in self._nodes:
            if nd.layer not in self.layers:
                self.layers.append(nd.layer)
        self.optimizer = None
        self.loss = None
        self.metrics_fns = []
        self.stop_training = False
        self._built = False
        self._plans: Dict[tuple, Plan] = {}
        self._graphs: Dict[tuple, object] = {}
        self._dist = None
        self.use_cuda_graph = os.environ.get("B200_NO_CUDA_GRAPH", "0") != "1"
        # filter-gradient kernels on a second stream (forked / joined inside the captured step): see _run_bwd
        # (C2: 5.06 -> 4.61 ms/step; B200_OVERLAP_WGRAD=0 restores the single-stream order)
        self.overlap_wgrad = os.environ.get("B200_OVERLAP_WGRAD", "1") == "1"
        self._side_stream = None
        # EXPERIMENTAL (off): Adam of a layer's kernel range on the second stream right behind its wgrad (N=1 only)
        self.overlap_adam = os.environ.get("B200_OVERLAP_ADAM", "0") == "1"
        # data parallel: BatchNormalization statistics over the GLOBAL batch (per-channel sums all-reduced inside the
        # forward and the backward pass), so N ranks x B/N samples train exactly like one rank x B samples
        # (B200_SYNC_BN=0 keeps per-replica statistics)
        self.sync_batchnorm = os.environ.get("B200_SYNC_BN", "1") == "1"
        # deterministic=True (or B200_DETERMINISTIC=1; set before the first step): filter gradients through per-CTA
        # partial slabs + a fixed-order reduce kernel instead of fp32 vector atomics into the gradient buffer (the
        # default, ~3 % faster, sums in an order that varies run to run).  Activations and activation gradients are
        # bit-reproducible either way (no atomics on those paths: tests/test_model_zz_determinism_gpu.py); the short
        # per-channel reductions (gamma / beta / bias gradients, loss sums) still end in one fp32 atomic per block,
        # reproducible to ~1e-7 relative, not bitwise.
        self.deterministic = os.environ.get("B200_DETERMINISTIC", "0") == "1"
        self.input_shape = self.inputs[0].shape
        self.output_shape = self.outputs[0].shape

    # ------------------------------------------------------------------ parameters
    def _ensure_built(self):
        if self._built:
            return
        if not torch.cuda.is_available():
            raise B200Error("b200unet needs a CUDA device (sm_100a): there is no CPU fallback for the hot path")
        ops.lib()
        self._device = torch.device("cuda", torch.cuda.current_device())
        self._compute_dtype = torch.float32 if _POLICY["name"] == "float32" else torch.bfloat16
        rng = np.random.default_rng(_SEED["value"])
        self._pinfo: Dict[str, tuple] = {}
        self._stem_slots: Dict[str, int] = {}
        # Flat layout of the trainable parameters: [ every rank-4 kernel, in layer order | pad to 512 | every vector
        # (bias, gamma, beta), in layer order ].  Data parallelism shards the optimizer over the kernel region
        # (reduce-scatter -> Adam on the owned shard -> all-gather of the compute-dtype shadow) and keeps the tiny
        # vector region replicated, because the kernels read biases / gamma / beta from the fp32 master directly.
        off_n = 0
        host_t, host_n = [], []
        slots = []
        for ly in self.layers:
            for w in ly.weight_specs:
                val = w["value"] if w["value"] is not None else _init_array(w["init"], w["shape"], rng)
                n = int(np.prod(w["shape"]))
                slot = n
                if self._is_stem_kernel(ly, w):
                    # narrow-input 3x3 kernel [3,3,Cin,Cout] = a [9*Cin][Cout] matrix: its slot is zero-padded to
                    # [64][Cout] so that parameter, gradient and bf16 shadow are directly the operands of the
                    # tcgen05 1x1 path over the im2col tensor (the pad rows have zero gradients and stay zero)
                    slot = 64 * w["shape"][3]
                    self._stem_slots[w["name"]] = slot
                if w["trainable"]:
                    slots.append((w, val, (slot + 63) // 64 * 64, len(w["shape"]) == 4))
                else:
                    self._pinfo[w["name"]] = (False, off_n, w["shape"])
                    host_n.append((off_n, val)); off_n += (n + 63) // 64 * 64
                w["value"] = None
            ly._model = self
        off_t = 0
        for w, val, size, big in slots:
            if big:
                self._pinfo[w["name"]] = (True, off_t, w["shape"])
                host_t.append((off_t, val)); off_t += size
        self._kernel_region = off_t = (off_t + 511) // 512 * 512
        for w, val, size, big in slots:
            if not big:
                self._pinfo[w["name"]] = (True, off_t, w["shape"])
                host_t.append((off_t, val)); off_t += size
        off_t = (off_t + 511) // 512 * 512
        self._nparams = off_t
        flat = np.zeros(max(off_t, 512), np.float32)
        for o, v in host_t:
            flat[o:o + v.size] = v.ravel()
        flat_n = np.zeros(max(off_n, 64), np.float32)
        for o, v in host_n:
            flat_n[o:o + v.size] = v.ravel()
        dev = self._device
        self.P = torch.from_numpy(flat).to(dev)
        self.G = torch.zeros_like(self.P)
        self.NT = torch.from_numpy(flat_n).to(dev)
        self.S = self.P if self._compute_dtype == torch.float32 else torch.empty_like(self.P, dtype=self._compute_dtype)
        self._filters: Dict[str, ops.ConvFilter] = {}
        self._built = True
        self._refresh_shadow()

    def _is_stem_kernel(self, ly, w) -> bool:
        shape = w["shape"]
        return (self._compute_dtype == torch.bfloat16 and w["trainable"] and w["name"].endswith("/kernel")
                and len(shape) == 4 and shape[0] == 3 and shape[1] == 3 and 9 * shape[2] <= 64
                and (shape[3] in (32, 64) or shape[3] % 128 == 0) and type(ly).__name__ == "Conv2D"
                and os.environ.get("B200_STEM_SIMT", "0") != "1")

    def _stem_padded(self, ly):
        """(bf16 shadow as a [1,1,64,Cout] filter, flat fp32 gradient slot) of a stem kernel, or None."""
        key = f"{ly.name}/kernel"
        if key not in self._stem_slots:
            return None
        _, off, shape = self._pinfo[key]
        n = self._stem_slots[key]
        if key not in self._filters:
            f = ops.ConvFilter.__new__(ops.ConvFilter)
            f.hwio = self.S[off:off + n].view(1, 1, 64, shape[3])
            f.kh, f.kw, f.cin, f.cout, f.ohwi = 1, 1, 64, shape[3], None
            self._filters[key] = f
        return self._filters[key], self.G[off:off + n]

    def _refresh_shadow(self):
        self._sync_master()
        if self.S is not self.P:
            ops.cast(self.P, self.S)
        for f in self._filters.values():
            f.repack()

    def _view(self, flat, ly, name):
        key = f"{ly.name}/{name}"
        if key not in self._pinfo:
            return None
        trainable, off, shape = self._pinfo[key]
        n = int(np.prod(shape))
        return flat[off:off + n].view(shape)

    def _param(self, ly, name):
        key = f"{ly.name}/{name}"
        if key not in self._pinfo:
            return None
        return self._view(self.P if self._pinfo[key][0] else self.NT, ly, name)

    def _grad(self, ly, name):
        key = f"{ly.name}/{name}"
        if key not in self._pinfo or not self._pinfo[key][0]:
            return None
        return self._view(self.G, ly, name)

    def _grad_range(self, ly, name):
        key = f"{ly.name}/{name}"
        if key not in self._pinfo or not self._pinfo[key][0]:
            return None
        _, off, shape = self._pinfo[key]
        return (off, self._stem_slots.get(key, int(np.prod(shape))))

    def _shadow(self, ly, name):
        return self._view(self.S, ly, name)

    def _filter(self, ly) -> ops.ConvFilter:
        if ly.name not in self._filters:
            hwio = self._shadow(ly, "kernel")
            kh, kw, cin, cout = hwio.shape
            f = ops.ConvFilter.__new__(ops.ConvFilter)   # a view of the shadow buffer: no copy
            f.hwio, f.kh, f.kw, f.cin, f.cout = hwio, kh, kw, cin, cout
            f.ohwi = None
            # Cout = 64 layers run as CTA pairs (cta_group::2), whose fprop wants the weights K-major ([kh][kw][cout][cin]):
            # a small packed copy (36..74 K elements per layer), refreshed by the plan at the start of every step
            if (hwio.dtype == torch.bfloat16 and (kh, kw) == (3, 3) and cout == 64 and cin % 64 == 0
                    and os.environ.get("B200_CONV_CG2", "1") == "1"):
                f.ohwi = torch.empty((kh, kw, cout, cin), dtype=hwio.dtype, device=hwio.device)
                f.repack()
            self._filters[ly.name] = f
        return self._filters[ly.name]

    # weights API ---------------------------------------------------------------
    def _layer_weights(self, ly):
        self._ensure_built()
        self._sync_master()
        return [self._param(ly, w["name"].split("/", 1)[1]).detach().cpu().numpy().copy() for w in ly.weight_specs]

    def _push_layer_weights(self, ly):
        self._ensure_built()
        for w in ly.weight_specs:
            if w["value"] is not None:
                self._param(ly, w["name"].split("/", 1)[1]).copy_(torch.from_numpy(w["value"]).to(self._device))
                w["value"] = None
        self._refresh_shadow()

    @property
    def weights(self):
        return [w for ly in self.layers for w in ly.weight_specs]

    def get_weights(self):
        return [a for ly in self.layers for a in self._layer_weights(ly)]

    def set_weights(self, values):
        self._ensure_built()
        values = list(values)
        i = 0
        for ly in self.layers:
            k = len(ly.weight_specs)
            if k:
                for w, v in zip(ly.weight_specs, values[i:i + k]):
                    v = np.asarray(v, np.float32)
                    if tuple(v.shape) != w["shape"]:
                        raise ValueError(f"{w['name']}: shape {v.shape} != {w['shape']}")
                    self._param(ly, w["name"].split("/", 1)[1]).copy_(torch.from_numpy(v).to(self._device))
                i += k
        if i != len(values):
            raise ValueError(f"expected {i} weight arrays, got {len(values)}")
        self._refresh_shadow()

    def count_params(self):
        return sum(ly.count_params() for ly in self.layers)

    def get_layer(self, name):
        for ly in self.layers:
            if ly.name == name:
                return ly
        raise ValueError(f"No such layer: {name}")

    # ------------------------------------------------------------------ checkpoints (keras-3 layout)
    def _h5_layer_names(self):
        """The group name keras gives each layer inside model.weights.h5: the snake-cased CLASS name, numbered per class
        in `model.layers` order ("conv2d", "conv2d_1", ... -- keras.src.saving.saving_lib._save_container_state; layer
        names are deliberately not used there).  Our `layers` order equals keras' (pinned by the 885 summary rows)."""
        used, out = {}, []
        for ly in self.layers:
            base = _to_snake_case(type(ly).__name__)
            if base in used:
                used[base] += 1
                out.append(f"{base}_{used[base]}")
            else:
                used[base] = 0
                out.append(base)
        return out

    def _weights_tree(self, include_optimizer=True):
        tree = {"vars": {}, "layers": {}}
        for ly, h5name in zip(self.layers, self._h5_layer_names()):
            tree["layers"][h5name] = {"vars": {str(i): a for i, a in enumerate(self._layer_weights(ly))}}
        opt = self.optimizer
        if include_optimizer and opt is not None and getattr(opt, "_state", None) is not None:
            # keras Adam: [iterations, learning_rate, m_0, v_0, m_1, v_1, ...] over model.trainable_variables
            st = opt._state
            ov = {"0": np.array(int(opt.iterations), dtype=np.int64), "1": np.array(opt.current_lr(), dtype=np.float32)}
            k = 2
            m_all, v_all = st["m"], st["v"]
            if self._sharded():
                m_all, v_all = self._gather_flat(m_all), self._gather_flat(v_all)
            for ly in self.layers:
                for w in ly.weight_specs:
                    if w["trainable"]:
                        nm = w["name"].split("/", 1)[1]
                        ov[str(k)] = self._view(m_all, ly, nm).detach().cpu().numpy().copy()
                        ov[str(k + 1)] = self._view(v_all, ly, nm).detach().cpu().numpy().copy()
                        k += 2
            tree["optimizer"] = {"vars": ov}
        return tree

    def _gather_flat(self, flat):
        """Collective: a copy of a flat per-parameter buffer (Adam m / v) with every sharded bucket all-gathered."""
        out = flat.clone()
        plan = getattr(self, "_last_train_plan", None)
        if plan is not None and self._sharded():
            from ..parallel import all_gather_bucket
            dist, group = self._dist
            for b in self._buckets(plan):
                if b["sharded"]:
                    all_gather_bucket(dist, out, b, group=group)
        return out

    def save(self, path, include_optimizer=True, **kwargs):
        """``model.save("x.keras")`` / ``save_weights("x.weights.h5")`` in the keras-3 on-disk layout: a `.keras` file is a
        zip of config.json, metadata.json and model.weights.h5, whose groups are ``layers/<class_snake[_k]>/vars/<i>`` (+
        ``optimizer/vars/<i>``) -- the file keras' ``load_weights`` reads (train_adaptive_unet.py:511-516,
        evaluate_model.py:57-91).  The HDF5 bytes come from the in-tree writer (h5lite; h5py is absent here).
        config.json describes the layers (class + get_config()) but is not a keras functional config:
        ``keras.models.load_model`` on it takes the reference evaluator's rebuild-and-load_weights branch."""
        from .. import h5lite
        path = str(path)
        blob = h5lite.write_h5(self._weights_tree(include_optimizer))
        if not path.endswith(".keras"):
            with open(path, "wb") as f:
                f.write(blob)
            return
        cfg = {"module": "b200unet.keras", "class_name": "Functional", "name": self.name,
               "layers": [{"class": type(ly).__name__, "name": ly.name, "config": ly.get_config()} for ly in self.layers],
               "weight_names": [w["name"] for w in self.weights]}
        meta = {"keras_version": "3.3.3", "date_saved": time.strftime("%Y-%m-%d@%H:%M:%S"), "writer": "b200unet"}
        with zipfile.ZipFile(path, "w") as z:
            z.writestr("metadata.json", json.dumps(meta))
            z.writestr("config.json", json.dumps(cfg, default=str))
            z.writestr("model.weights.h5", blob)

    save_weights = save

    def load_weights(self, path, skip_mismatch=False, **kwargs):
        """Load a keras-3 `.keras` archive / `.weights.h5` file (written by keras + h5py or by ``save`` above), or a
        round-1 archive with an ``.npz`` member.  Every layer's variable count and shapes are checked against the model;
        a file from another depth / variant fails with the first mismatching layer named."""
        from .. import h5lite
        path = str(path)
        with open(path, "rb") as f:
            head = f.read(8)
        tree = None
        if head[:4] == b"PK\x03\x04":
            with zipfile.ZipFile(path) as z:
                names = z.namelist()
                if "model.weights.h5" in names:
                    tree = h5lite.read_h5(z.read("model.weights.h5"))
                elif "model.weights.npz" in names:           # round-1 archives: positional arrays + stored weight names
                    data = np.load(io.BytesIO(z.read("model.weights.npz")))
                    cfg = json.loads(z.read("config.json")) if "config.json" in names else {}
                    stored = cfg.get("weight_names")
                    mine = [w["name"] for w in self.weights]
                    if stored is not None and list(stored) != mine:
                        raise B200Error(f"{path}: the archive holds weights {stored[:3]}... of another model "
                                        f"({len(stored)} arrays; this model has {len(mine)}: {mine[:3]}...)")
                    self.set_weights([data[k] for k in sorted(data.files)])
                    return
                else:
                    raise B200Error(f"{path}: no model.weights.h5 member -- not a keras-3 archive (members: {names})")
        else:
            try:
                tree = h5lite.read_h5(open(path, "rb").read())
            except h5lite.H5Error as e:
                raise B200Error(f"{path}: neither a .keras zip archive nor a readable HDF5 weights file ({e})") from e
        layers = tree.get("layers")
        if not isinstance(layers, dict):
            raise B200Error(f"{path}: no 'layers' group in the weights file (top-level: {sorted(tree)}); files saved by "
                            "keras 2.x / tf.keras use another layout")
        values = []
        for ly, h5name in zip(self.layers, self._h5_layer_names()):
            specs = ly.weight_specs
            got = layers.get(h5name, {}).get("vars", {}) if isinstance(layers.get(h5name), dict) else {}
            if len(got) != len(specs):
                if skip_mismatch:
                    values += [None] * len(specs)
                    continue
                raise B200Error(f"{path}: layer '{ly.name}' ({h5name}) expects {len(specs)} variables, the file holds "
                                f"{len(got)} -- a checkpoint of a different architecture (depth / variant)?")
            for i, w in enumerate(specs):
                a = np.asarray(got[str(i)])
                if tuple(a.shape) != tuple(w["shape"]):
                    if skip_mismatch:
                        values.append(None)
                        continue
                    raise B200Error(f"{path}: {w['name']} ({h5name}/vars/{i}) has shape {tuple(a.shape)} in the file, "
                                    f"{tuple(w['shape'])} in the model")
                values.append(a.astype(np.float32))
        extra = sorted(set(layers) - set(self._h5_layer_names()))
        if extra and not skip_mismatch:
            raise B200Error(f"{path}: the file holds layers this model does not have: {extra[:5]}")
        if any(v is None for v in values):
            cur = self.get_weights()
            values = [c if v is None else v for v, c in zip(values, cur)]
        self.set_weights(values)
        self._load_optimizer_state(tree.get("optimizer"))

    def _load_optimizer_state(self, node):
        """Adam state of a checkpoint (keras layout, see _weights_tree) into a compiled model; silently skipped when the
        file has none and with a warning when the variable count does not match (keras does the same)."""
        opt = self.optimizer
        if not isinstance(node, dict) or opt is None or self._sharded():
            return
        ov = node.get("vars", {})
        trainable = [(ly, w["name"].split("/", 1)[1], w["shape"]) for ly in self.layers for w in ly.weight_specs if w["trainable"]]
        if len(ov) != 2 + 2 * len(trainable):
            if ov:
                import warnings
                warnings.warn(f"skipping optimizer state: the file holds {len(ov)} optimizer variables, this model's Adam "
                              f"has {2 + 2 * len(trainable)}")
            return
        self._ensure_built()
        opt.ensure_state(self)
        st = opt._state
        k = 2
        for ly, nm, shape in trainable:
            m, v = np.asarray(ov[str(k)], np.float32), np.asarray(ov[str(k + 1)], np.float32)
            if tuple(m.shape) != tuple(shape) or tuple(v.shape) != tuple(shape):
                return
            self._view(st["m"], ly, nm).copy_(torch.from_numpy(m).to(self._device))
            self._view(st["v"], ly, nm).copy_(torch.from_numpy(v).to(self._device))
            k += 2
        it = int(np.asarray(ov["0"]))
        opt.iterations = it
        st["step"].fill_(it)

    # ------------------------------------------------------------------ summary
    def summary(self, print_fn=print, line_length=78):
        # one row per LAYER, as keras prints it: a layer called several times (the shared enc_down / dec_up resizes) appears
        # where it is first called, with the inputs of all its calls and the output shape of its last call
        rows, first = [], {}
        for nd in self._nodes:
            ly = nd.layer
            conn = [t.name for t in nd.inputs]
            shape = "(" + ", ".join("None" if d is None else str(d) for d in nd.output.shape) + ")"
            if id(ly) in first:
                row = rows[first[id(ly)]]
                row[1], row[3] = shape, row[3] + conn
                continue
            first[id(ly)] = len(rows)
            rows.append([f"{ly.name} ({type(ly).__name__})", shape, f"{ly.count_params():,}", conn])
        rows = [(a, b, c, ", ".join(d) or "-") for a, b, c, d in rows]
        w = [max(len(r[i]) for r in rows + [("Layer (type)", "Output Shape", "Param #", "Connected to")]) for i in range(4)]
        fmt = lambda r: " | ".join(s.ljust(w[i]) for i, s in enumerate(r))
        print_fn(f'Model: "{self.name}"')
        print_fn(fmt(("Layer (type)", "Output Shape", "Param #", "Connected to")))
        print_fn("-+-".join("-" * x for x in w))
        for r in rows:
            print_fn(fmt(r))
        total = self.count_params()
        train = sum(int(np.prod(x["shape"])) for x in self.weights if x["trainable"])
        mb = lambda n: f"{n * 4 / 2**20:.2f} MB"
        print_fn(f" Total params: {total:,} ({mb(total)})")
        print_fn(f" Trainable params: {train:,} ({mb(train)})")
        print_fn(f" Non-trainable params: {total - train:,} ({mb(total - train)})")

    # ------------------------------------------------------------------ compile
    def compile(self, optimizer=None, loss=None, metrics=None, jit_compile=False, **kwargs):
        from . import losses as LS
        from . import optimizers as O
        self.optimizer = optimizer if optimizer is not None else O.Adam()
        if isinstance(self.optimizer, str):
            self.optimizer = {"adam": O.Adam}[self.optimizer.lower()]()
        self.loss = LS.resolve_loss(loss)
        if _POLICY["name"] == "mixed_float16" and hasattr(self.optimizer, "dynamic_loss_scale"):
            self.optimizer.dynamic_loss_scale = True
        self.metrics_fns = list(metrics or [])
        self._plans = {k: v for k, v in self._plans.items() if not k[1]}
        self._graphs = {}

    def distribute(self, process_group=None):
        """Batch-sharded data parallelism: one process per GPU, gradients summed with NCCL
        (torch.distributed) inside the captured step.  Call after torch.distributed is initialised."""
        import torch.distributed as dist
        self._dist = (dist, process_group)
        self._graphs = {}
        self._ensure_built()
        dist.broadcast(self.P, src=0, group=process_group)
        dist.broadcast(self.NT, src=0, group=process_group)
        self._refresh_shadow()
        self._peer = None
        if (os.environ.get("B200_DP_P2P", "0") == "1" and self._sharded() and self.S is not self.P
                and dist.get_backend(process_group) == "nccl"):
            self._enable_peer_exchange()

    def _enable_peer_exchange(self):
        """Move the gradient buffer and the weight shadow into memory every rank of the node can map, so that the
        exchange of the sharded step runs on copy engines (peer.py / csrc/peer.cu) instead of NCCL kernels.  All ranks
        agree on the outcome: if the mapping fails anywhere (no peer access, other node) everyone stays on NCCL."""
        from ..peer import PeerExchange
        dist, group = self._dist
        px, err = None, ""
        try:
            px = PeerExchange(dist, group, self._device, self.P.numel(), self.S.dtype,
                              float(os.environ.get("B200_DP_P2P_TIMEOUT", "30")))
        except Exception as e:     # noqa: BLE001 -- any local failure: report, then agree below
            err = f"{type(e).__name__}: {e}"
        ok = torch.tensor([0 if px is None else 1], dtype=torch.int32, device=self._device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            if err:
                print(f"[b200unet] rank {dist.get_rank(group)}: peer exchange unavailable ({err}); using NCCL", flush=True)
            return
        px.S.copy_(self.S)
        self.G, self.S = px.G, px.S
        self._filters, self._plans, self._graphs, self._bucket_cache = {}, {}, {}, None
        self._peer = px
        self._refresh_shadow()

    # ------------------------------------------------------------------ plans
    def _plan(self, batch, training) -> Plan:
        self._ensure_built()
        key = (batch, training)
        if key not in self._plans:
            self._plans[key] = Plan(self, batch, training)
        return self._plans[key]

    def _buckets(self, plan):
        from ..parallel import plan_buckets
        key = id(plan)
        if getattr(self, "_bucket_cache", None) is None:
            self._bucket_cache = {}
        if key not in self._bucket_cache:
            # 64 MB buckets, or a quarter of the kernel region if that is more: measured on 8 GPUs, the 528 MB of C3's
            # gradients in 7 buckets of 64 MB scale 0.925 of independent replicas, in 4 buckets of ~130 MB 0.947
            # (profiles/r02_dp_breakdown_c3_n8_{head,b128}.json): every collective more is a launch + a tail on NCCL's stream
            auto_mb = max(64.0, self._kernel_region * 4 / (1 << 20) / 4) if self._sharded() else 64.0
            elems = int(float(os.environ.get("B200_BUCKET_MB", auto_mb)) * (1 << 20) / 4)
            world = self._world()
            if self._sharded():
                # kernel region: buckets that divide evenly over the ranks (reduce-scatter / sharded Adam / all-gather)
                head = int(float(os.environ.get("B200_HEAD_BUCKET_MB", "1.5")) * (1 << 20) / 4)
                bs = plan_buckets(self._kernel_region, plan.bwd_writes, elems, align=64 * world, head_elems=head)
                # the head (first layers; complete only when backward ends, read first by the next forward pass) is
                # replicated like the vector region: one all-reduce, Adam on every rank, no all-gather to wait for
                replicate_head = os.environ.get("B200_HEAD_REPLICATED", "1") == "1"
                for b in bs:
                    b["sharded"] = not (replicate_head and b.get("head", False))
                # vector region (biases, gamma, beta): one small replicated bucket, complete when backward ends
                bs.append({"lo": self._kernel_region, "hi": self.G.numel(), "ready_after": len(plan.bwd_steps) - 1,
                           "sharded": False})
                bs.sort(key=lambda b: b["ready_after"])
            else:
                bs = plan_buckets(self.G.numel(), plan.bwd_writes, elems)
                for b in bs:
                    b["sharded"] = False
            self._bucket_cache[key] = bs
        return self._bucket_cache[key]

    def _sharded(self) -> bool:
        """Sharded optimizer (ZeRO-1 style) under data parallelism; B200_DP_SHARD=0 selects all-reduce + replicated Adam."""
        return (self._dist is not None and self._world() > 1 and self._kernel_region > 0
                and os.environ.get("B200_DP_SHARD", "1") == "1")

    def _world(self):
        if self._dist is None:
            return 1
        return self._dist[0].get_world_size(self._dist[1])

    # ---- one training step, as launch sequences ------------------------------------------------
    def _seg_forward(self, plan: Plan, st):
        self._seg_forward_part(plan, 0, len(plan.steps))
        self._seg_loss(plan, st)

    def _seg_forward_part(self, plan: Plan, a: int, b: int):
        if a == 0:
            self.G.zero_()
            plan.run_pre()
        for f in plan.steps[a:b]:
            f()

    def _forward_segments(self, plan: Plan):
        """Sharded optimizer: the forward pass cut where it first needs another bucket of the bf16 shadow, so that the
        all-gathers issued behind Adam overlap the NEXT step's forward pass (first layers' bucket first) instead of
        standing between two steps: [(first_step, last_step_exclusive, [indices of the buckets to wait for])]."""
        n = len(plan.steps)
        if not self._sharded():
            return [(0, n, [])]
        buckets = self._buckets(plan)
        segs, seen, start, waits = [], set(), 0, []
        for i, rd in enumerate(plan.fwd_reads):
            need = [k for k, bk in enumerate(buckets) if bk["sharded"] and k not in seen
                    and any(o < bk["hi"] and o + c > bk["lo"] for (o, c) in rd)]
            if need:
                if i > start:
                    segs.append((start, i, waits))
                    start, waits = i, []
                waits = waits + need
                seen.update(need)
        segs.append((start, n, waits))
        return segs

    def _seg_loss(self, plan: Plan, st):
        self.loss.launch(plan, st, grad_scale=1.0 / self._world())
        ls = self.optimizer.loss_scale_state() if hasattr(self.optimizer, "loss_scale_state") else None
        if ls is not None:      # dynamic loss scaling: d(loss)/d(pred) *= scale (device-resident, so the graph stays valid)
            ops.loss_scale_apply(self.loss.grad_tensor(plan, st), ls)

    def _segments(self, plan: Plan):
        """Backward cut at the points where a gradient bucket becomes complete:
        [(first_step, last_step_exclusive, [buckets ready after this segment])]."""
        if self._dist is None:
            return [(0, len(plan.bwd_steps), [])]
        segs, start = [], 0
        by_ready = {}
        for b in self._buckets(plan):
            by_ready.setdefault(min(b["ready_after"], len(plan.bwd_steps) - 1), []).append(b)
        for ready in sorted(by_ready):
            segs.append((start, ready + 1, by_ready[ready]))
            start = ready + 1
        if start < len(plan.bwd_steps):
            segs.append((start, len(plan.bwd_steps), []))
        return segs

    def _run_bwd(self, plan: Plan, a: int, b: int):
        """Backward steps [a, b) in stream order.  With ``overlap_wgrad`` the filter-gradient launches go to a side stream:
        a wgrad only feeds the optimizer, so after waiting for everything issued before it (its dz is final) it may run
        next to the dgrad / norm-backward kernels of the earlier layers; the side stream is joined at the end of the range
        (before the bucket's exchange / the optimizer).  The deep levels' kernels fill a fraction of the 148 SMs, so
        two of them really run side by side; wgrad kernels are serialised among themselves (shared workspace)."""
        if not self.overlap_wgrad:
            for f in plan.bwd_steps[a:b]:
                f()
            return
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream()
        main, side, forked = torch.cuda.current_stream(), self._side_stream, False
        scaled = getattr(self.optimizer, "dynamic_loss_scale", False)     # needs the finite check over the COMPLETE gradient
        early = self._early_adam = [] if (self.overlap_adam and self._dist is None and not scaled) else None
        pending = []          # kernel ranges whose wgrad is on the side stream but whose dgrad may still read the shadow
        if early is not None:
            self.optimizer.advance()
        for i in range(a, b):
            if plan.bwd_tags[i].startswith("wgrad"):
                side.wait_stream(main)          # ... which the wait just issued settles: the layer's dgrad is behind us
                with torch.cuda.stream(side):
                    if early is not None and pending:
                        self.optimizer.apply_ranges(self, pending)
                        early += pending
                    plan.bwd_steps[i]()
                pending = [(lo, lo + n) for lo, n in plan.bwd_writes[i]] if early is not None else []
                forked = True
            else:
                plan.bwd_steps[i]()
        if forked:
            if early is not None and pending:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    self.optimizer.apply_ranges(self, pending)
                early += pending
            main.wait_stream(side)

    def _train_body(self, plan: Plan, st):
        """The launches of one training step, in stream order (eager form; also what gets captured)."""
        self._seg_forward(plan, st)
        works = []
        for (a, b, buckets) in self._segments(plan):
            self._run_bwd(plan, a, b)
            works += self._reduce_async(buckets)
        for w in works:
            w.wait()
        self._check_finite()
        self._apply_optimizer(plan)
        self._gather_updated(plan)
        self._last_train_plan = plan

    def _check_finite(self):
        """Dynamic loss scaling: raise the optimizer's found_inf flag if the (exchanged) gradient holds an inf / nan; under
        data parallelism the flag is max-reduced so that every rank skips -- or takes -- the step together."""
        ls = self.optimizer.loss_scale_state() if hasattr(self.optimizer, "loss_scale_state") else None
        if ls is None:
            return
        ops.loss_scale_check(self.G, ls)
        if self._dist is not None and self._world() > 1:
            dist, group = self._dist
            dist.all_reduce(ls[2:3], op=dist.ReduceOp.MAX, group=group)

    def _reduce_async(self, buckets, coalesce=True):
        """Start the exchange of gradient buckets that just became complete -> handles to wait on (one per bucket, in
        order, with coalesce=False)."""
        if not buckets:
            return []
        from ..parallel import reduce_scatter_bucket
        dist, group = self._dist
        whole = [b for b in buckets if not b["sharded"]]
        if not (coalesce and len(whole) > 1 and dist.get_backend(group) == "nccl"):
            return [reduce_scatter_bucket(dist, self.G, b, group=group, async_op=True) if b["sharded"] else
                    dist.all_reduce(self.G[b["lo"]:b["hi"]], group=group, async_op=True) for b in buckets]
        works = [reduce_scatter_bucket(dist, self.G, b, group=group, async_op=True) for b in buckets if b["sharded"]]
        # replicated head + vector region complete together: one grouped NCCL launch instead of two latencies
        from torch.distributed.distributed_c10d import _coalescing_manager
        with _coalescing_manager(group, self.G.device, async_ops=True) as cm:
            for b in whole:
                dist.all_reduce(self.G[b["lo"]:b["hi"]], group=group)
        works.append(cm)
        return works

    def _update_ranges(self, plan):
        """Flat ranges this rank's Adam updates: everything, or (sharded) its shard of every kernel bucket plus the
        replicated vector region."""
        if not self._sharded():
            return None
        from ..parallel import shard_of
        dist, group = self._dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        return [shard_of(b, rank, world) if b["sharded"] else (b["lo"], b["hi"]) for b in self._buckets(plan)]

    def _apply_optimizer(self, plan):
        early = getattr(self, "_early_adam", None)
        if early:
            # the kernel ranges were updated behind their wgrad kernels (_run_bwd): finish with everything else
            from ..parallel import complement_ranges
            self.optimizer.apply_ranges(self, complement_ranges(early, self.G.numel()))
            self._early_adam = None
            return
        self.optimizer.apply(self, self._update_ranges(plan))

    def _gather_updated(self, plan, order=None):
        """Sharded optimizer: publish the updated shards of the compute-dtype shadow (the fp32 master of the kernel
        region stays sharded until someone reads the weights, see _sync_master).  With `order` (bucket indices in the
        order the next forward pass needs them) the all-gathers are only ISSUED, in that order, and their handles kept in
        self._ag_works: the next step waits for each one right before the forward segment that reads its bucket."""
        if not self._sharded():
            return
        from ..parallel import all_gather_bucket
        dist, group = self._dist
        buckets = self._buckets(plan)
        self._finish_gather()
        if order is None:
            works = [all_gather_bucket(dist, self.S, b, group=group, async_op=True) for b in buckets if b["sharded"]]
            for w in works:
                w.wait()
        else:
            rest = [k for k, b in enumerate(buckets) if b["sharded"] and k not in order]
            self._ag_works = {k: all_gather_bucket(dist, self.S, buckets[k], group=group, async_op=True)
                              for k in list(order) + rest}
        self._master_plan = plan if self.S is not self.P else None

    def _finish_gather(self):
        """Wait (stream-wise) for every shadow all-gather still in flight: before anything but the segmented training
        step reads the shadow (evaluation, weight IO, another plan)."""
        works = getattr(self, "_ag_works", None)
        if works:
            for w in works.values():
                w.wait()
            works.clear()
        if getattr(self, "_peer", None) is not None:
            self._peer.wait_gathers()

    def gathered_gradients(self, plan=None):
        """Collective (every rank): a copy of the flat gradient buffer with every bucket fully reduced -- under the
        sharded optimizer a rank only holds the summed gradient of its own shards."""
        g = self.G.clone()
        plan = plan or getattr(self, "_last_train_plan", None)
        if self._sharded() and plan is not None:
            from ..parallel import all_gather_bucket
            dist, group = self._dist
            for b in self._buckets(plan):
                if b["sharded"]:
                    all_gather_bucket(dist, g, b, group=group)
        return g

    def _sync_master(self):
        """Collective (call on every rank): all-gather the fp32 master of the kernel region after sharded steps."""
        self._finish_gather()
        plan = getattr(self, "_master_plan", None)
        if plan is None or self._dist is None:
            return
        from ..parallel import all_gather_bucket
        dist, group = self._dist
        for b in self._buckets(plan):
            if b["sharded"]:
                all_gather_bucket(dist, self.P, b, group=group)
        self._master_plan = None

    def _train_state(self, batch):
        key = ("train", batch)
        if key in self._graphs:
            return self._graphs[key]
        self._finish_gather()          # (the warm-up below reads the shadow outside the segmented step)
        plan = self._plan(batch, True)
        st = self.loss.make_state(plan)
        self.optimizer.ensure_state(self)
        entry = {"plan": plan, "state": st, "graph": None, "segments": None}
        if self.use_cuda_graph and not plan.sync_bn:      # (host-issued collectives inside the passes: eager launches)
            # warm-up on a side stream (sets kernel attributes, initialises NCCL), then capture
            snap = (self.P.clone(), self.optimizer.snapshot(), self.NT.clone())
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._train_body(plan, st)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self.P.copy_(snap[0]); self.optimizer.restore(snap[1]); self.NT.copy_(snap[2])
            self._refresh_shadow()
            if self._dist is None:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._train_body(plan, st)
                entry["graph"] = g
            else:
                # data parallel: the NCCL all-reduces stay OUTSIDE the graphs (launched eagerly, async, on
                # NCCL's stream) and the backward pass is captured in segments that end where a gradient
                # bucket becomes complete, so each bucket's exchange overlaps the following segment
                fsegs = []
                for (a, b, waits) in self._forward_segments(plan):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._seg_forward_part(plan, a, b)
                    fsegs.append((g, waits))
                entry["fsegs"] = fsegs
                entry["ag_order"] = [k for _, waits in fsegs for k in waits]
                segs = []
                for k, (a, b, buckets) in enumerate(self._segments(plan)):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        if k == 0:
                            self._seg_loss(plan, st)
                        self._run_bwd(plan, a, b)
                    segs.append((g, buckets))
                ga = torch.cuda.CUDAGraph()
                with torch.cuda.graph(ga):
                    self._apply_optimizer(plan)
                entry["segments"], entry["adam_graph"] = segs, ga
                # Pipelined optimizer (B200_DP_PIPELINE_ADAM=1; OFF by default: measured on 8 GPUs it puts the all-gathers on
                # NCCL's stream behind the reduce-scatters of the backward pass, where they delay the tail -- C3 0.93 -> 0.84
                # of independent replicas, C2 0.946 -> 0.938 -- while the all-gathers issued behind Adam find that stream
                # idle during the next forward pass; on 2 GPUs it gains 1 %).  Sharded, no loss scaling (whose finite check
                # needs the complete gradient): a
                # bucket's shard is updated and its shadow all-gathered one backward segment after its reduce-scatter was
                # issued (the dgrad that still reads its first layer's weights sits at the head of that next segment), so
                # at the end of the step only the LAST bucket's reduce-scatter -> Adam -> all-gather chain is exposed.
                if (self._sharded() and not getattr(self.optimizer, "dynamic_loss_scale", False)
                        and os.environ.get("B200_DP_PIPELINE_ADAM", "0") == "1"):
                    from ..parallel import shard_of
                    dist, group = self._dist
                    rank, world = dist.get_rank(group), dist.get_world_size(group)
                    adv = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(adv):
                        self.optimizer.advance()
                    per = {}
                    for k, bk in enumerate(self._buckets(plan)):
                        rng = shard_of(bk, rank, world) if bk["sharded"] else (bk["lo"], bk["hi"])
                        gk = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(gk):
                            self.optimizer.apply_ranges(self, [rng])
                        per[k] = gk
                    entry["adam_advance"], entry["adam_buckets"] = adv, per
        self._graphs[key] = entry
        return entry

    def _run_step(self, entry):
        """Launch one captured (or eager) training step on the data already in the input buffers."""
        if entry["graph"] is not None:
            entry["graph"].replay()
        elif entry["segments"] is not None and getattr(self, "_peer", None) is not None:
            self._run_step_p2p(entry)
        elif entry["segments"] is not None:
            pend = getattr(self, "_ag_works", None) or {}
            for g, waits in entry["fsegs"]:
                for k in waits:                      # the previous step's all-gather of the bucket this segment reads first
                    w = pend.pop(k, None)
                    if w is not None:
                        w.wait()
                g.replay()
            self._finish_gather()
            if "adam_buckets" in entry:
                self._run_bwd_pipelined(entry)
                return
            works = []
            for g, buckets in entry["segments"]:
                g.replay()
                works += self._reduce_async(buckets)
            for w in works:
                w.wait()
            self._check_finite()
            entry["adam_graph"].replay()
            overlap = os.environ.get("B200_DP_OVERLAP_GATHER", "1") == "1"
            self._gather_updated(entry["plan"], entry["ag_order"] if overlap else None)
            self._last_train_plan = entry["plan"]
        else:
            self._train_body(entry["plan"], entry["state"])

    def _run_step_p2p(self, entry):
        """The segmented data-parallel step with the exchange on the copy engines (peer.py): per gradient bucket
        signal -> wait for the peers -> pull my shard's slices -> sum, on the exchange stream behind the backward segment
        that completes the bucket; after Adam, signal -> wait -> pull the peers' shards of the shadow in the order the
        next forward pass reads them (it waits per bucket, on events).  The replicated buckets (head, vector region) still
        go through one grouped NCCL all-reduce."""
        from ..parallel import shard_of
        from ..peer import ADAM_SLOT
        px, plan = self._peer, entry["plan"]
        buckets = self._buckets(plan)
        index = {id(b): k for k, b in enumerate(buckets)}
        main, cs = torch.cuda.current_stream(), px.stream
        works = getattr(self, "_ag_works", None)
        if works:                                     # left by an eager / NCCL step
            for w in works.values():
                w.wait()
            works.clear()
        if px.ag_plan is not plan:
            px.wait_gathers()
        if px.peers_past_adam is not None:
            main.wait_event(px.peers_past_adam)       # nobody pulls last step's gradients out of my G any more: it may be zeroed
        for g, waits in entry["fsegs"]:
            for k in waits:
                ev = px.ag_events.pop(k, None)
                if ev is not None:
                    main.wait_event(ev)
            g.replay()
        px.wait_gathers()
        works = []
        for g, ready in entry["segments"]:
            g.replay()
            mine = [b for b in ready if b["sharded"]]
            if mine:
                for b in mine:
                    px.signal(index[id(b)])
                cs.wait_event(main.record_event())
                with torch.cuda.stream(cs):
                    for b in mine:
                        px.reduce_scatter(index[id(b)], *shard_of(b, px.rank, px.world))
            works += self._reduce_async([b for b in ready if not b["sharded"]])
        main.wait_stream(cs)
        for w in works:
            w.wait()
        self._check_finite()
        entry["adam_graph"].replay()
        px.signal(ADAM_SLOT)
        cs.wait_event(main.record_event())
        order = list(entry["ag_order"])
        order += [k for k, b in enumerate(buckets) if b["sharded"] and k not in order]
        with torch.cuda.stream(cs):
            px.wait_peers(ADAM_SLOT)
            px.peers_past_adam = cs.record_event()
            for k in order:
                px.all_gather([shard_of(buckets[k], r, px.world) for r in range(px.world)])
                px.ag_events[k] = cs.record_event()
        px.ag_plan = plan
        self._master_plan = plan if self.S is not self.P else None
        self._last_train_plan = plan

    def _run_bwd_pipelined(self, entry):
        """Backward segments with the optimizer pipelined behind them (see _train_state): reduce-scatter of a bucket
        right after the segment that completes it; its Adam shard update and the all-gather of its shadow one segment
        later; the handles of the all-gathers are left for the next step's forward segments (self._ag_works)."""
        from ..parallel import all_gather_bucket
        dist, group = self._dist
        plan = entry["plan"]
        buckets = self._buckets(plan)
        index = {id(b): k for k, b in enumerate(buckets)}
        self._ag_works = {}

        def finish(batch):
            for k, w in batch:
                w.wait()
                entry["adam_buckets"][k].replay()
                if buckets[k]["sharded"]:
                    self._ag_works[k] = all_gather_bucket(dist, self.S, buckets[k], group=group, async_op=True)

        entry["adam_advance"].replay()
        prev = []
        for g, ready in entry["segments"]:
            g.replay()
            finish(prev)
            prev = list(zip([index[id(b)] for b in ready], self._reduce_async(ready, coalesce=False)))
        finish(prev)
        self._master_plan = plan if self.S is not self.P else None
        self._last_train_plan = plan

    def release_graphs(self):
        """Drop every captured graph (call before torch.distributed.destroy_process_group)."""
        self._finish_gather()
        self._graphs = {}
        torch.cuda.synchronize()
        if getattr(self, "_peer", None) is not None:      # back to NCCL for whatever this model still trains
            self._peer.close()
            self._peer = None

    def _eval_state(self, batch, with_loss):
        key = ("eval", batch, with_loss)
        if key in self._graphs:
            return self._graphs[key]
        plan = self._plan(batch, False)
        st = self.loss.make_state(plan) if with_loss else None

        def body():
            plan.run_pre()
            plan.run_forward()
            if with_loss:
                self.loss.launch(plan, st, grad_scale=1.0, with_grad=False)

        entry = {"plan": plan, "state": st, "graph": None, "body": body}
        if self.use_cuda_graph:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                body()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body()
            entry["graph"] = g
        self._graphs[key] = entry
        return entry

    @staticmethod
    def _to_device(dst: torch.Tensor, src):
        if isinstance(src, np.ndarray):
            src = torch.from_numpy(src)
        if src.dtype != dst.dtype and not src.is_cuda:
            src = src.to(dst.dtype)
        if (src.is_cuda and src.dtype != dst.dtype and src.dtype in ops._DT and dst.dtype in ops._DT and dst.dim() == 4
                and src.is_contiguous() and src.numel() == dst.numel()):
            ops.copy_tensor(src.reshape(dst.shape), dst)      # device batches (DevicePatchDataset): converting copy kernel
            return
        dst.copy_(src.reshape(dst.shape), non_blocking=True)

    # ------------------------------------------------------------------ steps
    def _feed(self, e, x, y):
        """Inputs of one training step -> the plan's input / target buffers.

        Host batches go through a two-slot staging ring filled on a COPY stream, so the host->device copy of
        step i+1 overlaps the kernels of step i (the compute stream only does a device-to-device copy of the slot);
        device tensors and targets that need a conversion (one-hot -> class ids ...) take the direct path."""
        plan, st = e["plan"], e["state"]
        xb, tgt = plan.input_vals[0].buf, st["target"]
        xs = torch.from_numpy(x) if isinstance(x, np.ndarray) else x
        ys = torch.from_numpy(y) if isinstance(y, np.ndarray) else y
        staged = (not xs.is_cuda and not ys.is_cuda and xs.dtype == xb.dtype and ys.dtype == tgt.dtype
                  and xs.numel() == xb.numel() and ys.numel() == tgt.numel()
                  and type(self.loss).set_target is losses._PlanLoss.set_target)
        if not staged:
            self._to_device(xb, x)
            self.loss.set_target(st, y)
            return
        pipe = e.get("pipe")
        if pipe is None:
            pipe = e["pipe"] = {
                "i": 0, "stream": torch.cuda.Stream(),
                "x": [torch.empty_like(xb) for _ in range(2)], "y": [torch.empty_like(tgt) for _ in range(2)],
                "ready": [torch.cuda.Event() for _ in range(2)], "free": [torch.cuda.Event() for _ in range(2)],
            }
            cur = torch.cuda.current_stream()
            for ev in pipe["free"]:
                ev.record(cur)
        k = pipe["i"] % 2
        pipe["i"] += 1
        cs, cur = pipe["stream"], torch.cuda.current_stream()
        cs.wait_event(pipe["free"][k])            # the step that last read slot k has copied it out
        with torch.cuda.stream(cs):
            pipe["x"][k].copy_(xs.reshape(xb.shape), non_blocking=True)
            pipe["y"][k].copy_(ys.reshape(tgt.shape), non_blocking=True)
            pipe["ready"][k].record(cs)
        cur.wait_event(pipe["ready"][k])
        xb.copy_(pipe["x"][k])
        tgt.copy_(pipe["y"][k])
        pipe["free"][k].record(cur)

    def train_on_batch(self, x, y, return_tensors=False):
        """One optimisation step; returns {"loss": ..., metric: ...} (device tensors if asked)."""
        if self.loss is None:
            raise RuntimeError("compile() the model before training")
        batch = int(x.shape[0])
        e = self._train_state(batch)
        self._feed(e, x, y)
        self.optimizer.before_step()
        self._run_step(e)
        logs = self.loss.logs(e["state"])
        return logs if return_tensors else {k: float(v) for k, v in logs.items()}

    def train_on_batch_async(self, x, y):
        """train_on_batch without the host synchronisation: the step is enqueued (its host->device copy on the copy
        stream, see _feed), the loss / metrics are copied to a pinned host buffer behind it, and a handle is
        returned; ``handle.result()`` waits for THAT step and returns the float logs.  Keep at most two steps in
        flight (the pinned result slots are a ring of two)."""
        if self.loss is None:
            raise RuntimeError("compile() the model before training")
        batch = int(x.shape[0])
        e = self._train_state(batch)
        self._feed(e, x, y)
        self.optimizer.before_step()
        self._run_step(e)
        st = e["state"]
        res = e.get("results")
        if res is None:
            res = e["results"] = {"i": 0, "host": [torch.empty(st["out"].shape, dtype=st["out"].dtype).pin_memory()
                                                   for _ in range(2)],
                                  "done": [torch.cuda.Event() for _ in range(2)]}
        k = res["i"] % 2
        res["i"] += 1
        res["host"][k].copy_(st["out"], non_blocking=True)
        res["done"][k].record(torch.cuda.current_stream())
        names = ["loss"] + list(self.loss.metric_names)
        slots = [0] + list(self.loss.metric_slots)
        return PendingStep(res["done"][k], res["host"][k], names, slots)

    def test_on_batch(self, x, y):
        self._finish_gather()
        batch = int(x.shape[0])
        e = self._eval_state(batch, True)
        plan, st = e["plan"], e["state"]
        self._to_device(plan.input_vals[0].buf, x)
        self.loss.set_target(st, y)
        e["graph"].replay() if e["graph"] is not None else e["body"]()
        return self.loss.logs(st)

    def __call__(self, x, training=False):
        if training:
            raise NotImplementedError("Model(x, training=True): use train_on_batch / fit")
        self._finish_gather()
        batch = int(x.shape[0])
        e = self._eval_state(batch, False)
        plan = e["plan"]
        self._to_device(plan.input_vals[0].buf, x)
        e["graph"].replay() if e["graph"] is not None else e["body"]()
        return plan.output_val.buf.clone()

    def predict(self, x, batch_size=32, verbose=0):
        outs = []
        for i in range(0, len(x), batch_size):
            outs.append(self(x[i:i + batch_size]).float().cpu().numpy())
        return np.concatenate(outs, axis=0)

    # ------------------------------------------------------------------ fit / evaluate
    def evaluate(self, dataset, steps=None, return_dict=False, verbose=0):
        sums, weight = {}, 0.0
        for i, (x, y) in enumerate(dataset):
            if steps is not None and i >= steps:
                break
            logs = self.test_on_batch(x, y)
            b = float(x.shape[0])
            for k, v in logs.items():
                sums[k] = sums.get(k, 0.0) + float(v) * b
            weight += b
        res = self._weighted_mean(sums, weight)
        return res if return_dict else list(res.values())

    def _log_keys(self):
        keys = ["loss"] + list(getattr(self.loss, "extra_metric_names", ())) + list(getattr(self.loss, "metric_names", ()))
        return keys + (["_tp", "_fp", "_fn"] if getattr(self.loss, "keras_metrics", False) else [])

    def _weighted_mean(self, sums: Dict[str, float], weight: float) -> Dict[str, float]:
        """Epoch-level logs as keras computes them: every batch weighted by its sample count (a ragged last batch counts
        for less), i.e. sum(value * batch) / sum(batch).  Data parallel: the sums and the sample count are all-reduced, so
        callbacks (EarlyStopping, ModelCheckpoint, ReduceLROnPlateau) take the same decisions on every rank -- and EVERY
        rank takes part in the collective, also one that saw no batch (it contributes zeros)."""
        keys = self._log_keys() if self.loss is not None else sorted(sums)
        keys = keys + [k for k in sorted(sums) if k not in keys]
        if self._dist is not None and self._world() > 1:
            dist, group = self._dist
            t = torch.tensor([float(sums.get(k, 0.0)) for k in keys] + [float(weight)], dtype=torch.float64,
                             device=self._device)
            dist.all_reduce(t, group=group)
            vals = t.tolist()
            weight = vals[-1]
            sums = {k: v for k, v in zip(keys, vals[:-1])}
        if weight <= 0:
            return {}
        res = {k: float(sums[k]) / weight for k in keys if k in sums}
        if "_tp" in res:     # keras' stateful Precision / Recall: ratios of the counters accumulated over the epoch
            tp, fp, fn = res.pop("_tp"), res.pop("_fp"), res.pop("_fn")
            res["precision"] = tp / (tp + fp) if tp + fp > 0 else 0.0
            res["recall"] = tp / (tp + fn) if tp + fn > 0 else 0.0
        return res

    def fit(self, x=None, y=None, epochs=1, initial_epoch=0, steps_per_epoch=None, validation_data=None,
            validation_steps=None, validation_freq=1, callbacks=None, verbose=1, batch_size=None, **kwargs):
        from .callbacks import CallbackList
        history = History()
        cbs = CallbackList(callbacks or [], self)
        cbs.history = history
        self.stop_training = False
        cbs.on_train_begin()
        start_epoch = cbs.initial_epoch(initial_epoch)
        train_iter = iter(x)
        for epoch in range(start_epoch, epochs):
            if self.stop_training:
                break
            cbs.on_epoch_begin(epoch)
            t0 = time.time()
            acc: Dict[str, torch.Tensor] = {}
            steps, samples = 0, 0.0
            while steps_per_epoch is None or steps < steps_per_epoch:
                try:
                    xb, yb = next(train_iter)
                except StopIteration:
                    if steps_per_epoch is None:
                        break
                    train_iter = iter(x)
                    xb, yb = next(train_iter)
                logs = self.train_on_batch(xb, yb, return_tensors=True)
                b = float(xb.shape[0])
                for k, v in logs.items():
                    acc[k] = v.double() * b if k not in acc else acc[k] + v.double() * b
                steps += 1
                samples += b
            if steps_per_epoch is None:
                train_iter = iter(x)
            logs = self._weighted_mean({k: float(v) for k, v in acc.items()}, samples)
            if validation_data is not None and (epoch + 1) % validation_freq == 0:
                val = self.evaluate(validation_data, steps=validation_steps, return_dict=True)
                logs.update({f"val_{k}": v for k, v in val.items()})
            if hasattr(self.optimizer, "current_lr"):
                logs["learning_rate"] = self.optimizer.current_lr()
            dt = time.time() - t0
            if verbose:
                print(f"Epoch {epoch + 1}/{epochs}")
                print(format_epoch_line(steps, dt, logs), flush=True)
            history.epoch.append(epoch)
            for k, v in logs.items():
                history.history.setdefault(k, []).append(v)
            cbs.on_epoch_end(epoch, logs)
        cbs.on_train_end()
        return history
