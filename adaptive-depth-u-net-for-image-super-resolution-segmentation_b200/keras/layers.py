"""Keras-shaped layer classes (symbolic graph construction only).

These mirror the slice of ``keras.layers`` the reference uses
(/root/reference: Super_resolution/code/train_adaptive_unet.py:200-287,
Segmenation/code/train_adaptive_unet.py:325-362, Segmenation/code/unet_vinillia.py:42-91):
same class names, constructor arguments and call conventions.  Calling a layer on a
symbolic ``KTensor`` records a node; no arithmetic happens here -- ``engine.py`` lowers
the recorded graph to launches of the sm_100a kernels.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_UIDS: Dict[str, int] = {}
_NODE_SEQ = [0]


def clear_session():
    """Reset automatic layer naming (keras.backend.clear_session)."""
    _UIDS.clear()


def _auto_name(base: str) -> str:
    k = _UIDS.get(base, 0)
    _UIDS[base] = k + 1
    return base if k == 0 else f"{base}_{k}"


def _snake(name: str) -> str:
    out = []
    for i, ch in enumerate(name):
        if ch.isupper() and i and (not name[i - 1].isupper() or (i + 1 < len(name) and name[i + 1].islower())):
            out.append("_")
        out.append(ch.lower())
    s = "".join(out)
    return s.replace("conv2_d", "conv2d").replace("sampling2_d", "sampling2d").replace("pooling2_d", "pooling2d")


class KTensor:
    """Symbolic NHWC tensor: shape is (None, H, W, C)."""

    def __init__(self, shape, node=None, name=None):
        self.shape = tuple(shape)
        self.node = node
        self.name = name
        self.dtype = "float32"

    def __repr__(self):
        return f"<KTensor shape={self.shape} from={self.node.layer.name if self.node else None}>"


class Node:
    def __init__(self, layer, inputs: List[KTensor], output_shape):
        _NODE_SEQ[0] += 1
        self.seq = _NODE_SEQ[0]
        self.layer = layer
        self.inputs = inputs
        self.call_index = len(layer._nodes)
        self.output = KTensor(output_shape, self, f"{layer.name}[{self.call_index}][0]")


class Layer:
    """Base class: naming, weight bookkeeping and symbolic ``__call__``."""

    def __init__(self, name: Optional[str] = None, **kwargs):
        self.name = name or _auto_name(_snake(type(self).__name__))
        self._nodes: List[Node] = []
        self.built = False
        # weights: list of dicts {name, shape, init, trainable, value(np or None)}
        self.weight_specs: List[dict] = []
        self.trainable = True
        self._model = None  # set when a Model materialises the weights

    # --- to be overridden ---
    def compute_output_shape(self, input_shapes):
        return input_shapes[0]

    def build(self, input_shapes):
        pass

    def get_config(self):
        return {"name": self.name}

    # --- graph construction ---
    def __call__(self, inputs):
        ins = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        for t in ins:
            if not isinstance(t, KTensor):
                raise TypeError(
                    f"{type(self).__name__} layers are symbolic: call them on keras Input tensors "
                    f"(got {type(t).__name__}); run data through Model.__call__/predict.")
        shapes = [t.shape for t in ins]
        if not self.built:
            self.build(shapes)
            self.built = True
        node = Node(self, ins, self.compute_output_shape(shapes))
        self._nodes.append(node)
        return node.output

    def add_weight(self, name, shape, init, trainable=True):
        self.weight_specs.append({"name": f"{self.name}/{name}", "shape": tuple(shape), "init": init,
                                  "trainable": trainable, "value": None})

    def count_params(self):
        return int(sum(np.prod(w["shape"]) for w in self.weight_specs))

    # weights live in the owning Model's flat buffers once it is built
    def get_weights(self):
        if self._model is not None:
            return self._model._layer_weights(self)
        return [w["value"] for w in self.weight_specs]

    def set_weights(self, values):
        if len(values) != len(self.weight_specs):
            raise ValueError(f"{self.name}: expected {len(self.weight_specs)} arrays, got {len(values)}")
        for w, v in zip(self.weight_specs, values):
            v = np.asarray(v, dtype=np.float32)
            if tuple(v.shape) != w["shape"]:
                raise ValueError(f"{w['name']}: shape {v.shape} != {w['shape']}")
            w["value"] = v
        if self._model is not None:
            self._model._push_layer_weights(self)


class InputLayer(Layer):
    def __init__(self, shape, name=None):
        super().__init__(name=name or _auto_name("input_layer"))
        self.shape = tuple(shape)


def Input(shape, name=None, **kwargs) -> KTensor:
    """keras.Input(shape=(H, W, C), name=...)."""
    layer = InputLayer(shape, name=name)
    node = Node(layer, [], (None,) + tuple(shape))
    layer._nodes.append(node)
    return node.output


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


class Conv2D(Layer):
    """keras.layers.Conv2D -- only what the reference uses: stride 1, padding "same",
    square odd kernels (3 or 1), optional fused activation (relu / sigmoid / softmax)."""

    def __init__(self, filters, kernel_size, strides=1, padding="valid", use_bias=True, activation=None,
                 kernel_initializer="glorot_uniform", bias_initializer="zeros", name=None, **kwargs):
        super().__init__(name=name)
        self.filters = int(filters)
        self.kernel_size = _pair(kernel_size)
        if _pair(strides) != (1, 1):
            raise NotImplementedError("Conv2D: only strides=1 is on the reference's hot path")
        if self.kernel_size not in ((3, 3), (1, 1)):
            raise NotImplementedError("Conv2D: kernel_size must be 3 or 1")
        if padding != "same" and self.kernel_size != (1, 1):
            raise NotImplementedError('Conv2D: only padding="same" is on the reference\'s hot path')
        if activation not in (None, "linear", "relu", "sigmoid", "softmax"):
            raise NotImplementedError(f"Conv2D: activation {activation!r} unsupported")
        self.activation = None if activation == "linear" else activation
        self.use_bias = use_bias
        self.kernel_initializer = kernel_initializer
        self.bias_initializer = bias_initializer

    def build(self, input_shapes):
        cin = input_shapes[0][-1]
        kh, kw = self.kernel_size
        self.add_weight("kernel", (kh, kw, cin, self.filters), self.kernel_initializer)
        if self.use_bias:
            self.add_weight("bias", (self.filters,), self.bias_initializer)

    def compute_output_shape(self, input_shapes):
        return tuple(input_shapes[0][:-1]) + (self.filters,)

    def get_config(self):
        return {"name": self.name, "filters": self.filters, "kernel_size": self.kernel_size, "padding": "same",
                "use_bias": self.use_bias, "activation": self.activation}


class Conv2DTranspose(Layer):
    """keras.layers.Conv2DTranspose(filters, 2, strides=2, padding="same")."""

    def __init__(self, filters, kernel_size, strides=1, padding="valid", use_bias=True, name=None, **kwargs):
        super().__init__(name=name)
        if _pair(kernel_size) != (2, 2) or _pair(strides) != (2, 2):
            raise NotImplementedError("Conv2DTranspose: only kernel 2, strides 2 is on the reference's hot path")
        self.filters = int(filters)
        self.use_bias = use_bias

    def build(self, input_shapes):
        cin = input_shapes[0][-1]
        self.add_weight("kernel", (2, 2, self.filters, cin), "glorot_uniform_T")
        if self.use_bias:
            self.add_weight("bias", (self.filters,), "zeros")

    def compute_output_shape(self, s):
        n, h, w, _ = s[0]
        return (n, 2 * h, 2 * w, self.filters)


class LayerNormalization(Layer):
    """keras.layers.LayerNormalization(axis=-1): epsilon 1e-3, gamma ones, beta zeros."""

    def __init__(self, axis=-1, epsilon=1e-3, name=None, **kwargs):
        super().__init__(name=name)
        if axis not in (-1, 3):
            raise NotImplementedError("LayerNormalization: only axis=-1")
        self.epsilon = float(epsilon)

    def build(self, input_shapes):
        c = input_shapes[0][-1]
        self.add_weight("gamma", (c,), "ones")
        self.add_weight("beta", (c,), "zeros")


class BatchNormalization(Layer):
    """keras.layers.BatchNormalization(): momentum 0.99, epsilon 1e-3."""

    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, name=None, **kwargs):
        super().__init__(name=name)
        self.momentum, self.epsilon = float(momentum), float(epsilon)

    def build(self, input_shapes):
        c = input_shapes[0][-1]
        self.add_weight("gamma", (c,), "ones")
        self.add_weight("beta", (c,), "zeros")
        self.add_weight("moving_mean", (c,), "zeros", trainable=False)
        self.add_weight("moving_variance", (c,), "ones", trainable=False)


class Activation(Layer):
    def __init__(self, activation, name=None, **kwargs):
        super().__init__(name=name)
        if activation != "relu":
            raise NotImplementedError("Activation: only 'relu' is on the reference's hot path")
        self.activation = activation


class Concatenate(Layer):
    def __init__(self, axis=-1, name=None, **kwargs):
        super().__init__(name=name)
        if axis not in (-1, 3):
            raise NotImplementedError("Concatenate: only the channel axis")

    def compute_output_shape(self, s):
        for t in s[1:]:
            if t[:3] != s[0][:3]:
                raise ValueError(f"Concatenate: spatial shapes differ: {s}")
        return tuple(s[0][:3]) + (sum(t[3] for t in s),)


class MaxPooling2D(Layer):
    def __init__(self, pool_size=2, strides=None, padding="valid", name=None, **kwargs):
        super().__init__(name=name)
        if _pair(pool_size) != (2, 2) or (strides is not None and _pair(strides) != (2, 2)):
            raise NotImplementedError("MaxPooling2D: only pool_size 2, stride 2")

    def compute_output_shape(self, s):
        n, h, w, c = s[0]
        return (n, h // 2, w // 2, c)


class UpSampling2D(Layer):
    def __init__(self, size=2, interpolation="nearest", name=None, **kwargs):
        super().__init__(name=name)
        if _pair(size) != (2, 2) or interpolation != "bilinear":
            raise NotImplementedError('UpSampling2D: only size 2, interpolation="bilinear"')

    def compute_output_shape(self, s):
        n, h, w, c = s[0]
        return (n, 2 * h, 2 * w, c)
