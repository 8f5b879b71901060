"""Lowering of a recorded Keras-shaped graph to a static plan of sm_100a kernel launches.

A ``Plan`` is built once per (batch size, training flag): every activation,
gradient and scratch buffer is allocated up front (device memory is plentiful --
180 GB -- so nothing is aliased except the deliberate cases below), every launch is
a C-ABI call with fixed pointers, and the whole step is captured in a CUDA graph.

Deliberate buffer sharing:
  * Concatenate is never executed: its inputs' buffers (and their gradient buffers)
    are channel slices of the concat output, so producers store straight into it and
    the consumer's dgrad splits the gradient for free.
  * Activation("relu") nodes are folded into the producing conv epilogue or
    normalisation kernel.

The forward/backward order is the graph's topological order / its reverse, which is
what keras' ``train_step`` does with a GradientTape
(/root/reference: Super_resolution/code/train_adaptive_unet.py:489-494, 622-632).
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional

import torch

from .. import ops
from .._ffi import ACT_NONE, ACT_RELU, ACT_SIGMOID
from ..shared.custom_layers import ClippedResidualAdd, ResizeByScale, ResizeToMatch
from . import layers as L


_POISON = os.environ.get("B200_POISON", "0") == "1"


class Val:
    """A concrete activation tensor of the plan (one per KTensor, minus folded nodes)."""

    def __init__(self, name, h, w, c, dtype):
        self.name, self.h, self.w, self.c, self.dtype = name, h, w, c, dtype
        self.buf: Optional[torch.Tensor] = None
        self.grad: Optional[torch.Tensor] = None
        self.parent: Optional["Val"] = None   # concat output this value is a slice of
        self.offset = 0
        self.needs_grad = True
        self._grad_written = False
        self.children: List["Val"] = []

    @property
    def grad_written(self):
        return self._grad_written or (self.parent is not None and self.parent.grad_written)

    def mark_grad_written(self):
        self._grad_written = True


class Op:
    def __init__(self, kind, layer=None, inputs=(), output=None, **attrs):
        self.kind, self.layer, self.inputs, self.output = kind, layer, list(inputs), output
        self.__dict__.update(attrs)


def lower(model) -> (List[Op], Dict[int, Val]):
    """Graph nodes -> op list (pure structure; no device memory yet)."""
    nodes = model._nodes
    consumers: Dict[int, List[L.Node]] = {}
    for nd in nodes:
        for t in nd.inputs:
            consumers.setdefault(id(t), []).append(nd)
    cdt = model._compute_dtype
    vals: Dict[int, Val] = {}
    ops_: List[Op] = []
    producer: Dict[int, Op] = {}   # id(KTensor) -> op that produced its Val

    def new_val(kt, dtype=None):
        _, h, w, c = kt.shape
        v = Val(kt.name, h, w, c, dtype or cdt)
        vals[id(kt)] = v
        return v

    for nd in nodes:
        ly, out = nd.layer, nd.output
        ins = [vals[id(t)] for t in nd.inputs]
        if isinstance(ly, L.InputLayer):
            v = new_val(out, torch.float32)
            v.needs_grad = False
            if cdt != torch.float32:
                vc = Val(out.name + ":cast", v.h, v.w, v.c, cdt)
                vc.needs_grad = False
                ops_.append(Op("cast_input", None, [v], vc))
                v.compute_view = vc
            else:
                v.compute_view = v
            continue
        # model inputs are consumed in the compute dtype, except by the clip-add (fp32 there, as keras)
        cin = [getattr(v, "compute_view", v) for v in ins]
        if isinstance(ly, L.Conv2D):
            act = {None: ACT_NONE, "relu": ACT_RELU, "sigmoid": ACT_SIGMOID, "softmax": ACT_NONE}[ly.activation]
            v = new_val(out)
            if ly.activation == "softmax":
                z = Val(out.name + ":logits", v.h, v.w, v.c, cdt)
                op = Op("conv", ly, cin, z, act=ACT_NONE, norm_bias=False)
                ops_.append(op)
                ops_.append(Op("softmax", None, [z], v))
                producer[id(out)] = ops_[-1]
            else:
                op = Op("conv", ly, cin, v, act=act, norm_bias=False)
                ops_.append(op)
                producer[id(out)] = op
        elif isinstance(ly, (L.LayerNormalization, L.BatchNormalization)):
            v = new_val(out)
            kind = "ln" if isinstance(ly, L.LayerNormalization) else "bn"
            src = producer.get(id(nd.inputs[0]))
            conv_src = src if (src is not None and src.kind == "conv" and src.act == ACT_NONE
                               and len(consumers[id(nd.inputs[0])]) == 1 and src.layer.use_bias) else None
            if conv_src is not None:
                conv_src.norm_bias = True   # its bias gradient comes out of the norm backward kernel
            op = Op(kind, ly, cin, v, relu=False, conv_src=conv_src)
            ops_.append(op)
            producer[id(out)] = op
        elif isinstance(ly, L.Activation):
            src = producer.get(id(nd.inputs[0]))
            single = len(consumers[id(nd.inputs[0])]) == 1
            if src is not None and single and src.kind in ("ln", "bn") and not src.relu:
                src.relu = True
            elif src is not None and single and src.kind == "conv" and src.act == ACT_NONE and not src.norm_bias:
                src.act = ACT_RELU
            else:
                raise NotImplementedError(
                    f"Activation '{ly.name}': a stand-alone ReLU is not on the reference's hot path "
                    "(it must directly follow a Conv2D / LayerNormalization / BatchNormalization)")
            vals[id(out)] = ins[0]          # folded: same value
            producer[id(out)] = src
        elif isinstance(ly, (ResizeByScale, ResizeToMatch, L.UpSampling2D)):
            v = new_val(out)
            aa = getattr(ly, "antialias", False)
            op = Op("resize", ly, [cin[0]], v, antialias=aa)
            ops_.append(op)
            producer[id(out)] = op
        elif isinstance(ly, L.MaxPooling2D):
            v = new_val(out)
            ops_.append(Op("maxpool", ly, cin, v))
            producer[id(out)] = ops_[-1]
        elif isinstance(ly, L.Conv2DTranspose):
            v = new_val(out)
            ops_.append(Op("convT", ly, cin, v))
            producer[id(out)] = ops_[-1]
        elif isinstance(ly, L.Concatenate):
            v = new_val(out)
            off = 0
            copies = []
            for t, vin in zip(nd.inputs, cin):
                others_after = any(c.seq > nd.seq for c in consumers[id(t)] if c is not nd)
                can_alias = vin.parent is None and vin.needs_grad and not others_after and not vin.children
                if can_alias:
                    vin.parent, vin.offset = v, off
                    v.children.append(vin)
                else:
                    copies.append((vin, off))
                off += vin.c
            ops_.append(Op("concat", ly, cin, v, copies=copies))
            producer[id(out)] = ops_[-1]
        elif isinstance(ly, ClippedResidualAdd):
            v = new_val(out)
            ops_.append(Op("clipadd", ly, [ins[0], cin[1]], v))   # the raw fp32 input, as keras casts to fp32
            producer[id(out)] = ops_[-1]
        else:
            raise NotImplementedError(f"layer type {type(ly).__name__} is not supported by the b200 engine")
    return ops_, vals


class Plan:
    def __init__(self, model, batch: int, training: bool):
        self.model, self.batch, self.training = model, batch, training
        self.dev = model._device
        self.ops, self.vals = lower(model)
        self.cdt = model._compute_dtype
        self.input_vals = [self.vals[id(t)] for t in model.inputs]
        self.output_val = self.vals[id(model.outputs[0])]
        self.steps: List[Callable[[], None]] = []       # forward launches
        self.bwd_steps: List[Callable[[], None]] = []   # backward launches (training only)
        self.pre_steps: List[Callable[[], None]] = []   # (unused: weight repacks sit in front of their layer's forward step)
        self._repacked = set()
        self._resample_cache = {}
        # synchronised BatchNorm: training plans of a distributed model that has BatchNormalization layers
        self.sync_bn = bool(training and model._dist is not None and model._world() > 1 and model.sync_batchnorm
                            and any(op.kind == "bn" for op in self.ops))
        self._alloc()
        self._build_forward()
        if training:
            self._build_backward()


    # ------------------------------------------------------------------ buffers
    def _new(self, v: Val, grad=False):
        if _POISON:
            # debug mode (B200_POISON=1): every activation / gradient buffer starts as NaN instead of zeros, so a kernel
            # that reads an element nobody wrote (or accumulates into a never-initialised one) poisons the loss
            return torch.full((self.batch, v.h, v.w, v.c), float("nan"), dtype=v.dtype, device=self.dev)
        return torch.zeros((self.batch, v.h, v.w, v.c), dtype=v.dtype, device=self.dev)

    def _scratch(self, shape, dtype):
        """Scratch every element of which is written before it is read (NaN-filled under B200_POISON=1)."""
        if _POISON and dtype in (torch.float32, torch.float64, torch.bfloat16):
            return torch.full(shape if isinstance(shape, tuple) else (shape,), float("nan"), dtype=dtype, device=self.dev)
        return torch.empty(shape, dtype=dtype, device=self.dev)

    def _alloc(self):
        all_vals = []
        seen = set()

        def add(v):
            if id(v) not in seen:
                seen.add(id(v)); all_vals.append(v)

        for v in self.vals.values():
            add(v)
        for op in self.ops:
            for v in op.inputs:
                add(v)
            add(op.output)
        roots = [v for v in all_vals if v.parent is None]
        for v in roots:
            v.buf = self._new(v)
            if self.training and v.needs_grad:
                v.grad = self._new(v)
        for v in all_vals:
            if v.parent is not None:
                p = v.parent
                v.buf = p.buf[..., v.offset:v.offset + v.c]
                if self.training:
                    v.grad = p.grad[..., v.offset:v.offset + v.c]
        self.all_vals = all_vals

    def _plans(self, in_hw, out_hw, antialias):
        key = (in_hw, out_hw, antialias)
        if key not in self._resample_cache:
            self._resample_cache[key] = (ops.ResamplePlan(in_hw[0], out_hw[0], antialias, self.dev),
                                         ops.ResamplePlan(in_hw[1], out_hw[1], antialias, self.dev))
        return self._resample_cache[key]

    # ------------------------------------------------------------------ forward
    def _build_forward(self):
        m = self.model
        npix = lambda v: self.batch * v.h * v.w
        self.step_tags: List[str] = []

        self.fwd_reads: List[list] = []   # flat kernel ranges (offset, count) of the compute-dtype shadow each forward step reads
        self._cur_reads = []

        class _Tagged(list):
            def append(inner, fn, _tags=self.step_tags):
                _tags.append(self._cur_tag)
                self.fwd_reads.append(list(self._cur_reads))
                self._cur_reads = []
                list.append(inner, fn)

        def reads(ly):
            r = m._grad_range(ly, "kernel")
            if r is not None:
                self._cur_reads.append(r)

        S = self.steps = _Tagged()
        # split-K scratch of the small-spatial (deep-level) convolutions, fprop and dgrad: ONE buffer per plan, handed to
        # every call that may need it (all of them run on the plan's main stream, so they may share it)
        ws_need = 0
        for op in self.ops:
            if op.kind == "conv" and op.inputs[0].buf.dtype == torch.bfloat16 and op.inputs[0].c % 64 == 0:
                filt = m._filter(op.layer)
                ws_need = max(ws_need, ops.conv2d_workspace(op.inputs[0].buf, filt, False))
                if self.training:
                    ws_need = max(ws_need, ops.conv2d_workspace(op.output.buf, filt, True))
            elif op.kind == "convT" and op.inputs[0].buf.dtype == torch.bfloat16:
                x, y = op.inputs[0], op.output
                ws_need = max(ws_need, ops.convT2x2_workspace(x.buf, x.c, y.c, False))
                if self.training:
                    ws_need = max(ws_need, ops.convT2x2_workspace(y.buf, y.c, x.c, True))
        self.conv_ws = ops.new_workspace(ws_need, self.dev)
        for op in self.ops:   # narrow-input 3x3 convs (the RGB stem): im2col tensor for the tcgen05 1x1 path
            if op.kind == "conv" and op.layer.kernel_size == (3, 3) and m._stem_padded(op.layer) is not None:
                x = op.inputs[0]
                op.xcol = torch.zeros((self.batch, x.h, x.w, 64), dtype=torch.bfloat16, device=self.dev)
        # a model input whose bf16 cast is read by stem convolutions only: im2col rounds to bf16 itself, skip the cast
        for cast in [o for o in self.ops if o.kind == "cast_input"]:
            readers = [o for o in self.ops if cast.output in o.inputs]
            if readers and all(getattr(o, "xcol", None) is not None for o in readers):
                for o in readers:
                    o.stem_raw = cast.inputs[0]
                cast.skip = True
        for op in self.ops:   # Conv2D whose only consumer is a LayerNormalization: run as one fused call
            if op.kind == "ln" and op.conv_src is not None and op.conv_src.output is op.inputs[0] \
                    and op.conv_src.layer.kernel_size == (3, 3):
                op.conv_src.fused_into_ln = True
        for op in self.ops:
            k = op.kind
            self._cur_tag = k + (":tc" if k == "conv" and self.is_tc(op) else (":simt" if k == "conv" else ""))
            if k == "cast_input":
                if getattr(op, "skip", False):
                    continue
                S.append(lambda a=op.inputs[0], b=op.output: ops.copy_tensor(a.buf, b.buf))
            elif k == "conv":
                if getattr(op, "fused_into_ln", False):
                    continue     # executed by the following LayerNormalization op (one fused kernel)
                filt, bias = m._filter(op.layer), m._param(op.layer, "bias")
                src = self._stem_source(op, S)
                self._emit_repack(filt, S)
                reads(op.layer)
                if src is not None:
                    filt = m._stem_padded(op.layer)[0]
                    S.append(lambda xc=src, f=filt, b=bias, y=op.output, o=op:
                             ops.conv2d_fprop(xc, f, b, y.buf, o.act, ws=self.conv_ws))
                else:
                    S.append(lambda x=op.inputs[0], f=filt, b=bias, y=op.output, o=op:
                             ops.conv2d_fprop(x.buf, f, b, y.buf, o.act, ws=self.conv_ws))
            elif k == "ln":
                g, b = m._param(op.layer, "gamma"), m._param(op.layer, "beta")
                op.mean = self._scratch(npix(op.output), torch.float32)
                op.rstd = self._scratch(npix(op.output), torch.float32)
                cs = op.conv_src
                if cs is not None and cs.output is op.inputs[0] and cs.layer.kernel_size == (3, 3):
                    # Conv2D -> LayerNormalization -> ReLU: one call (fused tcgen05 epilogue for Cout 64/128)
                    self._cur_tag = "conv+ln" + (":tc" if self.is_tc(cs) else ":simt")
                    src = self._stem_source(cs, S)
                    xin = src if src is not None else cs.inputs[0].buf
                    filt = m._stem_padded(cs.layer)[0] if src is not None else m._filter(cs.layer)
                    self._emit_repack(filt, S)
                    reads(cs.layer)
                    self._cur_tag = "conv+ln" + (":tc" if self.is_tc(cs) else ":simt")
                    S.append(lambda x=xin, f=filt, cb=m._param(cs.layer, "bias"),
                             z=op.inputs[0], y=op.output, o=op, g=g, b=b:
                             ops.conv2d_ln_fprop(x, f, cb, g, b, o.layer.epsilon, o.relu, z.buf, y.buf, o.mean, o.rstd,
                                                 ws=self.conv_ws))
                else:
                    S.append(lambda z=op.inputs[0], y=op.output, o=op, g=g, b=b:
                             ops.layernorm_fwd(z.buf, g, b, o.layer.epsilon, o.relu, y.buf, o.mean, o.rstd))
            elif k == "bn":
                g, b = m._param(op.layer, "gamma"), m._param(op.layer, "beta")
                mm, mv = m._param(op.layer, "moving_mean"), m._param(op.layer, "moving_variance")
                c = op.output.c
                op.save_mean = self._scratch(c, torch.float32)
                op.save_rstd = self._scratch(c, torch.float32)
                op.stats_ws = self._scratch(2 * c, torch.float64)
                if self.training and self.sync_bn:
                    # synchronised BatchNorm: per-channel sums over the GLOBAL batch (all-reduced between the phases;
                    # a host-issued collective, so plans with it are not captured in a CUDA graph)
                    def fwd_sync(z=op.inputs[0], y=op.output, o=op, g=g, b=b, mm=mm, mv=mv):
                        dist, group = m._dist
                        ops.batchnorm_stats(z.buf, o.stats_ws)
                        dist.all_reduce(o.stats_ws, group=group)
                        ops.batchnorm_fwd_apply(z.buf, g, b, o.layer.epsilon, o.layer.momentum, o.relu, y.buf, o.save_mean,
                                                o.save_rstd, mm, mv, o.stats_ws, npix(y) * m._world())
                    S.append(fwd_sync)
                elif self.training:
                    S.append(lambda z=op.inputs[0], y=op.output, o=op, g=g, b=b, mm=mm, mv=mv:
                             ops.batchnorm_fwd_train(z.buf, g, b, o.layer.epsilon, o.layer.momentum, o.relu, y.buf,
                                                     o.save_mean, o.save_rstd, mm, mv, o.stats_ws))
                else:
                    S.append(lambda z=op.inputs[0], y=op.output, o=op, g=g, b=b, mm=mm, mv=mv:
                             ops.batchnorm_fwd_infer(z.buf, g, b, o.layer.epsilon, o.relu, mm, mv, y.buf))
            elif k == "resize":
                x, y = op.inputs[0], op.output
                if (x.h, x.w) == (y.h, y.w):
                    op.identity = True
                    S.append(lambda x=x, y=y: ops.copy_tensor(x.buf, y.buf))
                else:
                    op.identity = False
                    op.ph, op.pw = self._plans((x.h, x.w), (y.h, y.w), op.antialias)
                    S.append(lambda x=x, y=y, o=op: ops.resample2d(x.buf, y.buf, o.ph, o.pw))
            elif k == "maxpool":
                S.append(lambda x=op.inputs[0], y=op.output: ops.maxpool2_fwd(x.buf, y.buf))
            elif k == "convT":
                kern, bias = m._shadow(op.layer, "kernel"), m._param(op.layer, "bias")
                reads(op.layer)
                S.append(lambda x=op.inputs[0], y=op.output, kk=kern, b=bias:
                         ops.convT2x2_fprop(x.buf, kk, b, y.buf, ws=self.conv_ws))
            elif k == "concat":
                for vin, off in op.copies:
                    S.append(lambda a=vin, y=op.output, off=off: ops.copy_tensor(a.buf, y.buf[..., off:off + a.c]))
            elif k == "clipadd":
                S.append(lambda a=op.inputs[0], r=op.inputs[1], y=op.output: ops.clipadd_fwd(a.buf, r.buf, y.buf))
            elif k == "softmax":
                S.append(lambda z=op.inputs[0], p=op.output: ops.softmax_fwd(z.buf, p.buf))
            else:
                raise AssertionError(k)

    def _emit_repack(self, filt, S):
        """The K-major pack of a filter that has one (Model._filter) is re-derived from the shadow right before the layer's
        first forward use -- behind the all-gather of its bucket under the sharded optimizer."""
        if filt is None or filt.ohwi is None or id(filt) in self._repacked:
            return
        self._repacked.add(id(filt))
        tag, rd = self._cur_tag, self._cur_reads
        self._cur_tag, self._cur_reads = "repack", []
        S.append(filt.repack)
        self._cur_tag, self._cur_reads = tag, rd

    def _stem_source(self, op, S):
        """Narrow-input 3x3 conv under the bf16 policy: emit the im2col launch and return the [B,H,W,64]
        tensor the tcgen05 1x1 path consumes (None when the op is not a stem)."""
        if getattr(op, "xcol", None) is None:
            return None
        x = getattr(op, "stem_raw", None) or op.inputs[0]     # the raw fp32 model input when op reads its bf16 cast
        tag = self._cur_tag
        self._cur_tag = "im2col"
        S.append(lambda x=x, xc=op.xcol: ops.im2col3x3(x.buf, xc))
        self._cur_tag = tag
        return op.xcol

    # ------------------------------------------------------------------ backward
    def _build_backward(self):
        m = self.model
        self.bwd_tags: List[str] = []

        self.bwd_writes: List[list] = []   # flat-G ranges (offset, count) written by each backward step

        class _Tagged(list):
            def append(inner, fn, _tags=self.bwd_tags):
                _tags.append(self._cur_tag)
                self.bwd_writes.append(list(self._cur_writes))
                self._cur_writes = []
                list.append(inner, fn)

        self._cur_writes = []

        def writes(ly, *names):
            for nm in names:
                r = m._grad_range(ly, nm)
                if r is not None:
                    self._cur_writes.append(r)

        B = self.bwd_steps = _Tagged()
        ws_bytes = 0
        for op in self.ops:
            if op.kind == "conv" and op.layer.kernel_size == (3, 3):
                if getattr(op, "xcol", None) is not None:
                    ws_bytes = max(ws_bytes, ops.conv2d_wgrad_workspace(op.xcol, op.output.buf, 1, 1))
                else:
                    ws_bytes = max(ws_bytes, ops.conv2d_wgrad_workspace(op.inputs[0].buf, op.output.buf, 3, 3))
        self.wgrad_ws = self._scratch(max(ws_bytes, 16) // 4, torch.float32)

        def write_flag(v: Val) -> bool:
            acc = v.grad_written
            v.mark_grad_written()
            return acc

        self.output_val.mark_grad_written()   # the loss kernel writes d(output)
        for op in reversed(self.ops):
            k, out = op.kind, op.output
            if k == "cast_input" or not out.needs_grad:
                continue
            self._cur_tag = k
            if k == "conv":
                x, ly = op.inputs[0], op.layer
                filt = m._filter(ly)
                dbias = m._grad(ly, "bias")
                kh = ly.kernel_size[0]
                sfx = ":tc" if self.is_tc(op) else ":simt"
                self._cur_tag = "bias_act"
                if op.act != ACT_NONE or (not op.norm_bias and dbias is not None):
                    writes(ly, "bias")
                if op.act != ACT_NONE:
                    B.append(lambda o=out, a=op.act, db=dbias: ops.bias_act_bwd(o.grad, o.buf, a, o.grad, db))
                elif not op.norm_bias and dbias is not None:
                    B.append(lambda o=out, db=dbias: ops.bias_act_bwd(o.grad, o.buf, ACT_NONE, o.grad, db))
                dw = m._grad(ly, "kernel").view(-1)
                self._cur_tag = "wgrad" + sfx
                writes(ly, "kernel")
                # the step zeroes the flat gradient buffer first, so the tcgen05 wgrad kernels add their partial sums
                # straight into it (vector atomics): no partial slabs, no reduce launch
                atomic = (sfx == ":tc" and os.environ.get("B200_WGRAD_ATOMIC", "1") == "1"
                          and not getattr(m, "deterministic", False))
                if getattr(op, "xcol", None) is not None:     # stem: 1x1 wgrad over the im2col tensor
                    if atomic:
                        B.append(lambda xc=op.xcol, o=out, dw=m._stem_padded(ly)[1]: ops.conv2d_wgrad_atomic(xc, o.grad, 1, 1, dw))
                    else:
                        B.append(lambda xc=op.xcol, o=out, dw=m._stem_padded(ly)[1]:
                                 ops.conv2d_wgrad(xc, o.grad, 1, 1, dw, self.wgrad_ws))
                elif atomic:
                    B.append(lambda x=x, o=out, dw=dw, kh=kh: ops.conv2d_wgrad_atomic(x.buf, o.grad, kh, kh, dw))
                else:
                    B.append(lambda x=x, o=out, dw=dw, kh=kh: ops.conv2d_wgrad(x.buf, o.grad, kh, kh, dw, self.wgrad_ws))
                if x.needs_grad:
                    acc = write_flag(x)
                    # (measured, tools/lnb_probe.py: the fused epilogue is latency-bound -- 194 us against 75 + 85 us for
                    # the two kernels at 64 x 128^2 x 64 -- so the engine keeps them apart unless B200_FUSE_LN_BWD=1)
                    lnop = self._ln_producer(x) if (sfx == ":tc" and not acc and getattr(op, "xcol", None) is None
                                                    and os.environ.get("B200_FUSE_LN_BWD", "0") == "1") else None
                    if lnop is not None and ops.conv2d_dgrad_ln_bwd_supported(out.grad, filt, lnop.inputs[0].grad):
                        # the input is a LayerNormalization(+ReLU) output read by this convolution only: its backward pass
                        # runs in this dgrad's epilogue (dy of the LayerNorm output never reaches memory); the "ln" op
                        # below then has nothing left to do
                        z, lly = lnop.inputs[0], lnop.layer
                        lnop.bwd_fused = True
                        z.mark_grad_written()
                        dbias = m._grad(lnop.conv_src.layer, "bias") if lnop.conv_src is not None else None
                        writes(lly, "gamma", "beta")
                        if lnop.conv_src is not None:
                            writes(lnop.conv_src.layer, "bias")
                        self._cur_tag = "dgrad+ln" + sfx
                        B.append(lambda o=out, f=filt, z=z, lo=lnop, g=m._param(lly, "gamma"), b=m._param(lly, "beta"),
                                 dg=m._grad(lly, "gamma"), db=m._grad(lly, "beta"), dbias=dbias:
                                 ops.conv2d_dgrad_ln_bwd(o.grad, f, z.buf, lo.mean, lo.rstd, g, b, lo.relu, z.grad, dg, db, dbias))
                    else:
                        self._cur_tag = "dgrad" + sfx
                        B.append(lambda x=x, o=out, f=filt, acc=acc: ops.conv2d_dgrad(o.grad, f, x.grad, acc, ws=self.conv_ws))
            elif k == "ln":
                z, ly = op.inputs[0], op.layer
                if getattr(op, "bwd_fused", False):
                    continue                        # done in the epilogue of the consumer's dgrad (see "conv" above)
                assert not z.grad_written, "LayerNormalization input must have a single consumer"
                z.mark_grad_written()
                dbias = m._grad(op.conv_src.layer, "bias") if op.conv_src is not None else None
                writes(ly, "gamma", "beta")
                if op.conv_src is not None:
                    writes(op.conv_src.layer, "bias")
                B.append(lambda z=z, o=out, op=op, g=m._param(ly, "gamma"), b=m._param(ly, "beta"),
                         dg=m._grad(ly, "gamma"), db=m._grad(ly, "beta"), dbias=dbias:
                         ops.layernorm_bwd(o.grad, z.buf, op.mean, op.rstd, g, b, op.relu, z.grad, dg, db, dbias))
            elif k == "bn":
                z, ly = op.inputs[0], op.layer
                assert not z.grad_written, "BatchNormalization input must have a single consumer"
                z.mark_grad_written()
                writes(ly, "gamma", "beta")
                if self.sync_bn:
                    def bwd_sync(z=z, o=out, op=op, g=m._param(ly, "gamma"), b=m._param(ly, "beta"),
                                 dg=m._grad(ly, "gamma"), db=m._grad(ly, "beta")):
                        dist, group = m._dist
                        ops.batchnorm_stats(z.buf, op.stats_ws, o.grad, op.save_mean, op.save_rstd, g, b, op.relu, dg, db)
                        dist.all_reduce(op.stats_ws, group=group)
                        ops.batchnorm_bwd_apply(o.grad, z.buf, op.save_mean, op.save_rstd, g, b, op.relu, z.grad,
                                                op.stats_ws, self.batch * z.h * z.w * m._world())
                    B.append(bwd_sync)
                else:
                    B.append(lambda z=z, o=out, op=op, g=m._param(ly, "gamma"), b=m._param(ly, "beta"),
                             dg=m._grad(ly, "gamma"), db=m._grad(ly, "beta"):
                             ops.batchnorm_bwd(o.grad, z.buf, op.save_mean, op.save_rstd, g, b, op.relu, z.grad, dg, db,
                                               op.stats_ws))
            elif k == "resize":
                x = op.inputs[0]
                if x.needs_grad:
                    acc = write_flag(x)
                    if op.identity and not acc:
                        B.append(lambda x=x, o=out: ops.copy_tensor(o.grad, x.grad))
                    elif op.identity:
                        # same-size resize whose input also feeds a skip connection (e.g. the 1x1 -> 1x1 level of a
                        # depth-5 net at scale 0.25): dx += dy through the identity span tables
                        op.ph, op.pw = self._plans((x.h, x.w), (out.h, out.w), op.antialias)
                        B.append(lambda x=x, o=out, op=op: ops.resample2d_bwd(o.grad, x.grad, op.ph, op.pw, True))
                    else:
                        B.append(lambda x=x, o=out, op=op, acc=acc: ops.resample2d_bwd(o.grad, x.grad, op.ph, op.pw, acc))
            elif k == "maxpool":
                x = op.inputs[0]
                if x.needs_grad:
                    acc = write_flag(x)
                    B.append(lambda x=x, o=out, acc=acc: ops.maxpool2_bwd(x.buf, o.buf, o.grad, x.grad, acc))
            elif k == "convT":
                x, ly = op.inputs[0], op.layer
                writes(ly, "kernel", "bias")
                B.append(lambda x=x, o=out, dk=m._grad(ly, "kernel"), db=m._grad(ly, "bias"):
                         ops.convT2x2_wgrad(x.buf, o.grad, dk, db))
                if x.needs_grad:
                    if write_flag(x):
                        raise NotImplementedError("Conv2DTranspose input with several consumers")
                    B.append(lambda x=x, o=out, kk=m._shadow(ly, "kernel"):
                             ops.convT2x2_dgrad(o.grad, kk, x.grad, ws=self.conv_ws))
            elif k == "concat":
                for vin, off in op.copies:
                    if vin.needs_grad:
                        if write_flag(vin):
                            raise NotImplementedError("copied concat input with several consumers")
                        B.append(lambda a=vin, o=out, off=off: ops.copy_tensor(o.grad[..., off:off + a.c], a.grad))
                # aliased inputs: their gradient IS the slice of d(concat), already written by the consumer
            elif k == "clipadd":
                inp, res = op.inputs
                if write_flag(res):
                    raise NotImplementedError("residual with several consumers")
                B.append(lambda a=inp, r=res, o=out: ops.clipadd_bwd(a.buf, r.buf, o.grad, r.grad))
            elif k == "softmax":
                # softmax + categorical cross-entropy: the loss kernel writes d(logits) directly
                op.inputs[0].mark_grad_written()
            else:
                raise AssertionError(k)

    def _ln_producer(self, v: Val):
        """The LayerNormalization op whose output is `v`, when `v` is read by ONE op only (so that the gradient this
        consumer produces is the whole gradient), is not part of a concat buffer and is 64 bf16 channels wide."""
        if v.parent is not None or v.children or v.c != 64 or v.dtype != torch.bfloat16:
            return None
        prod = [o for o in self.ops if o.output is v]
        cons = [o for o in self.ops if v in o.inputs]
        if len(prod) != 1 or prod[0].kind != "ln" or len(cons) != 1 or v is self.output_val:
            return None
        return prod[0]

    @staticmethod
    def is_tc(op) -> bool:
        """True when the conv op's shapes select the tcgen05 kernels (mirrors conv_tc_supported)."""
        x, y, ly = op.inputs[0], op.output, op.layer
        if getattr(op, "xcol", None) is not None:
            return True    # stem: im2col + 1x1 on the tcgen05 kernels
        return (ly.kernel_size == (3, 3) and x.dtype == torch.bfloat16 and (x.c % 64 == 0 or x.c == 32)
                and (y.c in (32, 64) or y.c % 128 == 0))

    # ------------------------------------------------------------------ execution
    def run_pre(self):
        for f in self.pre_steps:
            f()

    def run_forward(self):
        for f in self.steps:
            f()

    def run_backward(self):
        for f in self.bwd_steps:
            f()
