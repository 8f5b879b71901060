"""Run-to-run band of the captured training step, and where two runs first part ways.

Built to root-cause the round-1 in-suite failure of test_wgrad_side_stream_matches_single_stream (the loss
trajectories of the single-stream and the two-stream step agreed in isolation and differed by 3.5e-4 in-suite).

For every configuration (single stream / wgrad on the side stream, atomic / slab wgrad) the probe trains the test's
net (depth 3, scale 0.5, 64x64, batch 8, bf16 policy) for a few steps several times and reports

  * the loss trajectories and their spread over repetitions of the SAME configuration (the noise band);
  * after step 0: per-tensor differences of the parameter gradients, and which elements of the bf16 weight
    shadow differ (a shadow weight that rounds the other way is a discrete event the fp32 noise can trigger);
  * with the state (P, S, Adam m / v / step) of the reference run copied into the other run before step 1:
    the first activation (forward order) and the first activation gradient (backward order) that are not
    BIT-identical.  Forward and activation-gradient kernels use no atomics, so on identical weights and inputs
    any differing bit there is a race or an uninitialised read, not summation-order noise.

    python tools/trajectory_probe.py [--reps 3] [--steps 4] [--out gpurun_out/trajectory_probe.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _build(overlap, det=False):
    from b200unet import builders as B
    from b200unet.keras import clear_session, mixed_precision
    from b200unet.keras.optimizers import Adam
    clear_session()
    mixed_precision.set_global_policy("mixed_bfloat16")
    model, _ = B.build_super_resolution_unet(0.5, depth_override=3, input_size=64)
    model.overlap_wgrad, model.overlap_adam = overlap, False
    if hasattr(model, "deterministic"):
        model.deterministic = det
    head = model.get_layer("residual_rgb")
    head.weight_specs[0]["value"] = np.random.default_rng(3).uniform(-0.2, 0.2, (1, 1, 64, 3)).astype(np.float32)
    loss, metrics = B.build_losses_and_metrics("charbonnier")
    model.compile(optimizer=Adam(learning_rate=1e-4), loss=loss, metrics=metrics)
    return model


def _tensor_names(model):
    out = []
    for ly in model.layers:
        for w in ly.weight_specs:
            if w["trainable"]:
                nm = w["name"].split("/", 1)[1]
                rng = model._grad_range(ly, nm)
                out.append((w["name"], rng[0], rng[1]))
    return out


def _snapshot_acts(plan):
    acts, grads = [], []
    for v in plan.all_vals:
        if v.parent is None:
            acts.append((v.name, v.buf.clone()))
            if v.grad is not None:
                grads.append((v.name, v.grad.clone()))
    extra = []
    for op in plan.ops:
        if op.kind == "ln":
            extra.append((f"{op.output.name}:mean", op.mean.clone()))
            extra.append((f"{op.output.name}:rstd", op.rstd.clone()))
    return acts, grads, extra


def _bits_differ(a, b):
    if a.dtype == torch.bfloat16:
        a, b = a.view(torch.int16), b.view(torch.int16)
    elif a.dtype == torch.float32:
        a, b = a.view(torch.int32), b.view(torch.int32)
    return int((a != b).sum().item())


def run_once(lr, hr, overlap, steps, state_from=None, det=False):
    """-> dict with losses, G after step 0, (P, S, m, v) after step 0, activations / gradients of step 1."""
    model = _build(overlap, det)
    res = {"losses": [model.train_on_batch(lr, hr)["loss"]]}
    torch.cuda.synchronize()
    opt = model.optimizer._state
    res["G0"] = model.G.clone()
    if state_from is not None:      # enter step 1 from the reference run's exact state
        model.P.copy_(state_from["P"]); model.S.copy_(state_from["S"])
        opt["m"].copy_(state_from["m"]); opt["v"].copy_(state_from["v"]); opt["step"].copy_(state_from["step"])
    res["state"] = {"P": model.P.clone(), "S": model.S.clone(), "m": opt["m"].clone(), "v": opt["v"].clone(),
                    "step": opt["step"].clone()}
    res["losses"].append(model.train_on_batch(lr, hr)["loss"])
    torch.cuda.synchronize()
    plan = model._train_state(lr.shape[0])["plan"]
    res["acts"], res["grads"], res["extra"] = _snapshot_acts(plan)
    res["G1"] = model.G.clone()
    res["losses"] += [model.train_on_batch(lr, hr)["loss"] for _ in range(steps - 2)]
    res["names"] = _tensor_names(model)
    return res


def compare(ref, other):
    rep = {"losses": other["losses"], "loss_absdiff": [abs(a - b) for a, b in zip(ref["losses"], other["losses"])]}

    def per_tensor(ga, gb):
        worst = []
        for name, off, n in ref["names"]:
            a, b = ga[off:off + n].double(), gb[off:off + n].double()
            den = a.norm().item()
            worst.append((float((a - b).norm().item() / den) if den > 0 else 0.0, name))
        worst.sort(reverse=True)
        return [{"tensor": n, "rel_l2": e} for e, n in worst[:4]]

    rep["G0_worst_tensors"] = per_tensor(ref["G0"], other["G0"])
    rep["G1_worst_tensors"] = per_tensor(ref["G1"], other["G1"])
    rep["G1_bits_differ"] = _bits_differ(ref["G1"], other["G1"])
    sa, sb = ref["state"]["S"], other["state"]["S"]
    rep["shadow_elems_differ_entering_step1"] = _bits_differ(sa, sb)
    rep["master_bits_differ_entering_step1"] = _bits_differ(ref["state"]["P"], other["state"]["P"])
    for key in ("acts", "extra", "grads"):
        first, total = None, 0
        seq = list(zip(ref[key], other[key]))
        if key == "grads":
            seq = seq[::-1]            # backward order
        for (na, a), (nb, b) in seq:
            d = _bits_differ(a, b)
            if d:
                total += 1
                if first is None:
                    first = {"tensor": na, "elems": d, "of": a.numel(),
                             "max_abs": float((a.float() - b.float()).abs().max().item())}
        rep[f"{key}_first_bit_difference"] = first
        rep[f"{key}_tensors_differing"] = total
    return rep


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--out", default=None)
    ap.add_argument("--tag", default="isolated")
    args = ap.parse_args(argv)
    rng = np.random.default_rng(2)
    hr = rng.random((8, 64, 64, 3), dtype=np.float32)
    lr = np.clip(hr + 0.05 * rng.standard_normal(hr.shape).astype(np.float32), 0, 1)
    report = {"tag": args.tag, "wgrad_atomic": os.environ.get("B200_WGRAD_ATOMIC", "1"),
              "poison": os.environ.get("B200_POISON", "0"), "runs": []}
    ref = run_once(lr, hr, False, args.steps)
    report["reference_losses"] = ref["losses"]
    for overlap in (False, True):
        for free in (True, False):         # free-running, then with the reference state forced before step 1
            for r in range(args.reps):
                if not overlap and free and r == 0:
                    continue
                o = run_once(lr, hr, overlap, args.steps, None if free else ref["state"])
                c = compare(ref, o)
                c.update({"overlap_wgrad": overlap, "state_forced": not free, "rep": r})
                report["runs"].append(c)
                del o
                torch.cuda.empty_cache()
    txt = json.dumps(report, indent=1)
    print(txt)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            f.write(txt)
    return report


if __name__ == "__main__":
    main()
