"""Per-layer table of the tcgen05 conv kernels at a config's shapes (CUDA events, eager launches),
plus the raw MMA issue-rate probe.  Run on the GPU box:  python tools/conv_table.py [c2|c2alt|c3]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200unet.ops as ops  # noqa: E402
from b200unet._ffi import check  # noqa: E402
from oracle import resize_np  # noqa: E402

CFG = {"c2": (4, 0.25, 128, 64), "c2alt": (4, 0.5, 128, 64), "c3": (5, 0.25, 128, 64)}


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def timeit_graph(fn, reps=10, iters=5):
    """GPU time per call with the launches captured in a CUDA graph (no host launch overhead, as in the
    training step, which replays one graph)."""
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (iters * reps)


def mma_rate():
    L = ops.lib()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    out = {}
    for n in (64, 128, 256):
        for grid in (1, sms):
            cyc = torch.zeros(grid, dtype=torch.int64, device="cuda")
            iters = 2000
            check(L.b200_debug_umma_rate(n, iters, 1024, cyc.data_ptr(), grid, torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            c = cyc.float().mean().item() / (iters * 4)
            out[f"N{n}_grid{grid}"] = {"cycles_per_mma": c, "ideal": 128 * n / 256, "frac": (128 * n / 256) / c}
    return out


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    depth, scale, P, B = CFG[name]
    sizes = resize_np.size_chain(P, scale, depth)
    shapes = set()
    nf, cin = 64, 3
    for d in range(depth):
        shapes.add((sizes[d], cin, nf)); shapes.add((sizes[d], nf, nf)); cin, nf = nf, nf * 2
    shapes.add((sizes[depth], cin, nf)); shapes.add((sizes[depth], nf, nf)); cin = nf
    for d in reversed(range(depth)):
        nf //= 2
        shapes.add((sizes[d], cin, nf)); shapes.add((sizes[d], 2 * nf, nf)); shapes.add((sizes[d], nf, nf)); cin = nf
    rows = []
    for (s, ci, co) in sorted(shapes, reverse=True):
        if ci % 64 or co % 64:
            continue
        x = torch.randn(B, s, s, ci, device="cuda").bfloat16()
        w = (torch.randn(3, 3, ci, co, device="cuda") * 0.05).bfloat16()
        dy = torch.randn(B, s, s, co, device="cuda").bfloat16()
        y = torch.empty(B, s, s, co, device="cuda", dtype=torch.bfloat16)
        dx = torch.empty_like(x)
        dw = torch.empty(3, 3, ci, co, device="cuda")
        bias = torch.zeros(co, device="cuda")
        f = ops.ConvFilter(w, packed=True)
        cws = ops.new_workspace(max(ops.conv2d_workspace(x, f, False), ops.conv2d_workspace(dy, f, True)), "cuda")
        ws = torch.empty(max(ops.conv2d_wgrad_workspace(x, dy, 3, 3), 16) // 4, device="cuda")
        fl = 2.0 * B * s * s * ci * co * 9
        tm = timeit_graph if os.environ.get("TABLE_GRAPH", "1") == "1" else timeit
        t_f = tm(lambda: ops.conv2d_fprop(x, f, bias, y, 1, ws=cws))
        t_d = tm(lambda: ops.conv2d_dgrad(dy, f, dx, ws=cws))
        t_f1 = tm(lambda: ops.conv2d_fprop(x, f, bias, y, 1, ops.ALGO_TCGEN05_1CTA, ws=cws))
        t_d1 = tm(lambda: ops.conv2d_dgrad(dy, f, dx, False, ops.ALGO_TCGEN05_1CTA, ws=cws))
        t_w = tm(lambda: ops.conv2d_wgrad(x, dy, 3, 3, dw, ws))
        rows.append({"hw": s, "cin": ci, "cout": co, "gflop": fl / 1e9,
                     "fprop_us": t_f * 1e3, "fprop_tf": fl / t_f / 1e9,
                     "dgrad_us": t_d * 1e3, "dgrad_tf": fl / t_d / 1e9,
                     "wgrad_us": t_w * 1e3, "wgrad_tf": fl / t_w / 1e9,
                     "fprop_1cta_us": t_f1 * 1e3, "fprop_1cta_tf": fl / t_f1 / 1e9,
                     "dgrad_1cta_us": t_d1 * 1e3, "dgrad_1cta_tf": fl / t_d1 / 1e9})
        r = rows[-1]
        print(f"hw {s:4d} {ci:5d}->{co:5d}  fprop {r['fprop_us']:7.1f} us {r['fprop_tf']:7.1f} TF | dgrad {r['dgrad_us']:7.1f} us "
              f"{r['dgrad_tf']:7.1f} TF | wgrad {r['wgrad_us']:7.1f} us {r['wgrad_tf']:7.1f} TF || 1-CTA fprop {r['fprop_1cta_us']:7.1f} us "
              f"{r['fprop_1cta_tf']:7.1f} TF dgrad {r['dgrad_1cta_us']:7.1f} us {r['dgrad_1cta_tf']:7.1f} TF", flush=True)
    print(json.dumps({"config": name, "mma_rate": mma_rate(), "layers": rows}))


if __name__ == "__main__":
    main()
