"""Probe of the weight-stationary tcgen05.mma.ws form: (1) layout / numerics of two A tiles multiplied by one B tile
kept in the collector (fill + lastuse), (2) issue rate against ordinary MMAs in the same two-tile loop."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200unet.ops as ops  # noqa: E402
from b200unet._ffi import check  # noqa: E402

torch.manual_seed(0)
a = (torch.randn(256, 64, device="cuda")).bfloat16()
b = (torch.randn(64, 64, device="cuda")).bfloat16()
out = torch.zeros(256, 64, device="cuda")
ops.umma_probe(a, b, 0, 1024, 0, 2, out)
torch.cuda.synchronize()
ref = a.float() @ b.float().t()
err = ((out - ref).norm() / ref.norm()).item()
print(f"ws pair probe: rel-L2 vs A @ B^T = {err:.3e}  (rows 0..127 {((out[:128]-ref[:128]).norm()/ref[:128].norm()).item():.2e}, rows 128..255 {((out[128:]-ref[128:]).norm()/ref[128:].norm()).item():.2e})")
L = ops.lib()
sms = torch.cuda.get_device_properties(0).multi_processor_count
for n in (64, 128):
    for mode, name in ((2049, "ordinary pairs"), (2048, "ws pairs (B kept)")):
        cyc = torch.zeros(sms, dtype=torch.int64, device="cuda")
        iters = 2000
        check(L.b200_debug_umma_rate(n, iters, mode, cyc.data_ptr(), sms, torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        print(f"N={n} {name:20s}: {cyc.float().mean().item() / (iters * 4):.2f} cycles/MMA (ideal {128 * n / 256:.0f})")
