"""Host side of the peer-memory exchange of the data-parallel step (``csrc/peer.cu`` holds the protocol).

One process per GPU on ONE node.  Every rank allocates its flat gradient buffer, its compute-dtype weight shadow and a
block of counters through ``b200_peer_alloc``, publishes the cudaIpc handles over the process group, and maps the
buffers of every peer.  Reduce-scatter and all-gather then are copy-engine pulls out of the peers' buffers plus one
summation kernel: no NCCL kernel shares the SMs with the convolutions of the step.

The reference trains on one GPU (``Super_resolution/code/train_adaptive_unet.py:622-632``); this is the exchange
of this framework's own batch-sharded step."""
import ctypes as C
from typing import Dict, List

import torch

from . import ops
from ._ffi import B200Error

N_SLOTS = 64            # counters per rank
ADAM_SLOT = N_SLOTS - 1  # "my optimizer step is done" (the gradient buckets use slots 0 .. N_SLOTS-2)
HANDLE_BYTES = 64


class _Blob:
    """cudaMalloc'ed memory behind the numba / cupy array protocol, so that torch can view it without a copy."""

    def __init__(self, ptr: int, count: int, typestr: str):
        self.ptr = ptr
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}


class PeerExchange:
    def __init__(self, dist, group, device: torch.device, elems: int, shadow_dtype: torch.dtype, timeout_s: float = 30.0):
        if shadow_dtype != torch.bfloat16:
            raise B200Error("peer exchange: the weight shadow must be bfloat16")
        self.dist, self.group, self.device = dist, group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if not 2 <= self.world <= 33:
            raise B200Error(f"peer exchange: world size {self.world} not supported")
        self.timeout_s = float(timeout_s)
        self.elems = elems
        lib = ops.lib()
        self._own: Dict[str, int] = {}
        self._mapped: Dict[int, Dict[str, int]] = {}
        mine, error = {}, None
        try:        # a local failure must not leave the other ranks alone in the collective below
            for name, nbytes in (("G", elems * 4), ("S", elems * 2), ("F", N_SLOTS * 8)):
                p = C.c_void_p()
                ops.check(lib.b200_peer_alloc(C.byref(p), nbytes), "peer_alloc")
                self._own[name] = p.value
                h = (C.c_ubyte * HANDLE_BYTES)()
                ops.check(lib.b200_peer_export(p, h), "peer_export")
                mine[name] = bytes(h)
        except B200Error as e:
            error = str(e)
        everyone: List[dict] = [None] * self.world
        dist.all_gather_object(everyone, {"device": device.index, "handles": mine, "error": error}, group=group)
        errors = [f"rank {r}: {o['error']}" for r, o in enumerate(everyone) if o["error"]]
        if errors:
            raise B200Error("peer exchange: " + "; ".join(errors))
        # torch views of my own buffers (the memory stays allocated for the life of the process)
        self._blobs = [_Blob(self._own["G"], elems, "<f4"), _Blob(self._own["S"], elems, "<i2")]
        self.G = torch.as_tensor(self._blobs[0], device=device)
        self.S = torch.as_tensor(self._blobs[1], device=device).view(torch.bfloat16)
        if self.G.data_ptr() != self._own["G"] or self.S.data_ptr() != self._own["S"]:
            raise B200Error("peer exchange: torch copied the exported buffers instead of viewing them")
        self.peers = [(self.rank + 1 + j) % self.world for j in range(self.world - 1)]    # staggered: no hot source
        self.peer_device = {p: everyone[p]["device"] for p in self.peers}
        if len({self.device.index, *self.peer_device.values()}) != self.world:
            raise B200Error("peer exchange: every rank needs its own GPU of the same node")
        for p in self.peers:
            self._mapped[p] = {}
            for name, hb in everyone[p]["handles"].items():
                out = C.c_void_p()
                ops.check(lib.b200_peer_open((C.c_ubyte * HANDLE_BYTES).from_buffer_copy(hb), C.byref(out)), "peer_open")
                self._mapped[p][name] = out.value
        n = len(self.peers)
        # device arrays for the wait kernel: slot -> the peers' addresses of that counter
        self._slot_ptrs = torch.tensor([[self._mapped[p]["F"] + 8 * s for p in self.peers] for s in range(N_SLOTS)],
                                       dtype=torch.int64, device=device)
        self._dev_arr = (C.c_int * n)(*[self.peer_device[p] for p in self.peers])
        self.stream = torch.cuda.Stream(device=device)
        self._staging = None
        self.ag_events: Dict[int, torch.cuda.Event] = {}
        self.ag_plan = None
        self.peers_past_adam = None        # event: every peer signalled its Adam (so nobody reads last step's G any more)
        torch.cuda.synchronize(device)
        # no barrier here: the caller agrees on success over the group (Model._enable_peer_exchange), which is one

    # ---- primitives (all on the CURRENT stream) ----------------------------------------------------------------
    def signal(self, slot: int):
        ops.check(ops.lib().b200_peer_signal(C.c_void_p(self._own["F"] + 8 * slot), ops._stream()), "peer_signal")

    def wait_peers(self, slot: int):
        ops.check(ops.lib().b200_peer_wait(C.c_void_p(self._slot_ptrs[slot].data_ptr()), len(self.peers),
                                           C.c_void_p(self._own["F"] + 8 * slot), self.timeout_s, ops._stream()), "peer_wait")

    def _stage(self, floats: int):
        if self._staging is None or self._staging.numel() < floats:
            self._staging = torch.empty(floats, dtype=torch.float32, device=self.device)
        return self._staging

    def reduce_scatter(self, slot: int, lo: int, hi: int):
        """My shard [lo, hi) of the bucket whose counter is `slot`: wait for the peers' buckets, pull their slices of my
        shard, add them into G[lo:hi]."""
        n, count = len(self.peers), hi - lo
        self.wait_peers(slot)
        stage = self._stage(n * count)
        src = (C.c_void_p * n)(*[self._mapped[p]["G"] + 4 * lo for p in self.peers])
        ops.check(ops.lib().b200_peer_gather_sum(C.c_void_p(self._own["G"] + 4 * lo), C.c_void_p(stage.data_ptr()), src,
                                                 self._dev_arr, n, count, self.device.index, ops._stream()),
                  "peer_gather_sum")

    def all_gather(self, shards):
        """shards[p] = (lo, hi) element range of the shadow that rank p owns in one bucket: pull every peer's shard."""
        n = len(self.peers)
        dst = (C.c_void_p * n)(*[self._own["S"] + 2 * shards[p][0] for p in self.peers])
        src = (C.c_void_p * n)(*[self._mapped[p]["S"] + 2 * shards[p][0] for p in self.peers])
        size = (C.c_size_t * n)(*[2 * (shards[p][1] - shards[p][0]) for p in self.peers])
        ops.check(ops.lib().b200_peer_pull(dst, src, self._dev_arr, size, n, self.device.index, ops._stream()), "peer_pull")

    def wait_gathers(self):
        """The current stream waits for every shadow pull still in flight."""
        cur = torch.cuda.current_stream()
        for ev in self.ag_events.values():
            cur.wait_event(ev)
        self.ag_events.clear()

    def close(self):
        """Stop using the peers' buffers.  The mappings are NOT unmapped here: cudaIpcCloseMemHandle needs the owning
        device to quiesce, and a rank that is already past this point may be spinning in a NCCL kernel that waits for the
        caller -- on 8 GPUs exactly that hung the teardown (profiles/r02_dp_check_n8_p2p.log: five correct steps, then the
        collective close never returned).  The driver unmaps everything when the process exits."""
        if not self._mapped:
            return
        torch.cuda.synchronize(self.device)
        self._mapped = {}
