"""The ordering of the data-parallel step's exchange, checked on CPU by randomised interleaving: the peer-memory protocol
(``Model._run_step_p2p`` + ``peer.py`` + ``csrc/peer.cu``) and, further down, the default NCCL schedule (``Model._run_step``:
reduce-scatter per bucket behind its backward segment, shadow all-gathers overlapping the next forward pass).

The REAL host code of the step (``Model._run_step_p2p``) is run for every rank against stand-ins for the CUDA streams,
events, captured graphs and the PeerExchange; everything it issues lands in per-rank, per-stream FIFO queues.  A scheduler
then executes the queues of all ranks in random interleavings that respect CUDA's semantics (stream order, event waits,
the counter waits of ``peer_wait_kernel``), with the hosts running maximally ahead (all steps enqueued before anything
executes).  Every read is version-checked:

  * a gradient pull must see the peer's gradient of THIS step (not a zeroed buffer, not the previous or the next step's);
  * Adam must see the exact sum over the ranks;
  * a shadow pull must see the peer's Adam of this step; forward and backward must read the shadow of the previous Adam
    in every shard of every bucket.

No interleaving may deadlock.  The checker has teeth: without the "every peer is past its Adam" guard in front of the
gradient buffer's reset, some interleaving makes a slow rank pull zeros (second test)."""
import contextlib
import random
import types

import pytest
import torch

from b200unet.keras.model import Model
from b200unet.parallel import shard_of
from b200unet.peer import ADAM_SLOT, N_SLOTS


class Violation(AssertionError):
    pass


class Event:
    def __init__(self):
        self.done = False


class Stream:
    def __init__(self, sim, rank, name):
        self.sim, self.rank, self.name, self.q = sim, rank, name, []

    def push(self, runnable, run, what):
        self.q.append((runnable, run, what))

    def record_event(self):
        ev = Event()
        self.push(lambda: True, lambda: setattr(ev, "done", True), "record")
        return ev

    def wait_event(self, ev):
        self.push(lambda: ev.done, lambda: None, "wait_event")

    def wait_stream(self, other):
        self.wait_event(other.record_event())


class Graph:
    def __init__(self, rank_ctx, run, what):
        self.ctx, self.run, self.what = rank_ctx, run, what

    def replay(self):
        step = self.ctx.step
        self.ctx.sim.current().push(lambda: True, lambda: self.run(step), self.what)


class FakePeer:
    """PeerExchange stand-in: the same calls, turned into queue entries with the semantics of csrc/peer.cu."""

    def __init__(self, ctx):
        self.ctx, self.rank, self.world = ctx, ctx.rank, ctx.sim.world
        self.stream = ctx.cs
        self.ag_events, self.ag_plan, self.peers_past_adam = {}, None, None
        self.guard = True

    def __setattr__(self, k, v):
        if k == "peers_past_adam" and not getattr(self, "guard", True):
            v = None                       # mutation: drop the guard in front of the gradient buffer's reset
        object.__setattr__(self, k, v)

    def signal(self, slot):
        F = self.ctx.sim.F
        r = self.rank
        self.ctx.sim.current().push(lambda: True, lambda: F[r].__setitem__(slot, F[r][slot] + 1), f"signal {slot}")

    def wait_peers(self, slot):
        F, r, world = self.ctx.sim.F, self.rank, self.world
        self.ctx.sim.current().push(lambda: all(F[p][slot] >= F[r][slot] for p in range(world) if p != r), lambda: None,
                                    f"wait_peers {slot}")

    def reduce_scatter(self, slot, lo, hi):
        sim, r, k, step = self.ctx.sim, self.rank, slot, self.ctx.step
        assert (lo, hi) == shard_of(sim.buckets[k], r, self.world)
        self.wait_peers(slot)
        staged = {}

        def pull(p):
            tag, val = sim.G[p][k][r]
            if tag != step:
                raise Violation(f"rank {r} step {step}: gradient of bucket {k} pulled from rank {p} has tag {tag}")
            staged[p] = val

        for p in [(r + 1 + j) % self.world for j in range(self.world - 1)]:
            sim.current().push(lambda: True, lambda p=p: pull(p), f"pull G[{p}][{k}]")

        def total():
            tag, own = sim.G[r][k][r]
            if tag != step:
                raise Violation(f"rank {r} step {step}: own gradient of bucket {k} has tag {tag}")
            sim.G[r][k][r] = (("sum", step), own + sum(staged.values()))

        sim.current().push(lambda: True, total, f"sum {k}")

    def all_gather(self, shards):
        sim, r, step = self.ctx.sim, self.rank, self.ctx.step
        k = [i for i, b in enumerate(sim.buckets) if b["sharded"] and shard_of(b, 0, self.world) == shards[0]][0]

        def pull(p):
            ver = sim.S[p][k][p]
            if ver != step:
                raise Violation(f"rank {r} step {step}: shadow shard of rank {p}, bucket {k} has version {ver}")
            sim.S[r][k][p] = ver

        for p in [(r + 1 + j) % self.world for j in range(self.world - 1)]:
            sim.current().push(lambda: True, lambda p=p: pull(p), f"pull S[{p}][{k}]")

    def wait_gathers(self):
        cur = self.ctx.sim.current()
        for ev in self.ag_events.values():
            cur.wait_event(ev)
        self.ag_events.clear()


class Sim:
    def __init__(self, world, n_buckets, steps, seed, guard=True):
        self.world, self.steps, self.rng = world, steps, random.Random(seed)
        per = 64 * world
        # buckets as Model._buckets returns them: ordered by readiness, the LAST parameters first; one replicated bucket
        self.buckets = [{"lo": (n_buckets - 1 - k) * per, "hi": (n_buckets - k) * per, "ready_after": k, "sharded": True}
                        for k in range(n_buckets)]
        self.buckets.append({"lo": n_buckets * per, "hi": n_buckets * per + 8, "ready_after": n_buckets - 1, "sharded": False})
        self.F = [[0] * N_SLOTS for _ in range(world)]
        self.G = [[[(("z", -1), 0.0)] * world for _ in range(n_buckets)] for _ in range(world)]
        self.S = [[[-1] * world for _ in range(n_buckets)] for _ in range(world)]
        self.ranks = [self._rank(r, guard) for r in range(world)]
        self._cur = None

    def grad(self, r, k, s, step):
        return (step + 1) * 1000.0 + r * 100 + k * 10 + s

    def current(self):
        return self._cur

    def _rank(self, r, guard):
        sim, nb, world = self, len(self.buckets) - 1, self.world
        ctx = types.SimpleNamespace(sim=self, rank=r, step=-1)
        ctx.main, ctx.cs = Stream(self, r, "main"), Stream(self, r, "exchange")
        ctx.peer = FakePeer(ctx)
        ctx.peer.guard = guard

        def check_shadow(step, which, ks):
            for k in ks:
                for s in range(world):
                    if sim.S[r][k][s] != step - 1:
                        raise Violation(f"rank {r} step {step}: {which} reads shadow bucket {k} shard {s} at version "
                                        f"{sim.S[r][k][s]}")

        def fwd(i, step):
            if i == 0:       # the first forward segment zeroes the gradient buffer and reads only the replicated head
                for k in range(nb):
                    sim.G[r][k] = [(("z", step), 0.0)] * world
            else:
                check_shadow(step, f"forward segment {i}", [nb - i])     # bucket nb-1 holds the first sharded layers

        def bwd(j, step):
            check_shadow(step, f"backward segment {j}", range(nb))       # dgrad reads every weight
            for s in range(world):
                tag, _ = sim.G[r][j][s]
                if tag != ("z", step):
                    raise Violation(f"rank {r} step {step}: bucket {j} not freshly zeroed (tag {tag})")
                sim.G[r][j][s] = (step, sim.grad(r, j, s, step))

        def adam(step):
            for k in range(nb):
                tag, val = sim.G[r][k][r]
                want = sum(sim.grad(p, k, r, step) for p in range(world))
                if tag != ("sum", step) or val != want:
                    raise Violation(f"rank {r} step {step}: Adam sees bucket {k} tag {tag} value {val}, wanted {want}")
                sim.S[r][k][r] = step

        plan = object()
        entry = {"plan": plan,
                 "fsegs": [(Graph(ctx, lambda step, i=i: fwd(i, step), f"fwd {i}"), [nb - i] if i else [])
                           for i in range(nb + 1)],
                 "segments": [(Graph(ctx, lambda step, j=j: bwd(j, step), f"bwd {j}"),
                               [self.buckets[j]] + ([self.buckets[nb]] if j == nb - 1 else [])) for j in range(nb)],
                 "adam_graph": Graph(ctx, adam, "adam"),
                 "ag_order": [nb - 1 - i for i in range(nb)]}
        model = types.SimpleNamespace(_peer=ctx.peer, _ag_works=None, S=object(), P=object(),
                                      _buckets=lambda _plan: sim.buckets, _reduce_async=lambda bs: [],
                                      _check_finite=lambda: None)
        ctx.model, ctx.entry = model, entry
        return ctx

    def enqueue_all(self, monkeypatch):
        @contextlib.contextmanager
        def use(stream):
            prev, self._cur = self._cur, stream
            try:
                yield
            finally:
                self._cur = prev

        monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: self._cur)
        monkeypatch.setattr(torch.cuda, "stream", use)
        for ctx in self.ranks:                       # the hosts run maximally ahead: every step is enqueued up front
            for step in range(self.steps):
                ctx.step = step
                with use(ctx.main):
                    Model._run_step_p2p(ctx.model, ctx.entry)

    def run(self):
        queues = [s for ctx in self.ranks for s in (ctx.main, ctx.cs)]
        favourite = None
        while any(s.q for s in queues):
            ready = [s for s in queues if s.q and s.q[0][0]()]
            if not ready:
                heads = [(s.rank, s.name, s.q[0][2]) for s in queues if s.q]
                raise Violation(f"deadlock: {heads}")
            # bursts: let one rank race ahead for a while (the write-after-read hazards need skew between the ranks)
            if favourite is None or self.rng.random() < 0.05:
                favourite = self.rng.randrange(self.world) if self.rng.random() < 0.7 else -1
            mine = [s for s in ready if s.rank == favourite]
            s = self.rng.choice(mine or ready)
            _, run, _ = s.q.pop(0)
            run()
        for r in range(self.world):                  # after the last step every rank holds every shard of the last Adam
            for k in range(len(self.buckets) - 1):
                assert self.S[r][k] == [self.steps - 1] * self.world, (r, k, self.S[r][k])


@pytest.mark.parametrize("world,n_buckets", [(2, 1), (2, 3), (3, 2), (4, 3), (8, 4)])
def test_peer_exchange_ordering_is_safe_under_random_interleavings(monkeypatch, world, n_buckets):
    for seed in range(40 if world < 8 else 10):
        sim = Sim(world, n_buckets, steps=4, seed=seed)
        sim.enqueue_all(monkeypatch)
        sim.run()
        assert all(f[ADAM_SLOT] == 4 for f in sim.F)
        assert all(f[k] == 4 for f in sim.F for k in range(n_buckets))


def test_peer_exchange_checker_detects_a_missing_guard(monkeypatch):
    """Without the wait for "every peer is past its Adam" in front of the gradient buffer's reset, a fast rank zeroes its
    buffer while a slow peer still pulls last step's gradient out of it: some interleaving must trip the version check."""
    caught = 0
    for seed in range(60):
        sim = Sim(3, 2, steps=4, seed=seed, guard=False)
        sim.enqueue_all(monkeypatch)
        try:
            sim.run()
        except Violation:
            caught += 1
    assert caught > 0


# ---- the same check for the default (NCCL) exchange: Model._run_step's segmented branch ---------------------------------
class Work:
    def __init__(self, sim, ev):
        self.sim, self.ev = sim, ev

    def wait(self):
        self.sim.current().wait_event(self.ev)


class Flat:
    """Stands for the flat G / S tensor: slicing yields (buffer, lo, hi)."""

    def __init__(self, name):
        self.name = name

    def __getitem__(self, sl):
        return (self.name, sl.start, sl.stop)


class FakeDist:
    """torch.distributed stand-in with NCCL's stream semantics: the collective is ordered behind everything issued before
    it on the calling stream, runs on the rank's communication stream, and completes only once every rank has joined it.
    A rank's buffer is read between its join and its completion: it must not change in that window (lock)."""

    class ReduceOp:
        SUM, MAX = "sum", "max"

    def __init__(self, ctx):
        self.ctx = ctx

    def get_rank(self, group=None):
        return self.ctx.rank

    def get_world_size(self, group=None):
        return self.ctx.sim.world

    def get_backend(self, group=None):
        return "nccl"

    def _collective(self, join, finish, what):
        sim, ctx = self.ctx.sim, self.ctx
        n = ctx.ncoll
        ctx.ncoll += 1
        state = sim.coll.setdefault(n, {"joined": set(), "data": {}})
        ctx.nccl.wait_event(sim.current().record_event())
        ctx.nccl.push(lambda: True, lambda: (join(state), state["joined"].add(ctx.rank)), f"join {what}")
        ctx.nccl.push(lambda: len(state["joined"]) == sim.world, lambda: finish(state), f"finish {what}")
        return Work(sim, ctx.nccl.record_event())

    def _bucket(self, lo, hi):
        return [k for k, b in enumerate(self.ctx.sim.buckets) if (b["lo"], b["hi"]) == (lo, hi)][0]

    def reduce_scatter_tensor(self, out, inp, op=None, group=None, async_op=False):
        sim, r, step = self.ctx.sim, self.ctx.rank, self.ctx.step
        k = self._bucket(inp[1], inp[2])
        assert inp[0] == "G" and (out[1], out[2]) == shard_of(sim.buckets[k], r, sim.world)

        def join(state):
            for s in range(sim.world):
                if sim.G[r][k][s][0] != step:
                    raise Violation(f"rank {r} step {step}: reduce-scatter of bucket {k} joins with tag {sim.G[r][k][s][0]}")
            state["data"][r] = [v for _, v in sim.G[r][k]]
            sim.locks.add(("G", r, k))

        def finish(state):
            sim.G[r][k][r] = (("sum", step), sum(state["data"][p][r] for p in range(sim.world)))
            sim.locks.discard(("G", r, k))

        return self._collective(join, finish, f"reduce_scatter {k}")

    def all_gather_into_tensor(self, out, inp, group=None, async_op=False):
        sim, r, step = self.ctx.sim, self.ctx.rank, self.ctx.step
        k = self._bucket(out[1], out[2])
        assert out[0] == "S" and (inp[1], inp[2]) == shard_of(sim.buckets[k], r, sim.world)

        def join(state):
            if sim.S[r][k][r] != step:
                raise Violation(f"rank {r} step {step}: all-gather of bucket {k} joins with own version {sim.S[r][k][r]}")
            state["data"][r] = step
            sim.locks.add(("S", r, k))
            for s in range(sim.world):
                if s != r:
                    sim.S[r][k][s] = "in flight"

        def finish(state):
            for s in range(sim.world):
                sim.S[r][k][s] = state["data"][s]
            sim.locks.discard(("S", r, k))

        return self._collective(join, finish, f"all_gather {k}")

    def all_reduce(self, t, op=None, group=None, async_op=False):
        return self._collective(lambda state: None, lambda state: None, "all_reduce")


class NcclSim(Sim):
    def __init__(self, *a, forward_waits=True, **kw):
        self.coll, self.locks, self.forward_waits = {}, set(), forward_waits
        super().__init__(*a, **kw)

    def _rank(self, r, guard):
        sim, nb, world = self, len(self.buckets) - 1, self.world
        ctx = types.SimpleNamespace(sim=self, rank=r, step=-1, ncoll=0)
        ctx.main, ctx.cs, ctx.nccl = Stream(self, r, "main"), Stream(self, r, "unused"), Stream(self, r, "nccl")

        def check_shadow(step, which, ks):
            for k in ks:
                for s in range(world):
                    if sim.S[r][k][s] != step - 1:
                        raise Violation(f"rank {r} step {step}: {which} reads shadow bucket {k} shard {s}: {sim.S[r][k][s]}")

        def fwd(i, step):
            if i == 0:
                for k in range(nb):
                    if ("G", r, k) in sim.locks:
                        raise Violation(f"rank {r} step {step}: gradient bucket {k} zeroed under a running reduce-scatter")
                    sim.G[r][k] = [(("z", step), 0.0)] * world
            else:
                check_shadow(step, f"forward segment {i}", [nb - i])

        def bwd(j, step):
            check_shadow(step, f"backward segment {j}", range(nb))
            for s in range(world):
                if sim.G[r][j][s][0] != ("z", step) or ("G", r, j) in sim.locks:
                    raise Violation(f"rank {r} step {step}: bucket {j} not freshly zeroed / in use ({sim.G[r][j][s][0]})")
                sim.G[r][j][s] = (step, sim.grad(r, j, s, step))

        def adam(step):
            for k in range(nb):
                tag, val = sim.G[r][k][r]
                want = sum(sim.grad(p, k, r, step) for p in range(world))
                if tag != ("sum", step) or val != want or ("S", r, k) in sim.locks:
                    raise Violation(f"rank {r} step {step}: Adam sees bucket {k} tag {tag} value {val}, wanted {want}")
                sim.S[r][k][r] = step

        plan = object()
        entry = {"plan": plan, "graph": None,
                 "fsegs": [(Graph(ctx, lambda step, i=i: fwd(i, step), f"fwd {i}"),
                            [nb - i] if i and sim.forward_waits else []) for i in range(nb + 1)],
                 "segments": [(Graph(ctx, lambda step, j=j: bwd(j, step), f"bwd {j}"),
                               [self.buckets[j]] + ([self.buckets[nb]] if j == nb - 1 else [])) for j in range(nb)],
                 "adam_graph": Graph(ctx, adam, "adam"),
                 "ag_order": [nb - i for i in range(1, nb + 1)]}
        model = types.SimpleNamespace(_peer=None, _ag_works=None, G=Flat("G"), S=Flat("S"), P=object(),
                                      _dist=(FakeDist(ctx), None), _buckets=lambda _plan: sim.buckets,
                                      _sharded=lambda: True, _check_finite=lambda: None)
        for name in ("_run_step", "_finish_gather", "_gather_updated", "_reduce_async"):
            setattr(model, name, types.MethodType(getattr(Model, name), model))
        ctx.model, ctx.entry = model, entry
        ctx.run_step = lambda: model._run_step(entry)
        return ctx

    def enqueue_all(self, monkeypatch):
        monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: self._cur)
        for ctx in self.ranks:
            for step in range(self.steps):
                ctx.step = step
                self._cur = ctx.main
                ctx.run_step()
        self._cur = None

    def run(self):
        for ctx in self.ranks:
            ctx.cs = ctx.nccl                        # the scheduler walks (main, cs): make the second one NCCL's stream
        super().run()


@pytest.mark.parametrize("world,n_buckets", [(2, 1), (2, 3), (4, 3), (8, 4)])
def test_nccl_exchange_ordering_is_safe_under_random_interleavings(monkeypatch, world, n_buckets):
    monkeypatch.delenv("B200_DP_OVERLAP_GATHER", raising=False)
    for seed in range(30 if world < 8 else 8):
        sim = NcclSim(world, n_buckets, steps=4, seed=seed)
        sim.enqueue_all(monkeypatch)
        sim.run()


def test_nccl_exchange_checker_detects_a_forward_pass_that_does_not_wait(monkeypatch):
    """If the forward segments did not wait for the all-gather of the bucket they read, some interleaving reads a shard
    that is still in flight."""
    monkeypatch.delenv("B200_DP_OVERLAP_GATHER", raising=False)
    caught = 0
    for seed in range(40):
        sim = NcclSim(3, 2, steps=4, seed=seed, forward_waits=False)
        sim.enqueue_all(monkeypatch)
        try:
            sim.run()
        except Violation:
            caught += 1
    assert caught > 0
