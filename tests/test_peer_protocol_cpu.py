"""The ordering protocol of the peer-memory exchange (``Model._run_step_p2p`` + ``peer.py`` + ``csrc/peer.cu``), checked
on CPU by randomised interleaving.

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
