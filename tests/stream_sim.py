"""A model of CUDA stream / event ordering for the engine's host orchestration -- TEST INFRASTRUCTURE.

tests/cabi_emulator.py executes every emulated kernel at the moment it is "launched", i.e. in host program order.  That checks
what the host enqueues, not whether it ORDERED it: on the GPU the work of different streams runs concurrently and only the
event edges the host recorded constrain it.  This module closes that gap without a GPU: launches, device copies and fills are
queued per stream; `Event.record` / `Stream.wait_event` / `wait_stream` become markers and waits with CUDA's semantics (a wait
binds to the latest record enqueued before it; waiting on a never-recorded event is a no-op; created streams do not
synchronise with the default stream); host synchronisation points run exactly what they have to.  A scheduling policy then
picks, among all executions CUDA would allow, adversarial ones: everything as late as possible, some streams as early as
possible and the others as late as possible, or random interleavings.  If the host forgot an edge, some legal schedule runs a
consumer before its producer (or lets a producer overwrite a buffer still being read) and the run no longer reproduces the
reference golden (tests/test_stream_order_cpu.py).  A wait cycle shows up as a RecursionError / "deadlock".
"""
import contextlib
import ctypes as C
import random

import torch

import cabi_emulator


class _Marker(object):
    __slots__ = ("stream", "done")

    def __init__(self, stream):
        self.stream, self.done = stream, False


class VStream(object):
    def __init__(self, sim, name):
        self.sim, self.name = sim, name
        self.items, self.pos = [], 0
        self.cuda_stream = len(sim.streams)          # the "handle" the C ABI receives
        sim.streams.append(self)

    # torch.cuda.Stream API used by the engine
    def wait_event(self, ev):
        if ev.marker is not None:                    # never recorded: CUDA treats the wait as satisfied
            self.items.append(("wait", ev.marker))
        self.sim.pump()

    def wait_stream(self, other):
        m = _Marker(other)
        other.items.append(("record", m))
        self.items.append(("wait", m))
        self.sim.pump()

    def synchronize(self):
        m = _Marker(self)
        self.items.append(("record", m))
        self.sim.complete(m)

    def runnable(self):
        if self.pos >= len(self.items):
            return False
        kind, x = self.items[self.pos]
        return kind != "wait" or x.done

    def step(self):
        kind, x = self.items[self.pos]
        self.pos += 1
        if kind == "op":
            self.sim.host_mode = False
            try:
                x()
            finally:
                self.sim.host_mode = True
            self.sim.executed += 1
        elif kind == "record":
            x.done = True


class VEvent(object):
    def __init__(self, *a, **k):
        self.marker = None

    def record(self, stream=None):
        s = stream if stream is not None else SIM.current()
        self.marker = _Marker(s)
        s.items.append(("record", self.marker))
        SIM.pump()

    def wait(self, stream=None):
        (stream if stream is not None else SIM.current()).wait_event(self)

    def synchronize(self):
        if self.marker is not None:
            SIM.complete(self.marker)

    def elapsed_time(self, other):
        return 1.0


class Sim(object):
    """policy: "lazy" | "eager" | "workers_eager" | "default_eager" | ("random", seed)."""

    def __init__(self, policy):
        self.policy = policy
        self.rng = random.Random(policy[1]) if isinstance(policy, tuple) else None
        self.streams = []
        self.default = VStream(self, "default")
        self.stack = [self.default]
        self.host_mode = True       # False while a queued operation executes (its own tensor ops run directly)
        self.active = False         # interception of tensor ops only while the model runs
        self.executed = 0
        self.keepalive = []

    def current(self):
        return self.stack[-1]

    def by_handle(self, st):
        v = st.value if isinstance(st, C.c_void_p) else st
        return self.streams[int(v or 0)]

    def enqueue(self, fn, stream=None):
        (stream or self.current()).items.append(("op", fn))
        self.pump()

    def _is_eager(self, s):
        p = self.policy
        return p == "eager" or (p == "workers_eager" and s is not self.default) or (p == "default_eager" and s is self.default)

    def pump(self):
        """Let the streams the policy runs ahead make progress (called after every host-side enqueue)."""
        if self.rng is not None:
            for _ in range(self.rng.randint(0, 3)):
                cand = [s for s in self.streams if s.runnable()]
                if not cand:
                    return
                self.rng.choice(cand).step()
            return
        progress = True
        while progress:
            progress = False
            for s in self.streams:
                while self._is_eager(s) and s.runnable():
                    s.step()
                    progress = True

    def complete(self, marker, depth=0):
        """A host synchronisation point: run what `marker` depends on, and nothing else."""
        if depth > 200:
            raise RuntimeError("deadlock: cyclic stream / event wait")
        s = marker.stream
        while not marker.done:
            kind, x = s.items[s.pos]
            if kind == "wait" and not x.done:
                self.complete(x, depth + 1)      # ends with a pump that may also advance s: re-read its head
                continue
            s.step()
        self.pump()

    def sync_all(self):
        for s in list(self.streams):
            s.synchronize()


SIM = None


class _SimLib(object):
    """cabi_emulator.FakeLib behind per-stream queues: an entry point with a stream argument only enqueues."""
    HOST = ("smot_abi_version", "smot_last_error", "smot_rpn_select_workspace", "smot_sort_nms_workspace", "smot_conv2d_algo")

    def __init__(self, fake):
        self._fake = fake

    def __getattr__(self, name):
        target = getattr(self._fake, name)
        if name in self.HOST:
            return target

        def launch(*args):
            SIM.enqueue(lambda: target(*args), SIM.by_handle(args[-1]))
            return 0
        launch.__name__ = name
        return launch


def install(monkeypatch, policy):
    """cabi_emulator.install + stream semantics.  Returns (sim, fake).  Set sim.active = True around model calls."""
    global SIM
    from siammot_b200 import _lib, engine, ops, preprocess
    fake = cabi_emulator.install(monkeypatch)
    SIM = sim = Sim(policy)
    simlib = _SimLib(fake)
    for mod in (_lib, engine, ops, preprocess):
        monkeypatch.setattr(mod, "lib", lambda: simlib)
    for mod in (_lib, ops):
        monkeypatch.setattr(mod, "stream_ptr", lambda: C.c_void_p(sim.current().cuda_stream))
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: sim.current())
    monkeypatch.setattr(torch.cuda, "Event", VEvent)
    monkeypatch.setattr(torch.cuda, "Stream", lambda *a, **k: VStream(sim, "s%d" % len(sim.streams)))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: sim.sync_all())

    @contextlib.contextmanager
    def stream_ctx(s):
        sim.stack.append(s)
        try:
            yield
        finally:
            sim.stack.pop()
    monkeypatch.setattr(torch.cuda, "stream", stream_ctx)

    # device-side tensor operations issued from host code: queued on the current stream
    def queued(method):
        orig = getattr(torch.Tensor, method)

        def wrapper(self, *a, **k):
            if sim.active and sim.host_mode:
                sim.enqueue(lambda: orig(self, *a, **k))
                return self
            return orig(self, *a, **k)
        return wrapper
    for method in ("copy_", "fill_", "zero_"):
        monkeypatch.setattr(torch.Tensor, method, queued(method))
    # The CUDA caching allocator hands a freed block only to later work of the same stream, so a temporary may die on the
    # host while kernels that use it are still queued.  Host memory has no such guarantee: keep every buffer allocated while
    # the model runs alive until the simulation has drained.
    def keeping(name):
        orig = getattr(torch, name)

        def wrapper(*a, **k):
            t = orig(*a, **k)
            if sim.active and sim.host_mode:
                sim.keepalive.append(t)
            return t
        return wrapper
    for name in ("zeros", "empty", "full"):
        monkeypatch.setattr(torch, name, keeping(name))
    orig_gather = engine.Engine.gather_templates

    def gather(self, feat, first, sources):
        if sim.active and sim.host_mode:
            sim.enqueue(lambda: orig_gather(self, feat, first, sources))
        else:
            orig_gather(self, feat, first, sources)
    monkeypatch.setattr(engine.Engine, "gather_templates", gather)
    return sim, fake
