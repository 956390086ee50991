"""Per-step host -> device staging.

Everything the train step draws on the HOST each iteration (numpy crop boxes, the colour-jitter order, CPU-generator
latents, Adam's step-dependent scalars) reaches the GPU through ``stage(producer, device)``:

* eager mode (default): ``producer()`` is copied into a ring of pinned buffers and sent with a non-blocking copy
  (a pageable ``.to(device)`` blocks the host until the stream drains - measured 2.4 ms/step);
* recording mode (``Recorder``, used by ``engine.GraphedTrainStep`` to turn the step into a CUDA graph): the call
  returns a STATIC device buffer and remembers the producer.  Before every replay the recorder calls the producers
  again in the recorded order - i.e. the host RNG streams advance exactly as in the eager loop - and refreshes the
  static buffers through pinned slots, outside the graph.
"""
import threading

import torch

_SLOTS = 8
_ACTIVE = None          # the Recorder that is currently recording, else None


class _PinnedRing(object):
    def __init__(self, slots=_SLOTS):
        self.slots, self.bufs, self.i = slots, {}, 0

    def next(self, shape, dtype):
        key = (tuple(shape), dtype)
        ring = self.bufs.get(key)
        if ring is None:
            ring = self.bufs[key] = [torch.empty(shape, dtype=dtype).pin_memory() for _ in range(self.slots)]
        self.i = (self.i + 1) % self.slots
        return ring[self.i]


class _ThreadLocalRing(threading.local):
    """One pinned ring per host thread: nn.DataParallel drives one replica per thread, and two replicas must never be
    handed the same pinned slot."""

    def __init__(self):
        self.ring = _PinnedRing()

    def next(self, shape, dtype):
        return self.ring.next(shape, dtype)


_RING = _ThreadLocalRing()


def recording():
    return _ACTIVE is not None


def stage(producer, device, shape=None, dtype=torch.float32, fill=False, late=False):
    """producer: zero-argument callable returning a CPU tensor (it may draw from host RNGs or advance host state).
    `shape` (optional) spares the recorder a throw-away call of the producer to learn the buffer size.
    fill=True (needs `shape`): the producer takes the pinned destination buffer and writes into it in place - no
    temporary, and no multi-threaded CPU copy (a 64k-element copy_ forks an OpenMP team: measured 1.6 ms on a busy
    128-core host).
    late=True: under a Recorder the producer must run at replay time, not ahead of it (it reads or advances host
    state that belongs to the step itself, e.g. Adam's step counters); it should be cheap."""
    device = torch.device(device)
    if _ACTIVE is not None:
        return _ACTIVE._record(producer, device, shape, dtype, fill, late)
    if device.type != "cuda":
        if fill:
            value = torch.empty(shape, dtype=dtype)
            producer(value)
            return value.to(device)
        return producer().to(device)
    if fill:
        buf = _RING.next(shape, dtype)
        producer(buf)
    else:
        value = producer()
        buf = _RING.next(value.shape, value.dtype)
        buf.copy_(value)
    return buf.to(device, non_blocking=True)


class _Entry(object):
    __slots__ = ("producer", "device_buf", "pinned", "events", "slot", "fill", "late")


class Recorder(object):
    """Collects the staged inputs of ONE step so that the step can be replayed as a CUDA graph.

    Two passes over the same step code:
      with rec.plan():     one ordinary EAGER step.  Every stage() call gets a static device buffer allocated from the
                           normal allocator (never from the graph's private pool: a buffer that is written from
                           outside the graph must not alias graph intermediates, which the pool would happily do once
                           an earlier intermediate is dead), filled for this step, and remembered.
      with rec.capture():  the stream capture.  The k-th stage() call returns the k-th planned buffer.
    Afterwards refresh() before each replay."""

    def __init__(self, slots=4):
        self.entries, self.slots, self.phase, self.cursor = [], slots, None, 0

    def plan(self):
        if self.entries:
            raise RuntimeError("staging.Recorder.plan(): already planned")
        self.phase = "plan"
        return self

    def capture(self):
        self.phase, self.cursor = "capture", 0
        return self

    def __enter__(self):
        global _ACTIVE
        if _ACTIVE is not None:
            raise RuntimeError("staging.Recorder is not re-entrant")
        if self.phase not in ("plan", "capture"):
            raise RuntimeError("use `with recorder.plan():` or `with recorder.capture():`")
        _ACTIVE = self
        return self

    def __exit__(self, exc_type, *exc):
        global _ACTIVE
        _ACTIVE = None
        if exc_type is None and self.phase == "capture" and self.cursor != len(self.entries):
            raise RuntimeError("staging: the captured step staged %d inputs, the planned step %d"
                               % (self.cursor, len(self.entries)))
        self.phase = None
        return False

    def _pinned(self, e):
        if e.pinned is None:
            host = torch.empty(e.device_buf.shape, dtype=e.device_buf.dtype)
            e.pinned = [host.clone().pin_memory() if e.device_buf.is_cuda else host.clone() for _ in range(self.slots)]
        return e.pinned

    def _produce(self, e):
        """Host half: draw this entry's next value into its next pinned slot."""
        e.slot = (e.slot + 1) % self.slots
        ev = e.events[e.slot]
        if ev is not None:
            ev.synchronize()                     # the upload that last used this pinned slot (`slots` steps ago)
        buf = self._pinned(e)[e.slot]
        if e.fill:
            e.producer(buf)
        else:
            buf.copy_(e.producer())

    def _upload(self, e):
        """Device half: enqueue the copy of the current pinned slot into the static buffer (current stream)."""
        e.device_buf.copy_(self._pinned(e)[e.slot], non_blocking=True)
        if e.device_buf.is_cuda:
            ev = e.events[e.slot]
            if ev is None:
                ev = e.events[e.slot] = torch.cuda.Event()
            ev.record()

    def _push(self, e):
        self._produce(e)
        self._upload(e)

    def _record(self, producer, device, shape, dtype, fill, late=False):
        if fill and shape is None:
            raise ValueError("staging.stage(fill=True) needs the buffer shape")
        if self.phase == "capture":
            if self.cursor >= len(self.entries):
                raise RuntimeError("staging: the captured step stages more inputs than the planned step")
            e = self.entries[self.cursor]
            self.cursor += 1
            if shape is not None and (tuple(shape) != tuple(e.device_buf.shape) or dtype != e.device_buf.dtype):
                raise RuntimeError("staging: input %d changed shape between the planned and the captured step" % (self.cursor - 1))
            e.producer, e.fill, e.late = producer, fill, late
            return e.device_buf
        e = _Entry()
        e.producer, e.fill, e.late = producer, fill, late
        e.pinned, e.events, e.slot = None, [None] * self.slots, 0
        if shape is None:
            first = producer()
            e.device_buf = torch.empty(first.shape, dtype=first.dtype, device=device)
            e.producer, e.fill = (lambda: first), False       # this step's value is already drawn
            self._push(e)
            e.producer, e.fill = producer, fill
        else:
            e.device_buf = torch.empty(tuple(shape), dtype=dtype, device=device)
            self._push(e)
        self.entries.append(e)
        return e.device_buf

    def produce(self):
        """Draw the next step's host values (recorded order) into pinned memory.  Host work only: call it right
        after launching a replay so that it overlaps the GPU."""
        for e in self.entries:
            if not e.late:
                self._produce(e)

    def upload(self):
        """Enqueue the copies of the produced values into the static device buffers on the current stream.
        Call right before ``graph.replay()``.  `late` entries are produced here."""
        for e in self.entries:
            if e.late:
                self._produce(e)
            self._upload(e)

    def refresh(self):
        """produce() + upload()."""
        for e in self.entries:
            self._push(e)
