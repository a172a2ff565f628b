"""A stream of host batches through several library contexts per GPU.

`o2a_build_batch` on a host batch is PCIe-bound, and a single context leaves the link idle while the last chunk of a
batch computes and its results are downloaded. `include/o2a.h` allows one host thread per `o2a_ctx`, so the remedy
is plain: `contexts` engines per device, one thread each, taking batches from a shared queue — the uploads of one
batch overlap the tail of another (measured on a B200 with BASELINE config-2 batches: 12.1 -> 11.2 ms per batch with
two contexts; `bench.py --e2e-engines 2` measures this class). Results come back in input order.

The reference has no counterpart (its unit of parallelism is one process per lncRNA, parallel.py:296-488); this is
the host-side scheduling a caller with many batches (e.g. chromosome-sized groups of lncRNAs) would otherwise write.
"""
from __future__ import annotations

import queue
import threading


def _default_factory(device):
    from .engine import Engine
    return Engine(device)


class BatchStream:
    """`contexts` engines on each of `devices`; `map()` feeds them and yields `BatchResult`s in order.
    `engines`: ready `Engine` objects to use instead (one worker thread each; they are left open)."""

    def __init__(self, devices=(0,), contexts=2, engine_factory=None, engines=None):
        if engines is not None:
            if not engines:
                raise ValueError("need at least one engine")
            self.engines = list(engines)
            self.slots = list(range(len(self.engines)))
            self.factory = lambda k: self.engines[k]
            return
        if contexts < 1 or not devices:
            raise ValueError("need at least one device and one context")
        self.engines = None
        self.slots = [d for d in devices for _ in range(contexts)]
        self.factory = engine_factory or _default_factory

    def map(self, batches, fitter="hist", fdr=True, alpha=0.05, gapopen=5, gapextend=2, **build_kwargs):
        """batches: iterable of host batches — `hostbatch.HostBatch` (the ingest's pinned narrow layout, passed as is)
        or `ColumnarBatch`. Yields results in the order of `batches`; at most 2 x len(slots) batches are in flight or
        waiting to be collected. An engine failure is re-raised here."""
        from .hostbatch import HostBatch
        work = queue.Queue(maxsize=len(self.slots))
        done = {}
        cond = threading.Condition()
        errors = []

        def worker(device):
            eng = None
            try:
                eng = self.factory(device)
                while True:
                    item = work.get()                     # None = no more work (one per worker, see the end of map)
                    if item is None:
                        break
                    i, batch = item
                    run = eng.build_host if isinstance(batch, HostBatch) else eng.build_batch
                    res = run(batch, fitter, fdr, alpha, gapopen, gapextend, **build_kwargs)
                    with cond:
                        done[i] = res
                        cond.notify_all()
            except Exception as e:                       # surfaced on the caller's thread
                with cond:
                    errors.append(e)
                    cond.notify_all()
            finally:
                if eng is not None and self.engines is None:
                    eng.close()

        threads = [threading.Thread(target=worker, args=(d,), daemon=True) for d in self.slots]
        for t in threads:
            t.start()

        def put(item):
            while True:                                   # a dead worker must not block the feeder for ever
                if errors:
                    raise errors[0]
                try:
                    work.put(item, timeout=0.05)
                    return
                except queue.Full:
                    continue

        try:
            fed = collected = 0
            it = iter(batches)
            exhausted = False
            while not exhausted or collected < fed:
                while not exhausted and fed - collected < 2 * len(self.slots):
                    try:
                        batch = next(it)
                    except StopIteration:
                        exhausted = True
                        break
                    put((fed, batch))
                    fed += 1
                if collected < fed:
                    with cond:
                        while collected not in done and not errors:
                            cond.wait(0.05)
                        if errors:
                            raise errors[0]
                        res = done.pop(collected)
                    collected += 1
                    yield res
        finally:
            # wake every worker at once (a polling get() here cost up to 50 ms per map() call)
            for _ in threads:
                try:
                    work.put(None, timeout=5)
                except queue.Full:
                    pass
            for t in threads:
                t.join(timeout=5)
