"""Two library contexts on one GPU fed from host batches: hand-rolled threads against stream.BatchStream
(tools script, GPU box): python tools/ab_stream.py [n_lnc] [steps]"""
import os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ortho2align_b200 import synth
from ortho2align_b200.engine import Engine
from ortho2align_b200.hostbatch import HostBatch
from ortho2align_b200.stream import BatchStream

n_lnc = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
batch = synth.workload(2, n_lnc=n_lnc, seed=0)
hb = HostBatch.from_columnar(batch, threads=os.cpu_count(), pinned=True)
e1, e2 = Engine(0), Engine(0)
p = e1.params("hist", True, 0.05, 5, 2, "block_length", "max")
s1, s2 = e1.hostbatch_struct(hb, False), e2.hostbatch_struct(hb, False)
for e, s in ((e1, s1), (e2, s2)):
    e.run(s, p); e.fetch(survivors=False, pinned=True)

def single():
    t0 = time.perf_counter()
    for _ in range(steps):
        e1.run(s1, p); e1.fetch(survivors=False, pinned=True)
    return (time.perf_counter() - t0) / steps * 1e3

def hand():
    def w(e, s):
        for _ in range(steps // 2):
            e.run(s, p); e.fetch(survivors=False, pinned=True)
    ths = [threading.Thread(target=w, args=a) for a in ((e1, s1), (e2, s2))]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    return (time.perf_counter() - t0) / steps * 1e3

def hand_build_host():
    def w(e):
        for _ in range(steps // 2):
            e.build_host(hb, "hist", True, 0.05, 5, 2, survivors=False, pinned=True)
    ths = [threading.Thread(target=w, args=(e,)) for e in (e1, e2)]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    return (time.perf_counter() - t0) / steps * 1e3

def stream():
    bs = BatchStream(engines=[e1, e2])
    t0 = time.perf_counter()
    out = list(bs.map((hb for _ in range(steps)), "hist", True, 0.05, 5, 2, survivors=False, pinned=True))
    return (time.perf_counter() - t0) / steps * 1e3

for name, f in (("single", single), ("hand", hand), ("hand_build_host", hand_build_host), ("stream", stream), ("hand", hand), ("stream", stream)):
    print(name, round(f(), 3), "ms per batch", flush=True)
