"""N>1 path on CPU: world_size-2 gloo. Each rank takes its LPT shard of the same seeded batch, runs
the per-shard compute (oracle injected — no GPU here), and the gathered result must equal the
single-process result on the whole batch."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from ortho2align_b200 import synth
    from ortho2align_b200.distributed import gather_best_orthologs, gather_counts, gather_rows, my_shard
    batch = synth.workload(5, n_lnc=120, seed=3)
    idx, sub = my_shard(batch, rank, world)
    res = O.build_batch(sub, "hist", True)
    dev = torch.device("cpu")
    counts = gather_counts([int(res.n_surv.sum()), int(res.n_keep.sum()), int(res.chain_len.sum()),
                            int((res.chain_len > 0).sum()), int((res.status != 0).sum())], dev)
    has = gather_best_orthologs(res, idx, batch.n_lnc, dev)
    thr_rows = gather_rows(res.thr, dev)
    idx_rows = gather_rows(np.asarray(idx, dtype=np.int64), dev)
    if rank == 0:
        thr = np.zeros(batch.n_lnc)
        for i, t in zip(idx_rows, thr_rows):
            thr[i] = t
        q.put((counts.sum(axis=0).tolist(), has.tolist(), thr.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_shards_reproduce_the_single_process_result():
    from oracle import oracle as O
    from ortho2align_b200 import synth
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    counts, has, thr = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    batch = synth.workload(5, n_lnc=120, seed=3)
    full = O.build_batch(batch, "hist", True)
    assert counts == [int(full.n_surv.sum()), int(full.n_keep.sum()), int(full.chain_len.sum()),
                      int((full.chain_len > 0).sum()), int((full.status != 0).sum())]
    assert has == (full.best_aln >= 0).astype(int).tolist()
    assert thr == full.thr.tolist()


def test_take_lnc_preserves_per_lncrna_results():
    from oracle import oracle as O
    from ortho2align_b200 import synth
    batch = synth.workload(5, n_lnc=40, seed=9)
    full = O.build_batch(batch, "hist", True)
    idx = [31, 2, 17, 5]
    sub = O.build_batch(batch.take_lnc(idx), "hist", True)
    assert sub.thr.tolist() == full.thr[idx].tolist()
    for k, l in enumerate(idx):
        fa = range(batch.aln_off[l], batch.aln_off[l + 1])
        sa0 = sum(int(batch.aln_off[i + 1] - batch.aln_off[i]) for i in idx[:k])
        for j, a in enumerate(fa):
            f = full.chain_idx[full.chain_off[a]:full.chain_off[a + 1]]
            s = sub.chain_idx[sub.chain_off[sa0 + j]:sub.chain_off[sa0 + j + 1]]
            assert f.tolist() == s.tolist()
