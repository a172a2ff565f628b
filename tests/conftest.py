import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def batch_from_lists(d):
    from ortho2align_b200.columnar import ColumnarBatch
    i64 = lambda k: np.asarray(d[k], dtype=np.int64)
    i32 = lambda k: np.asarray(d[k], dtype=np.int32)
    return ColumnarBatch(bg_off=i64("bg_off"), bg=i32("bg"), aln_off=i64("aln_off"), hsp_off=i64("hsp_off"),
                         qlen=i64("qlen"), slen=i64("slen"), qstart=i32("qstart"), qend=i32("qend"),
                         sstart=i32("sstart"), send=i32("send"), score=np.asarray(d["score"], dtype=np.float64),
                         names=d.get("names")).validate()


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from ortho2align_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()
