"""Ingest throughput: JSON directories -> columnar batch.

    python tools/bench_ingest.py [--lnc 48] [--threads 0]

Thin wrapper: the measurement lives in bench.py (`python bench.py --stage ingest ...`), next to the other
legs that are allowed to run the CPU oracle / reference harness as a baseline."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.argv = [os.path.join(ROOT, "bench.py"), "--stage", "ingest"] + sys.argv[1:]
runpy.run_path(sys.argv[0], run_name="__main__")
