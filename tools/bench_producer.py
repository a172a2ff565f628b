"""Producer-side throughput: BLAST tables -> to_genomic + cut_coordinates -> columnar.

    python tools/bench_producer.py [--lnc 10000] [--steps 10] [--warmup 3] [--cpu-lnc 2000] [--tables 200]

Thin wrapper: the measurement lives in bench.py (`python bench.py --stage producer ...`), next to the other
legs that are allowed to run the CPU oracle / reference harness as a baseline."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.argv = [os.path.join(ROOT, "bench.py"), "--stage", "producer"] + sys.argv[1:]
runpy.run_path(sys.argv[0], run_name="__main__")
