"""annotate_orthologs' interval join: orthologs x annotation -> neighbours + JI/OC.

    python tools/bench_annotate.py [--orthologs 4000000] [--annotation 4000000] [--steps 10] [--warmup 3]

Thin wrapper: the measurement lives in bench.py (`python bench.py --stage annotate ...`), next to the other
legs that are allowed to run the CPU oracle as a baseline."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.argv = [os.path.join(ROOT, "bench.py"), "--stage", "annotate"] + sys.argv[1:]
runpy.run_path(sys.argv[0], run_name="__main__")
