"""File-level end to end (SURVEY.md 8d E2E boundary): JSON directories -> the 7 + 4 output files.

    python tools/bench_e2e_files.py [--lnc 100] [--cores 0] [--runner gpu|oracle] [--reference]

Thin wrapper: the measurement lives in bench.py (`python bench.py --stage files ...`), next to the other
legs that are allowed to run the CPU oracle / reference harness as a baseline."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.argv = [os.path.join(ROOT, "bench.py"), "--stage", "files"] + sys.argv[1:]
runpy.run_path(sys.argv[0], run_name="__main__")
