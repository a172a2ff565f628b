"""Command line: the reference's `build_orthologs`, `get_best_orthologs` and `annotate_orthologs` sub-commands
with the same flags and '@config' files (`cli_scripts.py:7-12, 254-405, 556-561`), plus `-devices` / `-device`."""
from __future__ import annotations

import argparse

from .annotate import annotate_orthologs
from .pipeline import build_orthologs, get_best_orthologs


def make_parser():
    p = argparse.ArgumentParser(prog="ortho2align_b200", fromfile_prefix_chars="@",
                                description="B200-native build_orthologs / get_best_orthologs of ortho2align.")
    sub = p.add_subparsers(title="Subcommands", metavar="SUBCOMMAND", required=True)
    b = sub.add_parser("build_orthologs", help="Asses orthologous alignments and build orthologs.")
    b.set_defaults(func=build_orthologs)
    b.add_argument("-alignments", type=str, nargs="?", required=True)
    b.add_argument("-background", type=str, nargs="?", required=True)
    b.add_argument("-fitting", type=str, nargs="?", choices=["kde", "hist", "empirical"], default="kde")
    b.add_argument("-threshold", type=float, nargs="?", default=0.05)
    b.add_argument("-gapopen", type=int, nargs="?", default=5)
    b.add_argument("-gapextend", type=int, nargs="?", default=2)
    b.add_argument("--fdr", action="store_true")
    b.add_argument("-outdir", type=str, nargs="?", required=True)
    b.add_argument("-cores", type=int, nargs="?", default=1)
    b.add_argument("-timeout", type=int, nargs="?", default=None)
    b.add_argument("--silent", action="store_true")
    b.add_argument("-devices", type=int, nargs="*", default=None, help="CUDA device indices (default: 0)")
    b.add_argument("-cache", type=str, nargs="?", default=None,
                   help="columnar sidecar file: written after the first JSON parse, reused while inputs are unchanged")
    g = sub.add_parser("get_best_orthologs", help="Select one ortholog for each query gene.")
    g.set_defaults(func=get_best_orthologs)
    g.add_argument("-query_orthologs", type=str, nargs="?", required=True)
    g.add_argument("-query_total", type=str, nargs="?", required=True)
    g.add_argument("-subject_orthologs", type=str, nargs="?", required=True)
    g.add_argument("-subject_total", type=str, nargs="?", required=True)
    g.add_argument("-value", type=str, nargs="?", required=True,
                   choices=["total_length", "block_count", "block_length", "weight"])
    g.add_argument("-function", type=str, nargs="?", required=True, choices=["max", "min"])
    g.add_argument("-outfile_query", type=str, nargs="?", required=True)
    g.add_argument("-outfile_query_total", type=str, nargs="?", required=True)
    g.add_argument("-outfile_subject", type=str, nargs="?", required=True)
    g.add_argument("-outfile_subject_total", type=str, nargs="?", required=True)
    a = sub.add_parser("annotate_orthologs",
                       help="Annotate found orthologs with provided annotation of subject genome lncRNAs.")
    a.set_defaults(func=annotate_orthologs)
    a.add_argument("-subject_orthologs", type=str, nargs="?", required=True)
    a.add_argument("-subject_annotation", type=str, nargs="?", required=True)
    a.add_argument("-subject_name_regex", type=str, nargs="?", default=None)
    a.add_argument("-output", type=str, nargs="?", required=True)
    a.add_argument("-float_precision", type=int, nargs="?", default=8)
    a.add_argument("-device", type=int, nargs="?", default=0, help="CUDA device index (default: 0)")
    return p


def main(argv=None):
    args = make_parser().parse_args(argv)
    func = args.func
    kwargs = vars(args)
    del kwargs["func"]
    func(**kwargs)


if __name__ == "__main__":
    main()
