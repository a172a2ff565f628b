"""ctypes binding of libo2a_cuda.so (include/o2a.h). Fails loudly: no CPU fallback exists."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libo2a_cuda.so")

EXPORTS = ["o2a_abi_version", "o2a_create", "o2a_destroy", "o2a_last_error", "o2a_set_stream",
           "o2a_host_alloc", "o2a_host_free", "o2a_build_batch", "o2a_fetch_lnc", "o2a_fetch_aln",
           "o2a_fetch_survivors", "o2a_fetch_chains", "o2a_fitter_eval", "o2a_best_chain",
           "o2a_launch_count", "o2a_graph_state", "o2a_coords_zero_copy", "o2a_pipeline_chunks", "o2a_set_coords_mode", "o2a_coords_mode", "o2a_fp64_probe", "o2a_set_timing", "o2a_get_stage_ms",
           "o2a_cut_batch", "o2a_fetch_cut", "o2a_cut_view", "o2a_get_cut_ms",
           "o2a_annotate_batch", "o2a_fetch_annotation", "o2a_get_annotate_ms"]

STAGES = ["h2d", "fit", "filter", "pq", "gather", "chain", "best"]


class O2ABatch(C.Structure):
    _fields_ = [("n_lnc", C.c_int64), ("n_aln", C.c_int64), ("n_hsp", C.c_int64), ("n_bg", C.c_int64),
                ("bg_off", C.c_void_p), ("bg", C.c_void_p), ("kde_factor", C.c_void_p),
                ("aln_off", C.c_void_p), ("hsp_off", C.c_void_p), ("qlen", C.c_void_p), ("slen", C.c_void_p),
                ("qstart", C.c_void_p), ("qend", C.c_void_p), ("sstart", C.c_void_p), ("send", C.c_void_p),
                ("score", C.c_void_p), ("mem", C.c_int32), ("reserved", C.c_int32),
                ("max_aln_hsps", C.c_int64), ("max_lnc_bg", C.c_int64), ("coords", C.c_void_p),
                ("score_bits", C.c_int32), ("bg_bits", C.c_int32), ("score_q", C.c_void_p),
                ("exc_off", C.c_void_p), ("exc_pos", C.c_void_p), ("exc_val", C.c_void_p)]


class O2ACutInput(C.Structure):
    _fields_ = [("n_aln", C.c_int64), ("n_hsp", C.c_int64), ("hsp_off", C.c_void_p),
                ("qstart", C.c_void_p), ("qend", C.c_void_p), ("sstart", C.c_void_p), ("send", C.c_void_p),
                ("score", C.c_void_p), ("score_is_int", C.c_void_p), ("q_start", C.c_void_p), ("q_end", C.c_void_p), ("q_minus", C.c_void_p),
                ("s_start", C.c_void_p), ("s_end", C.c_void_p), ("s_minus", C.c_void_p), ("relative", C.c_void_p),
                ("qleft", C.c_void_p), ("qright", C.c_void_p), ("sleft", C.c_void_p), ("sright", C.c_void_p),
                ("mem", C.c_int32), ("reserved", C.c_int32)]


class O2AAnnInput(C.Structure):
    _fields_ = [("n_q", C.c_int64), ("n_a", C.c_int64),
                ("q_chrom", C.c_void_p), ("q_start", C.c_void_p), ("q_end", C.c_void_p), ("q_name", C.c_void_p),
                ("q_strand", C.c_void_p),
                ("a_chrom", C.c_void_p), ("a_start", C.c_void_p), ("a_end", C.c_void_p), ("a_name", C.c_void_p),
                ("a_strand", C.c_void_p),
                ("distance", C.c_int64), ("strandness", C.c_int32), ("mem", C.c_int32)]


class O2AParams(C.Structure):
    _fields_ = [("fitter", C.c_int32), ("fdr", C.c_int32), ("alpha", C.c_double),
                ("gapopen", C.c_int32), ("gapextend", C.c_int32),
                ("best_value", C.c_int32), ("best_function", C.c_int32)]


class O2ACounts(C.Structure):
    _fields_ = [("n_survivors", C.c_int64), ("n_retained", C.c_int64), ("n_chain_hsps", C.c_int64),
                ("n_chains", C.c_int64), ("n_failed_lnc", C.c_int64)]


class LibraryMissing(RuntimeError):
    pass


_lib = None


def load():
    """Load the in-tree CUDA library; raise if it has not been built (`__graft_entry__.build()` /
    `python -m ortho2align_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(f"{LIB_PATH} not built: run `python -m ortho2align_b200.build` "
                             "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    P, i32, i64, d = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    lib.o2a_abi_version.restype = C.c_int
    lib.o2a_create.restype = C.c_int
    lib.o2a_create.argtypes = [C.POINTER(P), C.c_int]
    lib.o2a_destroy.restype = None
    lib.o2a_destroy.argtypes = [P]
    lib.o2a_last_error.restype = C.c_char_p
    lib.o2a_last_error.argtypes = [P]
    lib.o2a_set_stream.restype = C.c_int
    lib.o2a_set_stream.argtypes = [P, P]
    lib.o2a_host_alloc.restype = P
    lib.o2a_host_alloc.argtypes = [C.c_size_t]
    lib.o2a_host_free.restype = None
    lib.o2a_host_free.argtypes = [P]
    lib.o2a_build_batch.restype = C.c_int
    lib.o2a_build_batch.argtypes = [P, C.POINTER(O2ABatch), C.POINTER(O2AParams), C.POINTER(O2ACounts)]
    lib.o2a_fetch_lnc.restype = C.c_int
    lib.o2a_fetch_lnc.argtypes = [P, P, P, P]
    lib.o2a_fetch_aln.restype = C.c_int
    lib.o2a_fetch_aln.argtypes = [P] + [P] * 6
    lib.o2a_fetch_survivors.restype = C.c_int
    lib.o2a_fetch_survivors.argtypes = [P] + [P] * 5
    lib.o2a_fetch_chains.restype = C.c_int
    lib.o2a_fetch_chains.argtypes = [P] + [P] * 3
    lib.o2a_fitter_eval.restype = C.c_int
    lib.o2a_fitter_eval.argtypes = [P, i32, P, i64, d, P, i64, P, P, i64, P, P]
    lib.o2a_best_chain.restype = C.c_int
    lib.o2a_best_chain.argtypes = [P, i64, P, P, P, P, P, i64, i64, i32, i32, P, P, P, P, P, P]
    lib.o2a_launch_count.restype = i64
    lib.o2a_launch_count.argtypes = [P]
    lib.o2a_graph_state.restype = C.c_int
    lib.o2a_graph_state.argtypes = [P]
    lib.o2a_coords_zero_copy.restype = C.c_int
    lib.o2a_coords_zero_copy.argtypes = [P]
    lib.o2a_pipeline_chunks.restype = C.c_int
    lib.o2a_pipeline_chunks.argtypes = [P]
    lib.o2a_set_coords_mode.restype = C.c_int
    lib.o2a_set_coords_mode.argtypes = [P, C.c_int]
    lib.o2a_coords_mode.restype = C.c_int
    lib.o2a_coords_mode.argtypes = [P]
    lib.o2a_fp64_probe.restype = C.c_int
    lib.o2a_fp64_probe.argtypes = [P, C.c_int, i64, C.POINTER(d)]
    lib.o2a_set_timing.restype = C.c_int
    lib.o2a_set_timing.argtypes = [P, C.c_int]
    lib.o2a_get_stage_ms.restype = C.c_int
    lib.o2a_get_stage_ms.argtypes = [P, P]
    lib.o2a_cut_batch.restype = C.c_int
    lib.o2a_cut_batch.argtypes = [P, C.POINTER(O2ACutInput), C.POINTER(i64)]
    lib.o2a_fetch_cut.restype = C.c_int
    lib.o2a_fetch_cut.argtypes = [P] + [P] * 10
    lib.o2a_cut_view.restype = C.c_int
    lib.o2a_cut_view.argtypes = [P, C.POINTER(O2ABatch)]
    lib.o2a_get_cut_ms.restype = C.c_int
    lib.o2a_get_cut_ms.argtypes = [P, P]
    lib.o2a_annotate_batch.restype = C.c_int
    lib.o2a_annotate_batch.argtypes = [P, C.POINTER(O2AAnnInput), C.POINTER(i64)]
    lib.o2a_fetch_annotation.restype = C.c_int
    lib.o2a_fetch_annotation.argtypes = [P] + [P] * 4
    lib.o2a_get_annotate_ms.restype = C.c_int
    lib.o2a_get_annotate_ms.argtypes = [P, P]
    if lib.o2a_abi_version() != 1:
        raise LibraryMissing("libo2a_cuda.so ABI version mismatch; rebuild")
    _lib = lib
    return lib
