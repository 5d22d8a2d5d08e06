"""ctypes binding of libagnn.so (the C ABI declared in include/agnn.h).

There is deliberately no fallback: if the library is missing or a call fails the
caller gets an exception.  ``build()`` compiles it in-tree with nvcc for sm_100a
(``analysisgnn_b200/csrc/Makefile``); the built ``.so`` is git-ignored but ships
to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AGNN_LIB_PATH") or os.path.join(_HERE, "lib", "libagnn.so")
CSRC = os.path.join(_HERE, "csrc")

MAX_SEG = 32
MAX_REL = 16
HGT_MAX_HEADS = 16
F32, BF16 = 0, 1
SCALE_NONE, SCALE_MEAN = 0, 1
COMBINE_CONCAT, COMBINE_SUM = 0, 1
REL_IDENTITY_IF_EMPTY = 1
REL_LOW_DEGREE = 2
EDGE_ABSDIFF, EDGE_ABSDIFF_DROW, EDGE_ABSDIFF_DNBR, EDGE_GATE, EDGE_GATE_DROW, EDGE_GATE_DNBR = range(6)
# HEAVY_ROW / HEAVY_CHUNK: build constants of the library (agnn_heavy_params), resolved on first use (__getattr__ below)
GEMM_TF32X3, GEMM_TF32, GEMM_BF16, GEMM_F16X3 = 0, 1, 2, 3
K_MAJOR, MN_MAJOR = 0, 1
GEMM_RELU, GEMM_ACCUMULATE, GEMM_OUT_BF16 = 1, 2, 4


class AgnnError(RuntimeError):
    pass


# Routes off the hand-written kernels -- an operand repack before agnn_gemm, a library kernel for a shape the row
# kernels do not take (odd widths, GRU sizes) -- are counted here; under ``set_strict(True)`` (bench.py, the full-size
# parity tests) they raise instead, so a measured or parity-checked hot path provably has none.
_strict = os.environ.get("AGNN_STRICT", "0") not in ("", "0")
library_routes = {}


def set_strict(on: bool) -> None:
    global _strict
    _strict = bool(on)


def strict() -> bool:
    return _strict


def library_route(what: str, counter=None, key=None) -> None:
    library_routes[what] = library_routes.get(what, 0) + 1
    if counter is not None:
        counter[key] = counter.get(key, 0) + 1
    if _strict:
        raise AgnnError(f"strict mode: {what}")


class Coo(C.Structure):
    _fields_ = [("row", C.c_void_p), ("col", C.c_void_p), ("etype", C.c_void_p), ("n_edges", C.c_int64),
                ("n_rows", C.c_int32), ("n_cols", C.c_int32), ("n_rel", C.c_int32), ("reserved", C.c_int32),
                ("rowptr_off", C.c_int64), ("edge_off", C.c_int64), ("heavy_off", C.c_int64), ("count_off", C.c_int64),
                ("heavy_cap", C.c_int64)]


class Rel(C.Structure):
    _fields_ = [("rowptr", C.c_void_p), ("col", C.c_void_p), ("src", C.c_void_p), ("ld_src", C.c_int64),
                ("nbr_deg_rowptr", C.c_void_p), ("out_col", C.c_int32), ("flags", C.c_int32),
                ("heavy_rows", C.c_void_p), ("n_heavy", C.c_void_p), ("heavy_cap", C.c_int64)]


class HgtRel(C.Structure):
    _fields_ = [("rowptr", C.c_void_p), ("col", C.c_void_p), ("t_rowptr", C.c_void_p), ("t_col", C.c_void_p),
                ("k", C.c_void_p), ("v", C.c_void_p), ("ld_kv", C.c_int64), ("dk", C.c_void_p), ("dv", C.c_void_p),
                ("ld_dkv", C.c_int64), ("n_src", C.c_int32), ("reserved", C.c_int32)]


class GemmProblem(C.Structure):
    """agnn_gemm_problem_t"""
    _fields_ = [("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64),
                ("a_hi", C.c_void_p), ("a_lo", C.c_void_p), ("lda", C.c_int64), ("amax_a", C.c_void_p),
                ("b_hi", C.c_void_p), ("b_lo", C.c_void_p), ("ldb", C.c_int64), ("amax_b", C.c_void_p),
                ("c", C.c_void_p), ("ldc", C.c_int64), ("bias", C.c_void_p), ("flags", C.c_int32),
                ("split_k", C.c_int32), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("amax_out", C.c_void_p)]


class SplitItem(C.Structure):
    """agnn_split_item_t"""
    _fields_ = [("x", C.c_void_p), ("rows", C.c_int64), ("cols", C.c_int64), ("ld_x", C.c_int64), ("hi", C.c_void_p),
                ("lo", C.c_void_p), ("ld_out", C.c_int64), ("amax", C.c_void_p)]


GEMM_MAX_GROUP = 12
SPLIT_MULTI_MAX = 24


class ParamChunk(C.Structure):
    _fields_ = [("param", C.c_void_p), ("param_off", C.c_int64), ("arena_off", C.c_int64), ("count", C.c_int32),
                ("param_aligned", C.c_int32)]


_PROTOTYPES = {
    "agnn_version": (C.c_int, []),
    "agnn_last_error": (C.c_char_p, []),
    "agnn_csr_build_workspace": (C.c_size_t, [C.c_int, C.POINTER(Coo)]),
    "agnn_csr_build": (C.c_int, [C.c_int, C.POINTER(Coo), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "agnn_gather_reduce": (C.c_int, [C.c_int32, C.c_int32, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Rel),
                                     C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                     C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "agnn_gather_reduce_f16": (C.c_int, [C.c_int32, C.c_int32, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Rel),
                                         C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                         C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "agnn_gather_heavy_workspace": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int32]),
    "agnn_heavy_params": (None, [C.c_void_p, C.c_void_p]),
    "agnn_edge_op": (C.c_int, [C.c_int, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                               C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int,
                               C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "agnn_rowscale_sum": (C.c_int, [C.c_int32, C.c_int32, C.c_int, C.c_int, C.POINTER(Rel), C.c_void_p, C.c_int64,
                                    C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "agnn_hgt_attn_fwd": (C.c_int, [C.c_int32, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(HgtRel), C.c_void_p,
                                    C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "agnn_hgt_attn_bwd_dst_blocks": (C.c_int, [C.c_int32]),
    "agnn_hgt_attn_bwd_dst": (C.c_int, [C.c_int32, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(HgtRel), C.c_void_p,
                                        C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                        C.c_void_p]),
    "agnn_hgt_attn_bwd_src": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(HgtRel), C.c_void_p, C.c_int64,
                                        C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p]),
    "agnn_optim_chunk_elems": (C.c_int, []),
    "agnn_sumsq_blocks": (C.c_int, [C.c_int64]),
    "agnn_sumsq_partials": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "agnn_adamw_clip_step": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                       C.c_float, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_float,
                                       C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "agnn_split_tf32": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                  C.c_void_p]),
    "agnn_gemm_split_k": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_int64]),
    "agnn_gemm_workspace": (C.c_size_t, [C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int]),
    "agnn_gemm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                            C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int,
                            C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "agnn_amax": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "agnn_split_f16": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int64, C.c_void_p]),
    "agnn_gemm_scaled": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "agnn_split_f16_multi": (C.c_int, [C.c_int, C.POINTER(SplitItem), C.c_void_p]),
    "agnn_gemm_pair_supported": (C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int]),
    "agnn_gemm_pair": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                 C.c_int, C.c_void_p, C.c_void_p]),
    "agnn_gemm_tickets": (C.c_int64, [C.c_int64, C.c_int64, C.c_int]),
    "agnn_gemm_group_split_k": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                          C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "agnn_gemm_grouped": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(GemmProblem), C.c_void_p, C.c_int64,
                                    C.c_void_p]),
    "agnn_gather_reduce_amax": (C.c_int, [C.c_int32, C.c_int32, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Rel),
                                          C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                          C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                          C.c_void_p]),
    "agnn_softmax_ce_bwd_padded": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float,
                                             C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                             C.c_void_p, C.c_void_p]),
    "agnn_layernorm_fwd_pair": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                          C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                          C.c_int64, C.c_void_p, C.c_float, C.c_void_p, C.c_uint32, C.c_void_p]),
    "agnn_layernorm_bwd_dropout": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_uint32,
                                             C.c_void_p, C.c_void_p]),
    "agnn_dropout_advance": (C.c_int, [C.c_void_p, C.c_void_p]),
    "agnn_dropout_apply": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_float,
                                     C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]),
    "agnn_split_f16_dropout": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int64, C.c_float, C.c_void_p, C.c_uint32, C.c_void_p]),
    "agnn_split_f16_shifted": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int64, C.c_float, C.c_void_p, C.c_uint32, C.c_int32, C.c_int32, C.c_void_p]),
    "agnn_sage_weights": (C.c_int, [C.c_int, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "agnn_row_blocks": (C.c_int, [C.c_int64]),
    "agnn_layernorm_fwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                     C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_void_p]),
    "agnn_layernorm_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_int, C.c_void_p]),
    "agnn_l2norm_relu_fwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int,
                                       C.c_int, C.c_float, C.c_void_p]),
    "agnn_l2norm_relu_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                       C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "agnn_colsum_partials": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "agnn_grad_prepare": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                    C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "agnn_grad_prepare_f16": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "agnn_ce_blocks": (C.c_int, [C.c_int64]),
    "agnn_softmax_ce_fwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_int64,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "agnn_softmax_ce_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float,
                                      C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "agnn_embedding_bwd_blocks": (C.c_int, [C.c_int64]),
    "agnn_embedding_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "agnn_softmax2_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64,
                                     C.c_void_p]),
    "agnn_run_heads_workspace": (C.c_size_t, [C.c_int64]),
    "agnn_run_heads": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_size_t, C.c_void_p]),
    "agnn_row_argmax": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                  C.c_void_p, C.c_void_p]),
    "agnn_decode_assign": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "agnn_score_graph_workspace": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int64]),
    "agnn_score_graph_build": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                         C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t,
                                         C.c_void_p]),
    "agnn_window_workspace": (C.c_size_t, [C.c_int64]),
    "agnn_window_subgraph": (C.c_int, [C.c_int32, C.c_int64] + [C.c_void_p] * 10 + [C.c_int64, C.c_void_p, C.c_void_p,
                                                                                    C.c_size_t, C.c_void_p]),
    "agnn_sample_init": (C.c_int, [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "agnn_sample_hop_workspace": (C.c_size_t, [C.c_int32, C.c_int32]),
    "agnn_sample_hop_count": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "agnn_sample_hop_draw": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.c_void_p, C.c_int32] +
                             [C.c_void_p] * 8 + [C.c_size_t, C.c_void_p]),
    "agnn_gru_supported": (C.c_int, [C.c_int]),
    "agnn_gru_mode": (C.c_int, [C.c_int]),
    "agnn_gru_fwd": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p]),
    "agnn_gru_bwd": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "agnn_gru_bwd_amax": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "agnn_gru_bwd_stepwise": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}

GRU_RESIDENT, GRU_STEPWISE = 1, 2      # agnn_gru_supported (include/agnn.h)

_lib = None
_launches = 0


def ptr_array(tensors):
    """A host array of device pointers (``const float* const*`` arguments)."""
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return arr


def count_launches(n: int) -> None:
    """Book-keeping of kernels launched through the C ABI (reported by bench.py)."""
    global _launches
    _launches += n


def launches() -> int:
    return _launches


def build(verbose: bool = False) -> str:
    """Compile libagnn.so in-tree (nvcc, -gencode arch=compute_100a,code=sm_100a)."""
    proc = subprocess.run(["make", "-C", CSRC, "-j", str(os.cpu_count() or 4)], capture_output=True, text=True)
    if proc.returncode != 0:
        raise AgnnError("building libagnn.so failed:\n" + proc.stdout[-4000:] + proc.stderr[-4000:])
    if verbose:
        print(proc.stdout)
    return LIB_PATH


def exported_symbols():
    return sorted(_PROTOTYPES)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise AgnnError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU or PyTorch fallback for the CUDA path)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int, what: str = "libagnn") -> None:
    if rc != 0:
        raise AgnnError(f"{what} failed ({rc}): {lib().agnn_last_error().decode()}")


_heavy = None


def heavy_params():
    """(HEAVY_ROW, HEAVY_CHUNK) the loaded library was built with: rows of at least HEAVY_ROW entries are listed by
    agnn_csr_build and split into HEAVY_CHUNK-edge chunks by agnn_gather_reduce."""
    global _heavy
    if _heavy is None:
        a, b = C.c_int32(0), C.c_int32(0)
        lib().agnn_heavy_params(C.byref(a), C.byref(b))
        _heavy = (int(a.value), int(b.value))
    return _heavy


def __getattr__(name):
    if name == "HEAVY_ROW":
        return heavy_params()[0]
    if name == "HEAVY_CHUNK":
        return heavy_params()[1]
    raise AttributeError(name)
