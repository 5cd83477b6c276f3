"""ctypes binding of ``libmingraph_b200.so`` (the C ABI declared in ``include/mingraph_b200.h``).

There is no CPU fallback: if the shared library is missing or a call fails, the product path
raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C mingraph_unet_b200/csrc -j8``.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Dict, List

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MINGRAPH_B200_LIB", os.path.join(_HERE, "lib", "libmingraph_b200.so"))
HEADER_PATH = os.path.abspath(os.path.join(_HERE, "..", "include", "mingraph_b200.h"))

MG_F32, MG_BF16, MG_I32, MG_I64 = 0, 1, 2, 3
MG_OK, MG_ERR_INVALID, MG_ERR_CUDA, MG_ERR_UNSUPPORTED = 0, -1, -2, -3


class MinGraphError(RuntimeError):
    """A C-ABI call returned a negative status."""

    def __init__(self, fn: str, code: int, text: str):
        super().__init__(f"{fn} failed ({code}): {text}")
        self.fn, self.code, self.text = fn, code, text


_p, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); must list every MG_API symbol of the header (tests check this)
PROTOTYPES: Dict[str, tuple] = {
    "mg_version": (_i, []),
    "mg_last_error": (C.c_char_p, []),
    "mg_launch_count": (_i64, []),
    "mg_grid_num_edges": (_i64, [_i, _i]),
    "mg_grid_edge_index": (_i, [_i, _i, _i, _i, _p, _p]),
    "mg_grid_csr": (_i, [_i, _i, _i, _p, _p, _p, _p, _p]),
    "mg_complete_edge_index": (_i, [_i, _i, _i, _p, _p]),
    "mg_complete_csr": (_i, [_i, _i, _p, _p, _p]),
    "mg_csr_work_bytes": (_i64, [_i, _i64]),
    "mg_csr_from_coo": (_i, [_p, _i64, _i, _i, _p, _p, _p, _p, _p, _p]),
    "mg_knn_graph": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "mg_pool_patches": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _p, _i, _p]),
    "mg_segment_work_bytes": (_i64, [_i, _i, _i, _i]),
    "mg_segment_mean": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "mg_gat_work_bytes": (_i64, [_i, _i, _i, _i, _i]),
    "mg_gat_uses_tensor_pipe": (_i, [_i] * 7),
    "mg_gat_forward": (_i, [_p, _i, _p, _p, _i, _i64, _p, _p, _i, _i, _i, _i, _f, _i, _f, C.c_uint64, _p, _p, _i, _p, _p, _p, _p]),
    "mg_edge_slot_map": (_i, [_p, _p, _i64, _p, _p, _p]),
    "mg_gat_backward_work_bytes": (_i64, [_i, _i64, _i, _i, _i, _i]),
    "mg_gat_backward": (_i, [_p, _i, _p, _p, _p, _p, _p, _i, _i64, _p, _p, _i, _i, _i, _i, _f, _i, _f, C.c_uint64, _p,
                             _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "mg_softmax_argmax": (_i, [_p, _i, _i, _p, _p, _p]),
    "mg_ncut_edge_weights": (_i, [_p, _i, _i, _p, _i64, _p, _p]),
    "mg_ncut_work_bytes": (_i64, [_i, _i, _i]),
    "mg_ncut_loss": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "mg_ncut_backward": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "mg_unpool_nearest": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _i, _i64, _p]),
    "mg_unpool_backward_work_bytes": (_i64, [_i, _i, _i, _i, _i]),
    "mg_unpool_nearest_backward": (_i, [_p, _i, _i64, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "mg_segment_mean_backward": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "mg_softmax_backward": (_i, [_p, _p, _i, _i, _p, _p]),
    "mg_block_prep_floats": (_i64, [_i] * 6),
    "mg_block_supported": (_i, [_i] * 9),
    "mg_block_prepare": (_i, [_p] * 6 + [_i] * 6 + [_p, _p]),
    "mg_block_forward": (_i, [_p] + [_i] * 10 + [_f] * 3 + [_p] * 9),
    "mg_feature_loss_work_bytes": (_i64, [_i, _i]),
    "mg_feature_consistency_loss": (_i, [_p, _i, _p, _i, _p, _i, _i, _i, _i, _f, _p, _p, _p, _p]),
    "mg_feature_consistency_loss_backward": (_i, [_p, _i, _p, _i, _p, _i, _i, _i, _i, _f, _p, _p, _p, _p]),
    "mg_tv_loss_work_bytes": (_i64, [_i] * 5),
    "mg_tv_loss": (_i, [_p, _i, _i, _i, _i, _i, _f, _p, _p, _p]),
    "mg_tv_loss_backward": (_i, [_p, _i, _i, _i, _i, _i, _f, _p, _p, _p]),
    "mg_region_map_gather": (_i, [_p, _i, _i, _p, _i, _i, _i, _i, _p, _i, _i64, _p]),
    "mg_block_forward_push": (_i, [_p] + [_i] * 10 + [_f] * 3 + [_p] * 10),
    "mg_block_forward_ex": (_i, [_p] + [_i] * 10 + [_f] * 3 + [_p] * 11),
    "mg_peer_wait": (_i, [_p, _i64, _i, _p, _p, _p]),
    "mg_peer_mem_alloc": (_i, [_i64, _p, _p]),
    "mg_peer_mem_open": (_i, [_p, _p]),
    "mg_peer_mem_close": (_i, [_p]),
    "mg_peer_mem_free": (_i, [_p]),
}



class PeerOut(C.Structure):
    """``mg_peer_out_t`` of the header: one (pipeline slot, rank) of the fused peer exchange."""
    _fields_ = [("peer_bufs_dev", _p), ("peer_flags_dev", _p), ("world", C.c_int32), ("rank", C.c_int32),
                ("slice_offset", _i64), ("parity_stride", _i64), ("flag_index", _i64), ("seq", _p), ("done", _p)]


class BlockFeatureLoss(C.Structure):
    """``mg_block_feature_loss_t`` of the header."""
    _fields_ = [("f_unet", _p), ("y", _p), ("margin", C.c_float), ("loss_per_image", _p)]


_lib = None


def header_symbols() -> List[str]:
    """Every ``MG_API`` function name declared in the public header."""
    with open(HEADER_PATH) as f:
        return re.findall(r"^MG_API\s+[\w\s\*]+?\b(mg_\w+)\s*\(", f.read(), flags=re.M)


def load() -> C.CDLL:
    """Load the library once and attach prototypes; raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA library of mingraph_unet_b200 is not built "
            "(run __graft_entry__.build() or make -C mingraph_unet_b200/csrc). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def call(name: str, *args) -> None:
    """Invoke a status-returning entry point; raise MinGraphError on a negative status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != MG_OK:
        raise MinGraphError(name, rc, lib.mg_last_error().decode(errors="replace"))


def launch_count() -> int:
    return int(load().mg_launch_count())
