"""Import the UNTOUCHED reference classes (TEST INFRASTRUCTURE).  Probed, in order: ``$MINGRAPH_REFERENCE_ROOT``,
``/root/reference/MinGraph-UNet`` (build container only), ``oracle/_ref/MinGraph-UNet`` (the verbatim, git-ignored copy
``oracle/fetch_ref.py`` makes in ``__graft_entry__.build()`` — the one that travels to the GPU box) and
``baseline/_ref/MinGraph-UNet``.  The root is put on ``sys.path`` and the modules are imported in place, unmodified.
Only ``tests/``, ``smoke()`` and ``bench.py``'s CPU legs may use this.
"""
from __future__ import annotations

import importlib
import os
import sys
import warnings

_REPO = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
CANDIDATES = [p for p in (os.environ.get("MINGRAPH_REFERENCE_ROOT"), "/root/reference/MinGraph-UNet",
                          os.path.join(_REPO, "oracle", "_ref", "MinGraph-UNet"),
                          os.path.join(_REPO, "baseline", "_ref", "MinGraph-UNet")) if p]


def _is_ref(root: str) -> bool:
    return os.path.isfile(os.path.join(root, "model", "gat", "graph_attention.py"))


REF_ROOT = next((p for p in CANDIDATES if _is_ref(p)), CANDIDATES[0])


def available() -> bool:
    return _is_ref(REF_ROOT)


def load():
    """Returns a namespace with the reference classes on the hot path.

    ``PatchSegmentPredictor`` lives in ``scripts/train_end_to_end.py``
    (``:40-70``), whose import pulls cv2/tqdm/yaml; if that fails the attribute
    is ``None`` and callers fall back to a plain ``GATNetwork`` predictor (the
    reference's own ``use_gnn=True`` branch is exactly that, ``:43-54``).
    """
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    ns = type("Ref", (), {})()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # reference docstrings contain invalid escapes
        ga = importlib.import_module("model.gat.graph_attention")
        mc = importlib.import_module("model.graph_partition.mincut_refinement")
        pg = importlib.import_module("preprocessing.graph_construction.patch_graph_construction")
        ns.GraphAttentionLayer = ga.GraphAttentionLayer
        ns.MultiHeadGATLayer = ga.MultiHeadGATLayer
        ns.GATNetwork = ga.GATNetwork
        ns.MinCutRefinement = mc.MinCutRefinement
        ns.PatchGraphConstructor = pg.PatchGraphConstructor
        try:
            te = importlib.import_module("scripts.train_end_to_end")
            ns.PatchSegmentPredictor = te.PatchSegmentPredictor
        except Exception:                         # pragma: no cover - optional deps
            ns.PatchSegmentPredictor = None
    return ns
