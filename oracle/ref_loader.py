"""Import the UNTOUCHED reference classes when the reference tree is mounted
(TEST INFRASTRUCTURE).  ``/root/reference`` exists only in the build container,
never on the GPU box, so everything that runs there uses the committed fixtures
in ``tests/golden/`` instead.  Nothing is copied: the reference root is put on
``sys.path`` and its modules are imported in place.
"""
from __future__ import annotations

import importlib
import os
import sys
import warnings

REF_ROOT = os.environ.get("MINGRAPH_REFERENCE_ROOT", "/root/reference/MinGraph-UNet")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "model", "gat", "graph_attention.py"))


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
