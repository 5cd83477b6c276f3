"""mingraph_unet_b200 — B200-native (sm_100a) graph block of MinGraph-UNet.

Public surface = the reference's graph-block classes (same constructor / ``forward`` signatures,
tensor layouts and ``state_dict`` keys) plus the batched :class:`GraphBlock`.  Everything computes
in ``libmingraph_b200.so`` (C ABI: ``include/mingraph_b200.h``); importing this package fails if the
library has not been built — there is no CPU or eager-PyTorch fallback.
"""
from . import _lib

_lib.load()          # fail loudly when the CUDA library is missing

from . import ops  # noqa: E402
from .block import GraphBlock, GraphBlockOutput  # noqa: E402
from .graph import Graph  # noqa: E402
from .runner import CapturedGraphBlock, CapturedTrainStep, PipelinedGraphBlock  # noqa: E402
from .modules import (FeatureConsistencyLoss, FeatureFusion, GATNetwork, GraphAttentionLayer, MinCutRefinement,  # noqa: E402
                      MultiHeadGATLayer, PatchGraphConstructor, PatchSegmentPredictor, StackedGATNetwork, TVLoss)

__all__ = ["ops", "Graph", "GraphBlock", "CapturedGraphBlock", "CapturedTrainStep", "PipelinedGraphBlock", "GraphBlockOutput", "GATNetwork", "GraphAttentionLayer", "MinCutRefinement",
           "MultiHeadGATLayer", "PatchGraphConstructor", "PatchSegmentPredictor", "FeatureConsistencyLoss", "TVLoss", "StackedGATNetwork", "FeatureFusion"]
__version__ = "0.1.0"
