"""CUDA-graph replay of the graph block for a fixed input shape.

The block is a short chain of small kernels around two HBM-streaming ones (pool, un-pool); at the
image sizes of the reference's configs the host-side launch path costs more than the kernels.
``CapturedGraphBlock`` records one forward (pool -> fused block -> un-pool into the caller's fusion
buffer) into a CUDA graph and replays it with a single driver call per step.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .block import GraphBlock, GraphBlockOutput


class CapturedGraphBlock:
    """``runner = CapturedGraphBlock(block, example, image_size, out=fusion[:, 32:])``;
    ``runner(x)`` copies ``x`` into the static input (skipped when ``x`` is None or already the
    static buffer) and replays.  Outputs are static tensors overwritten by every replay.

    ``example`` is a per-pixel feature map ``(B,C,H,W)`` (``kind='feature_map'``) or node features
    ``(B,N,in)`` (``kind='node_features'``).  Weight updates are picked up: the prepared-weight blob
    is refreshed in place before the replay when a parameter changed."""

    def __init__(self, block: GraphBlock, example: torch.Tensor, image_size: Optional[Tuple[int, int]] = None,
                 out: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None, want_dense: bool = True,
                 warmup: int = 2):
        if not example.is_cuda:
            raise RuntimeError("mingraph_unet_b200 runs on CUDA tensors only (there is no CPU fallback)")
        if block.training and torch.is_grad_enabled():
            raise RuntimeError("CapturedGraphBlock replays the inference path: call block.eval() first")
        self.block = block
        self.kind = "feature_map" if example.dim() == 4 else "node_features"
        self.static_in = example.clone()
        self._kw = dict(image_size=image_size, out=out, out_dtype=out_dtype, want_dense=want_dense)
        side = torch.cuda.Stream(device=example.device)
        side.wait_stream(torch.cuda.current_stream(example.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):
                self._forward()
        torch.cuda.current_stream(example.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.outputs: GraphBlockOutput = self._forward()

    def _forward(self) -> GraphBlockOutput:
        return self.block(**{self.kind: self.static_in}, **self._kw)

    def __call__(self, x: Optional[torch.Tensor] = None) -> GraphBlockOutput:
        if x is not None and x.data_ptr() != self.static_in.data_ptr():
            self.static_in.copy_(x, non_blocking=True)
        if self.block.fused:
            self.block._prepared()            # refresh the weight blob in place if a parameter changed
        self.graph.replay()
        return self.outputs
