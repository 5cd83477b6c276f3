"""CPU: the C-ABI library loads and exports every symbol of include/mingraph_b200.h, the host-side
mirror has the reference's interface (constructor signatures, state_dict keys/shapes, error
behaviour that does not need a device), and the multi-GPU plumbing works over gloo (world 2)."""
import ctypes
import inspect
import os
import subprocess
import sys

import pytest
import torch

from oracle import ref_loader

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def mg():
    import mingraph_unet_b200 as m
    return m


def test_library_exports_every_header_symbol(mg):
    from mingraph_unet_b200 import _lib
    names = _lib.header_symbols()
    assert len(names) >= 19 and len(set(names)) == len(names)
    assert set(names) == set(_lib.PROTOTYPES)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert _lib.load().mg_version() == 100
    assert _lib.load().mg_grid_num_edges(16, 16) == 960 and _lib.load().mg_grid_num_edges(1, 1) == 0
    assert _lib.load().mg_grid_num_edges(64, 64) == 16128


def test_argument_validation_without_device(mg):
    """Host-side argument checks return MG_ERR_INVALID before any CUDA call."""
    from mingraph_unet_b200 import _lib
    with pytest.raises(_lib.MinGraphError) as e:
        _lib.call("mg_grid_edge_index", 0, 4, 1, 0, None, None)
    assert e.value.code == _lib.MG_ERR_INVALID and "bad grid" in e.value.text
    with pytest.raises(_lib.MinGraphError) as e:
        _lib.call("mg_gat_forward", 1, 0, 1, None, 4, 0, 1, 1, 8, 8, 1, 0, 0.2, 0, 0.0, 0, None, 1, 0, 1, None, None, None)
    assert "empty edge_index" in e.value.text


def test_no_cpu_fallback(mg):
    x = torch.randn(4, 8)
    ei = torch.tensor([[0, 1], [1, 0]])
    layer = mg.MultiHeadGATLayer(8, 4, 2, 0.0, 0.2, concat=False).eval()
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        layer(x, ei)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        mg.PatchGraphConstructor(16).construct_patch_graph(torch.zeros(3, 32, 32), torch.zeros(4, 8))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        mg.ops.pool_patches(torch.zeros(1, 2, 32, 32), 16)


def test_missing_library_fails_loudly():
    code = ("import os; os.environ['MINGRAPH_B200_LIB']='/nonexistent/lib.so'\n"
            "try:\n    import mingraph_unet_b200\nexcept ImportError as e:\n    print('IMPORTERROR', 'no CPU fallback' in str(e).replace('There is no CPU fallback','no CPU fallback'))\n")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert "IMPORTERROR True" in out.stdout, out.stdout + out.stderr


def test_constructor_errors_match_reference(mg):
    with pytest.raises(AssertionError):
        mg.MultiHeadGATLayer(8, 5, 2, 0.0, 0.2, concat=True)           # graph_attention.py:138
    with pytest.raises(ValueError):
        mg.PatchGraphConstructor(16).construct_patch_graph(torch.zeros(3, 64, 64), torch.zeros(5, 8))
    with pytest.raises(ValueError):
        mg.PatchSegmentPredictor(8, 2, use_gnn=True)(torch.zeros(4, 8), None)   # train_end_to_end.py:66-67
    with pytest.raises(ValueError):
        mg.MinCutRefinement()(torch.zeros(4, 8), torch.zeros(2, 0, dtype=torch.long), 2, None)
    with pytest.raises(NotImplementedError):
        mg.PatchGraphConstructor().get_patch_features_from_unet_encoder(None, None)


def test_default_block_parameter_count(mg):
    blk = mg.GraphBlock()
    counts = [sum(p.numel() for p in m.parameters()) for m in (blk.patch_gat_model, blk.segment_predictor, blk.region_gat_model)]
    assert counts == [5632, 264, 16896]


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_interface_matches_live_reference(mg):
    R = ref_loader.load()
    pairs = [(R.GraphAttentionLayer, mg.GraphAttentionLayer), (R.MultiHeadGATLayer, mg.MultiHeadGATLayer),
             (R.GATNetwork, mg.GATNetwork), (R.MinCutRefinement, mg.MinCutRefinement),
             (R.PatchGraphConstructor, mg.PatchGraphConstructor)]
    if R.PatchSegmentPredictor is not None:
        pairs.append((R.PatchSegmentPredictor, mg.PatchSegmentPredictor))
        import importlib
        pairs.append((importlib.import_module("scripts.train_end_to_end").TVLoss, mg.TVLoss))
        pairs.append((importlib.import_module("model.unet.feature_loss").FeatureConsistencyLoss, mg.FeatureConsistencyLoss))
    pairs.append((R.GATNetwork, mg.StackedGATNetwork))
    for ref_cls, our_cls in pairs:
        assert str(inspect.signature(ref_cls.__init__)) == str(inspect.signature(our_cls.__init__)), ref_cls.__name__
        if hasattr(ref_cls, "forward"):
            assert list(inspect.signature(ref_cls.forward).parameters) == list(inspect.signature(our_cls.forward).parameters)
    for name in ("image_to_patches", "construct_patch_graph", "get_patch_features_from_unet_encoder"):
        assert list(inspect.signature(getattr(R.PatchGraphConstructor, name)).parameters) == \
            list(inspect.signature(getattr(mg.PatchGraphConstructor, name)).parameters)
    for name in ("compute_edge_weights_for_ncut", "normalized_cut_loss"):
        assert list(inspect.signature(getattr(R.MinCutRefinement, name)).parameters) == \
            list(inspect.signature(getattr(mg.MinCutRefinement, name)).parameters)
    # state_dict: same keys and shapes, loads both ways, same init distribution bounds
    for args in [(20, 128, 64, 4, 1, 0.1, 0.2), (64, 32, 2, 2, 1), (12, 16, 8, 2, 3)]:
        torch.manual_seed(0)
        ref = R.GATNetwork(*args)
        torch.manual_seed(0)
        ours = mg.GATNetwork(*args)
        rs, os_ = ref.state_dict(), ours.state_dict()
        assert list(rs) == list(os_)
        for k in rs:
            assert rs[k].shape == os_[k].shape and torch.equal(rs[k], os_[k])       # same RNG consumption order
        ours.load_state_dict(rs)
        ref.load_state_dict(os_)
    if R.PatchSegmentPredictor is not None:
        for kw in (dict(use_gnn=True, num_heads=2, hidden_dim=32), dict(use_gnn=False)):
            a, b = R.PatchSegmentPredictor(64, 2, **kw), mg.PatchSegmentPredictor(64, 2, **kw)
            assert {k: v.shape for k, v in a.state_dict().items()} == {k: v.shape for k, v in b.state_dict().items()}
    # image_to_patches is a view-level restatement: identical on CPU
    img = torch.randn(3, 70, 75)
    pr, gr = R.PatchGraphConstructor(16).image_to_patches(img)
    po, go = mg.PatchGraphConstructor(16).image_to_patches(img)
    assert gr == go and torch.equal(pr, po)


def test_shard_range():
    from mingraph_unet_b200.distributed import shard_range
    for B, world in [(64, 8), (64, 2), (16, 1), (10, 4), (3, 8), (32, 8)]:
        spans = [shard_range(B, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == B
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 4, 4)


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["MG_ROOT"])
from mingraph_unet_b200.distributed import shard_range, gather_block_outputs, allreduce_graph_grads
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
Bg, K, D, N = int(os.environ["MG_BG"]), 2, 4, 6
g = torch.Generator().manual_seed(0)
loss = torch.rand(Bg, generator=g); reg = torch.randn(Bg, K, D, generator=g)
lab = torch.randint(0, K, (Bg, N), generator=g).int()
lo, hi = shard_range(Bg, rank, world)
out = gather_block_outputs(loss[lo:hi], reg[lo:hi], lab[lo:hi], Bg)
assert torch.equal(out.l_partition, loss) and torch.equal(out.region_features, reg) and torch.equal(out.hard_labels, lab)
nf = out.node_features()
assert nf.shape == (Bg, N, D) and torch.equal(nf[1, 3], reg[1, lab[1, 3].item()])
m = torch.nn.Linear(3, 2)
for p in m.parameters():
    p.grad = torch.full_like(p, float(rank + 1))
n = allreduce_graph_grads(m)
assert n == 8 and all(torch.allclose(p.grad, torch.full_like(p, (world + 1) / 2)) for p in m.parameters())
# InlineGather: the packed per-slot payload (what the block kernel writes in place) gathered in step order
from mingraph_unet_b200.distributed import InlineGather
if Bg % world == 0:
    B = Bg // world
    ig = InlineGather(B, N, K, D, "cpu", depth=3)
    for step in range(5):
        slot = step % 3
        sl = slice(rank * B, (rank + 1) * B)
        pk = ig.packed[slot]
        pk[:B] = loss[sl] + step
        pk[B:B + B * K * D] = reg[sl].reshape(-1)
        pk[B + B * K * D:].view(torch.int32).copy_(lab[sl].reshape(-1))
        ig.gather(slot)
        v = ig.views(slot)
        assert torch.equal(v.l_partition, loss + step) and torch.equal(v.region_features, reg) and torch.equal(v.hard_labels, lab)
dist.destroy_process_group()
print("OK", rank)
"""


@pytest.mark.parametrize("Bg", [8, 5])
def test_gloo_world2_gather_and_grad_allreduce(tmp_path, Bg):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MG_ROOT=ROOT, MG_BG=str(Bg), MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + Bg), str(script)]
    out = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.count("OK") == 2, out.stdout[-2000:] + out.stderr[-4000:]


def test_bench_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--workload", "cfg1"], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0
    # the arm is the UNTOUCHED reference whenever a copy is reachable (oracle/_ref from build(), or /root/reference)
    from oracle import ref_loader
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_loader.available() else "port")


def test_reference_arm_falls_back_to_the_port_when_the_copy_is_absent(tmp_path):
    env = dict(os.environ, MINGRAPH_REFERENCE_ROOT=str(tmp_path))
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from oracle import ref_loader\n"
            "ref_loader.CANDIDATES[:] = [%r]; ref_loader.REF_ROOT = %r\n"
            "import bench\narm = bench.CpuArm('cfg1')\nprint(arm.kind, arm.run(1) > 0)\n") % (ROOT, str(tmp_path), str(tmp_path))
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "port True" in out.stdout, out.stdout + out.stderr[-2000:]


def test_fetched_reference_copy_is_verbatim():
    """oracle/_ref (made by oracle/fetch_ref.py in build()) is byte-identical to the mounted reference, file by file."""
    src = "/root/reference/MinGraph-UNet"
    dst = os.path.join(ROOT, "oracle", "_ref", "MinGraph-UNet")
    if not (os.path.isdir(src) and os.path.isdir(dst)):
        pytest.skip("needs both the mounted reference and the fetched copy (build container after build())")
    n = 0
    for root, _, files in os.walk(dst):
        for f in files:
            rel = os.path.relpath(os.path.join(root, f), dst)
            with open(os.path.join(dst, rel), "rb") as a, open(os.path.join(src, rel), "rb") as b:
                assert a.read() == b.read(), rel
            n += 1
    assert n >= 30


def test_prepared_weight_cache_tracks_every_kind_of_update(mg, monkeypatch):
    """GraphBlock._prepared() re-arranges the weights only when a parameter changed.  The module-tree walk is cached (it
    costs as much host time as a pipelined step), so every way weights can change must still be noticed."""
    import time
    calls = []
    monkeypatch.setattr(mg.ops, "block_prepare", lambda *stacks, out=None: calls.append(1) or torch.zeros(1))
    blk = mg.GraphBlock()
    blk._prepared(); blk._prepared()
    assert len(calls) == 1
    with torch.no_grad():                                                    # optimizer-style in-place update
        blk.region_gat_model.gat_layers[0].heads[3].a.weight.add_(1.0)
    blk._prepared(); blk._prepared()
    assert len(calls) == 2
    blk.patch_gat_model.load_state_dict(mg.GATNetwork(20, 128, 64, 4).state_dict())           # copy into the same tensors
    blk._prepared()
    assert len(calls) == 3
    blk.segment_predictor.gnn_predictor.load_state_dict(mg.GATNetwork(64, 32, 2, 2).state_dict(), assign=True)   # new tensors
    blk._prepared()
    assert len(calls) == 4
    blk.region_gat_model = mg.GATNetwork(64, 128, 64, 4)                     # a sub-network replaced
    blk._prepared(); blk._prepared()
    assert len(calls) == 5
    blk.double(); blk._prepared()                                            # .to()/.double() re-allocate the data
    assert len(calls) == 6
    def best_of(fn, reps=5, n=100):
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            ts.append((time.perf_counter() - t0) / n)
        return min(ts)
    nets = (blk.patch_gat_model, blk.segment_predictor.gnn_predictor, blk.region_gat_model)
    walk = best_of(lambda: tuple((p.data_ptr(), p._version) for net in nets for p in net.parameters()))
    cached = best_of(blk._prepared)
    assert len(calls) == 6 and cached < walk / 3, (cached, walk)            # the per-step host cost stays small


def test_loss_modules_error_behaviour_on_cpu(mg):
    """Argument checks of the f4 loss modules run before any device work, so they are testable here: same exceptions as
    the reference (model/unet/feature_loss.py:91-101), and no CPU fallback."""
    fl = mg.FeatureConsistencyLoss(margin=1.0)
    x2 = torch.randn(6, 8)
    # the reference's own call site (scripts/train_end_to_end.py:344) passes 2-D (N, D) tensors and dies in the shape
    # unpacking at feature_loss.py:91 with a ValueError; the mirror keeps that contract
    with pytest.raises(ValueError):
        fl(x2, x2, torch.zeros(6))
    x3 = torch.randn(2, 6, 8)
    with pytest.raises(ValueError, match="must have same dimensions"):
        fl(x3, torch.randn(2, 6, 9), torch.zeros(2, 6))
    with pytest.raises(ValueError, match="is not \\(Batch, Num_Patches\\)"):
        fl(x3, x3, torch.zeros(2, 7))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        fl(x3, x3, torch.zeros(2, 6))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        mg.TVLoss()(torch.rand(1, 1, 8, 8))
    if ref_loader.available():
        import importlib
        ref_loader.load()
        ref_fl = importlib.import_module("model.unet.feature_loss").FeatureConsistencyLoss(1.0)
        with pytest.raises(ValueError):
            ref_fl(x2, x2, torch.zeros(6))
        with pytest.raises(ValueError, match="must have same dimensions"):
            ref_fl(x3, torch.randn(2, 6, 9), torch.zeros(2, 6))
        with pytest.raises(ValueError, match="is not \\(Batch, Num_Patches\\)"):
            ref_fl(x3, x3, torch.zeros(2, 7))


def test_numa_binding_never_raises():
    """bind_to_gpu_numa is an optimisation of the N>1 bench ranks: without NVML / a GPU it reports why it did nothing."""
    from mingraph_unet_b200.distributed import bind_to_gpu_numa
    before = os.sched_getaffinity(0)
    msg = bind_to_gpu_numa("cuda:0")
    assert isinstance(msg, str) and (msg.startswith("bound to") or msg.startswith("not bound"))
    if msg.startswith("not bound"):
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)


def test_pdl_kernels_read_predecessor_outputs_with_ordinary_loads():
    """Source lint for csrc/gat_tc.cu, the only file that chains kernels with programmatic dependent launch: ``__ldg`` (ld.global.nc)
    may be scheduled above ``griddepcontrol.wait`` by ptxas, so it is reserved for the layer's inputs; everything a predecessor
    grid wrote (u, s, gmax, gsrc) goes through ``ld_pre``.  (The ordering itself is checked on the GPU by
    tests/test_gpu_tc.py::test_bf16_prepass_in_front_of_the_spilled_path.)"""
    import re
    src = open(os.path.join(os.path.dirname(__file__), "..", "mingraph_unet_b200", "csrc", "gat_tc.cu")).read()
    src = re.sub(r"//[^\n]*", "", src)                                       # comments mention __ldg too
    allowed = ("x +", "x_lane", "A.x", "rowptr", "col +", "A.col", "Wh +", "ah +", "A.W")
    calls = re.findall(r"__ldg\(([^;]*?)\)\s*[;:,)]", src)
    assert len(calls) >= 20
    for arg in calls:
        assert any(tok in arg for tok in allowed), f"__ldg on something that is not a layer input: {arg!r}"
        for bad in ("u_s", "s_in", "gmax", "gsrc", "s_tgt_lane", "A.s", "(s +"):
            assert bad not in arg, f"__ldg on a predecessor grid's output: {arg!r}"
    assert src.count("ld_pre(") >= 10
