"""GPU parity of scope row f4 — FeatureConsistencyLoss, TVLoss and the working multi-layer GAT stack —
through the C ABI against the fixtures of the untouched reference and the CPU oracle.
Tolerances: losses are fp32 sums of O(1e3..1e6) terms reduced in a different (fixed) order than
torch's, so they are compared at rel 2e-6 (f32 storage); features at the block's max-abs 1e-5."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import restate as O  # noqa: E402

FL_CASES = ["b3_n37_d64", "b1_n256_d64", "b2_n50_d7"]
TV_CASES = ["b2_c1_64x64", "b3_c2_37x53", "b1_c3_5x200", "b2_c2_19x8"]


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(scope="module")
def M():
    import mingraph_unet_b200 as m
    return m


@pytest.mark.parametrize("tag", FL_CASES)
@pytest.mark.parametrize("ydt", [torch.int64, torch.int32, torch.float32, torch.bool])
def test_feature_loss_golden(golden, M, tag, ydt):
    g = golden("losses.npz")
    fu = T(g[f"fl_{tag}_fu"]).cuda().requires_grad_(True)
    fg = T(g[f"fl_{tag}_fg"]).cuda().requires_grad_(True)
    y = T(g[f"fl_{tag}_y"]).cuda().to(ydt)
    mod = M.FeatureConsistencyLoss(margin=float(g[f"fl_{tag}_margin"]))
    loss = mod(fu, fg, y)
    assert loss.dim() == 0 and loss.dtype == torch.float32
    loss.backward()
    assert float(loss) == pytest.approx(float(g[f"fl_{tag}_loss"]), rel=2e-6)
    assert float((fu.grad.cpu() - T(g[f"fl_{tag}_gfu"])).abs().max()) <= 1e-6
    assert float((fg.grad.cpu() - T(g[f"fl_{tag}_gfg"])).abs().max()) <= 1e-6
    # no-grad path gives the same number, bit for bit, twice (fixed-order reductions)
    with torch.no_grad():
        a, b = mod(fu, fg, y), mod(fu, fg, y)
    assert torch.equal(a, b) and torch.equal(a, loss.detach())


def test_feature_loss_errors_and_per_image(M):
    mod = M.FeatureConsistencyLoss()
    fu = torch.randn(2, 5, 8, device="cuda")
    with pytest.raises(ValueError, match="must have same dimensions"):
        mod(fu, torch.randn(2, 5, 9, device="cuda"), torch.zeros(2, 5, device="cuda"))
    with pytest.raises(ValueError, match="is not \\(Batch, Num_Patches\\)"):
        mod(fu, fu.clone(), torch.zeros(2, 6, device="cuda"))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        mod(fu.cpu(), fu.cpu(), torch.zeros(2, 5))
    fg = torch.randn(2, 5, 8, device="cuda")
    y = torch.tensor([[0, 1, 1, 0, 1], [1, 1, 0, 0, 0]], device="cuda")
    loss, per = M.ops.feature_consistency_loss(fu, fg, y, 1.0, with_per_image=True)
    ref = torch.stack([O.feature_consistency_loss(fu[b:b + 1].cpu(), fg[b:b + 1].cpu(), y[b:b + 1].cpu()) for b in range(2)])
    assert torch.allclose(per.cpu(), ref, rtol=2e-6) and float(loss) == pytest.approx(float(ref.mean()), rel=2e-6)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_feature_loss_block_sized(M, dt):
    """cfg 2 shape: 16 images x 1024 patches x 64 features, bf16 or f32 storage, fp32 math on both sides."""
    gen = torch.Generator().manual_seed(3)
    fu = (0.1 * torch.randn(16, 1024, 64, generator=gen)).to(dt)
    fg = (0.1 * torch.randn(16, 1024, 64, generator=gen)).to(dt)
    y = torch.randint(0, 2, (16, 1024), generator=gen)
    want = O.feature_consistency_loss(fu.float(), fg.float(), y)
    got = M.FeatureConsistencyLoss()(fu.cuda(), fg.cuda(), y.cuda())
    assert float(got) == pytest.approx(float(want), rel=5e-6)
    # mixed storage: the U-Net side in bf16, the graph side (patch-GAT output) in f32
    got2 = M.FeatureConsistencyLoss()(fu.cuda(), fg.float().cuda(), y.cuda())
    assert float(got2) == pytest.approx(float(want), rel=5e-6)


@pytest.mark.parametrize("tag", TV_CASES)
def test_tv_loss_golden(golden, M, tag):
    g = golden("losses.npz")
    x = T(g[f"tv_{tag}_x"]).cuda().requires_grad_(True)
    mod = M.TVLoss(float(g[f"tv_{tag}_weight"]))
    loss = mod(x)
    assert loss.dim() == 0 and loss.dtype == torch.float32
    loss.backward()
    assert float(loss) == pytest.approx(float(g[f"tv_{tag}_loss"]), rel=2e-6)
    assert float((x.grad.cpu() - T(g[f"tv_{tag}_gx"])).abs().max()) <= 1e-7
    with torch.no_grad():
        assert torch.equal(mod(x), mod(x))


@pytest.mark.parametrize("shape,dt", [((16, 1, 512, 512), torch.float32), ((4, 2, 512, 512), torch.bfloat16),
                                      ((2, 3, 129, 1000), torch.float32), ((1, 1, 300, 24), torch.bfloat16),
                                      ((2, 1, 1, 16), torch.float32), ((2, 1, 16, 1), torch.float32)])
def test_tv_loss_sizes(M, shape, dt):
    """Full-size maps (cfg 2 segmentation probability map), vector and scalar paths, and the degenerate
    H == 1 / W == 1 maps where the reference divides 0 by 0."""
    gen = torch.Generator().manual_seed(11)
    x = torch.rand(*shape, generator=gen).to(dt)
    want = O.tv_loss(x.float().double(), 1.25)              # fp64 reference value of the same arithmetic
    got = M.TVLoss(1.25)(x.cuda())
    if shape[2] == 1 or shape[3] == 1:
        assert torch.isnan(got) and torch.isnan(O.tv_loss(x.float(), 1.25))
        return
    assert float(got) == pytest.approx(float(want), rel=3e-6)
    terms = M.ops.tv_loss(x.cuda(), 1.25, with_terms=True).cpu()
    xf = x.double()
    assert float(terms[1]) == pytest.approx(float(((xf[:, :, 1:] - xf[:, :, :-1]) ** 2).sum()), rel=3e-6)
    assert float(terms[2]) == pytest.approx(float(((xf[..., 1:] - xf[..., :-1]) ** 2).sum()), rel=3e-6)


def test_tv_loss_linearity_and_constant(M):
    """Size-independent properties: TV(c) = 0 for a constant map, TV(s*x) = s^2 TV(x) (power-of-two s: exact)."""
    x = torch.rand(3, 2, 96, 160, device="cuda")
    tv = M.TVLoss()
    assert float(tv(torch.full_like(x, 0.37))) == 0.0
    assert float(tv(4.0 * x)) == pytest.approx(16.0 * float(tv(x)), rel=1e-7)


@pytest.mark.parametrize("tag", ["2layer", "3layer"])
def test_multilayer_network_golden(golden, M, tag):
    g = golden("multilayer_gat.npz")
    hp, wp, fin, hidden, fout, heads, nl = (int(v) for v in g[f"{tag}_meta"])
    net = M.StackedGATNetwork(fin, hidden, fout, heads, nl, 0.1, 0.2).cuda().eval()
    sd = {}
    for i in range(nl):
        W, a = g[f"{tag}_W{i}"], g[f"{tag}_a{i}"]
        for k in range(heads):
            sd[f"gat_layers.{i}.heads.{k}.W.weight"] = T(W[k])
            sd[f"gat_layers.{i}.heads.{k}.a.weight"] = T(a[k]).view(1, -1)
    net.load_state_dict(sd)
    x = T(g[f"{tag}_x"]).cuda()
    _, ei = M.PatchGraphConstructor(16).construct_patch_graph(torch.zeros(1, hp * 16, wp * 16), x)
    with torch.no_grad():
        y = net(x, ei)
    assert float((y.cpu() - T(g[f"{tag}_y"])).abs().max()) <= 1e-5
    # the default construction keeps the reference's widths and therefore its failure (graph_attention.py:176-186)
    bad = M.GATNetwork(fin, hidden, fout, heads, nl, 0.1, 0.2).cuda().eval()
    with pytest.raises(RuntimeError, match="cannot be multiplied"):
        bad(x, ei)


def test_multilayer_network_trains(M):
    """Gradients flow through the stacked layers (every parameter gets a finite, non-zero gradient)."""
    torch.manual_seed(0)
    net = M.StackedGATNetwork(12, 16, 8, 2, 3, 0.0, 0.2).cuda().train()
    x = torch.randn(30, 12, device="cuda", requires_grad=True)
    _, ei = M.PatchGraphConstructor(16).construct_patch_graph(torch.zeros(1, 80, 96), x)
    net(x, ei).square().sum().backward()
    for n_, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all() and float(p.grad.abs().sum()) > 0, n_
    assert torch.isfinite(x.grad).all()


@pytest.mark.parametrize("train", [False, True])
def test_block_feature_loss(M, train):
    """GraphBlock(f_unet_patches=, patch_labels_y=) returns the reference's L_feature on its own patch-GAT output
    (train_end_to_end.py:344) — fused-kernel inference path and autograd training path."""
    torch.manual_seed(5)
    blk = M.GraphBlock(dropout_rate=0.0).cuda()
    blk.train(train)
    B, H, W = 3, 96, 80
    N = (H // 16) * (W // 16)
    x = torch.randn(B, N, 20, device="cuda")
    fu = (0.3 * torch.randn(B, N, 64, device="cuda")).requires_grad_(train)
    y = torch.randint(0, 2, (B, N), device="cuda")
    with torch.set_grad_enabled(train):
        out = blk(x, (H, W), f_unet_patches=fu, patch_labels_y=y, want_dense=False)
    want = O.feature_consistency_loss(fu.detach().cpu(), out.patch_features.detach().cpu(), y.cpu())
    assert float(out.l_feature.detach()) == pytest.approx(float(want), rel=5e-6)
    if not train:
        # inference: the loss is evaluated INSIDE the one-launch block kernel (rows of h still in registers) — the step
        # launches exactly as many kernels of the library as without the loss, and agrees with the stand-alone kernel
        from mingraph_unet_b200 import _lib
        with torch.no_grad():
            n0 = _lib.launch_count(); blk(x, (H, W), want_dense=False)
            n1 = _lib.launch_count(); blk(x, (H, W), f_unet_patches=fu, patch_labels_y=y, want_dense=False)
            n2 = _lib.launch_count()
        assert n2 - n1 == n1 - n0 == 1
        alone = M.FeatureConsistencyLoss()(fu, out.patch_features, y)
        assert float(out.l_feature) == pytest.approx(float(alone), rel=2e-6)
        # other margins / label dtypes, a padded grid and K = 3
        blk3 = M.GraphBlock(node_feature_dim=12, num_segments=3, dropout_rate=0.0).cuda().eval()
        x3 = torch.randn(2, 5 * 5, 12, device="cuda")
        fu3 = 0.5 * torch.randn(2, 25, 64, device="cuda")
        y3 = torch.randint(0, 2, (2, 25), device="cuda").float()
        with torch.no_grad():
            o3 = blk3(x3, (70, 75), f_unet_patches=fu3, patch_labels_y=y3, feature_loss_margin=0.7, want_dense=False)
        w3 = O.feature_consistency_loss(fu3.cpu(), o3.patch_features.cpu(), y3.cpu(), margin=0.7)
        assert float(o3.l_feature) == pytest.approx(float(w3), rel=5e-6)
    if train:
        (out.l_feature + out.l_partition.mean()).backward()
        assert fu.grad is not None and torch.isfinite(fu.grad).all()
        g = blk.patch_gat_model.gat_layers[0].heads[0].W.weight.grad
        assert g is not None and float(g.abs().sum()) > 0
    with pytest.raises(ValueError, match="patch_labels_y is required"):
        blk(x, (H, W), f_unet_patches=fu, want_dense=False)
