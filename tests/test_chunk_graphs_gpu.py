"""ChunkGraphEngine (mult_engine.py): the MulT step replayed as one captured CUDA graph per chunk and direction must equal the
eagerly issued step -- outputs bit for bit (same kernels, same order), parameter gradients up to fp32 summation order -- through
the static input buffers, with a modality mask, with a ragged last chunk, inside HierarchicalFusion, after an optimizer step,
and with dropout (fresh masks every step that forward and backward agree on).  Reference: models/fusion_layers.py:93-179."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("simple-multimodal_b200")
FL, ops, ME, K = pkg.fusion_layers, pkg.ops, pkg.mult_engine, pkg.kernels
LENS = (96, 64, 30)


class Cfg:
    def __init__(self, p=0.0):
        self.fusion_hidden_size, self.fusion_num_heads, self.fusion_dropout = 512, 8, p
        self.num_emotions, self.graph_hidden_size, self.graph_num_layers, self.graph_dropout = 7, 512, 3, p
        self.contrastive_temperature = 0.07


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def feats(B, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return [torch.randn(B, L, 512, device="cuda", generator=g).to(torch.bfloat16) for L in LENS]


def run(head, xs, mask=None, loss_w=None, **kw):
    for p in head.parameters():
        p.grad = None
    ins = [x.detach().clone().requires_grad_(True) for x in xs]
    out = head(*ins, mask=mask, **kw)
    f = out["fused_features"].float()
    loss = (f * f).mean() if loss_w is None else (f * loss_w).sum()
    for v in out.get("contrastive_losses", {}).values():
        loss = loss + 0.1 * v
    loss.backward()
    torch.cuda.synchronize()
    return out["fused_features"].detach().clone(), [i.grad.clone() for i in ins], {k: p.grad.clone() for k, p in head.named_parameters()}


def small_mult(p=0.0, chunk=2):
    torch.manual_seed(0)
    head = FL.MultimodalTransformer(Cfg(p)).cuda()
    head.train()
    head.chunk_size, head.graph_min_tokens = chunk, 0
    return head


def check_same(got, want, tol=2e-3):
    assert torch.equal(got[0], want[0])
    for a, b in zip(got[1], want[1]):
        assert torch.equal(a, b)
    for k in want[2]:
        assert rel(got[2][k], want[2][k]) < tol or float(want[2][k].norm()) < 1e-6, k


@pytest.mark.parametrize("B", [4, 5])            # 5: ragged last chunk (its own pair of graphs)
def test_engine_equals_eager(B):
    head = small_mult()
    mask = torch.tensor([[1, 1, 1], [0, 1, 1], [1, 0, 0], [1, 1, 0], [0, 0, 1]], device="cuda", dtype=torch.float32)[:B]
    xs, ys = feats(B, 1), feats(B, 2)
    head.graph_chunks = False
    want_x, want_y, want_m = run(head, xs), run(head, ys), run(head, xs, mask)
    head.graph_chunks = True
    got_x = run(head, xs)
    eng = head._engine
    assert eng is not None and len(eng.fwd_graphs) == (B + 1) // 2 and eng.kernel_launches > 100
    check_same(got_x, want_x)
    check_same(run(head, ys), want_y)                 # new data goes through the static input buffers
    check_same(run(head, xs, mask), want_m)           # modality mask folded into the staging copy / the input gradients
    check_same(run(head, xs), want_x)                 # and no state leaks from one step to the next
    assert head._engine is eng and head._engine_builds == 1
    # a dropped modality gets exactly zero input gradient
    assert float(run(head, xs, mask)[1][0][1].abs().max()) == 0.0
    head.release_graphs()


def test_engine_sees_updated_weights_and_grad_accumulation():
    head = small_mult()
    xs = feats(4, 3)
    run(head, xs)
    opt = pkg.FusedAdamW(head.parameters(), lr=1e-2)
    opt.step()                                        # writes the parameters through raw pointers: the version bump refreshes the operand copies
    got = run(head, xs)
    head.graph_chunks = False
    head.release_graphs()
    check_same(got, run(head, xs))
    # gradients returned by the engine are copied into .grad, never adopted: a second backward accumulates
    head.graph_chunks = True
    g1 = run(head, xs)[2]
    ins = [x.detach().clone().requires_grad_(True) for x in xs]
    (head(*ins)["fused_features"].float() ** 2).mean().backward()        # no zeroing in between
    torch.cuda.synchronize()
    for k, p in head.named_parameters():
        assert rel(p.grad, 2 * g1[k]) < 2e-3 or float(g1[k].norm()) < 1e-6, k
    head.release_graphs()


def test_only_one_forward_in_flight():
    head = small_mult()
    xs = feats(4, 4)
    a = head(*[x.requires_grad_(True) for x in xs])["fused_features"].float().sum()
    b = head(*xs)["fused_features"].float().sum()
    b.backward()
    with pytest.raises(pkg.B200FusionError, match="one forward may be in flight"):
        a.backward()
    head.release_graphs()


def test_engine_dropout_matches_eager_at_the_same_seeds_and_epoch():
    """forward and backward graphs regenerate one and the same mask per step: the replayed step equals an eagerly issued step
    that is given the engine's frozen seeds under the epoch of that step; consecutive steps draw different masks."""
    head = small_mult(p=0.1)
    xs = feats(4, 5)
    w = torch.randn(4, 512, device="cuda")
    s1 = run(head, xs, loss_w=w)
    s2 = run(head, xs, loss_w=w)
    eng = head._engine
    assert eng is not None and eng.step_no == 2
    assert not torch.equal(s1[0], s2[0])                                  # fresh masks every step
    assert all(torch.isfinite(g).all() for g in s2[2].values())
    W, plist = head._operands(torch.bfloat16)
    ins = [x.detach().clone().requires_grad_(True) for x in xs]
    for p in head.parameters():
        p.grad = None
    try:
        K.dropout_epoch(2)                                                # the epoch the engine raised around its second step
        pooled, _ = ME.MulTFn.apply(*ins, None, False, 512, 8, 2, 1 << 40, eng.drop, head._names, W, None, *plist)
        # (final_fusion's own elementwise dropout draws from the host-side counter stream: compare the pooled path only)
        pooled.float().pow(2).mean().backward()
        torch.cuda.synchronize()
    finally:
        K.dropout_epoch(0)
    want_dx = [i.grad.clone() for i in ins]
    want_pg = {n: p.grad.clone() for n, p in zip(head._names, plist)}
    for p in head.parameters():
        p.grad = None
    ins2 = [x.detach().clone().requires_grad_(True) for x in xs]
    eng.step_no = 1                                                       # replay step 2 again
    pooled2, _ = ME.MulTFn.apply(*ins2, None, False, 512, 8, 2, 0, None, head._names, W, eng, *plist)
    pooled2.float().pow(2).mean().backward()
    torch.cuda.synchronize()
    assert torch.equal(pooled2, pooled)
    for a, b in zip(ins2, want_dx):
        assert torch.equal(a.grad, b)
    for n, p in zip(head._names, plist):
        assert rel(p.grad, want_pg[n]) < 2e-3 or float(want_pg[n].norm()) < 1e-6, n
    head.release_graphs()


def test_hierarchical_with_engine_equals_eager():
    torch.manual_seed(1)
    head = FL.HierarchicalFusion(Cfg()).cuda()
    head.train()
    head.mult_fusion.chunk_size, head.mult_fusion.graph_min_tokens = 3, 0
    xs = feats(7, 6)
    mask = pkg.ModalityDropout(0.3, seed=5).sample_mask(7, "cuda")
    head.mult_fusion.graph_chunks = False
    want = run(head, xs, mask, compute_contrastive_loss=True)
    head.mult_fusion.graph_chunks = True
    got = run(head, xs, mask, compute_contrastive_loss=True)
    assert head.mult_fusion._engine is not None
    check_same(got, want)
    head.mult_fusion.release_graphs()


def test_engine_is_bypassed_where_it_does_not_apply():
    head = small_mult()
    xs = feats(4, 7)
    with torch.no_grad():
        head(*xs)
    assert head._engine is None                                           # inference
    head(*[x.float() for x in xs])
    assert head._engine is None                                           # fp32 parity mode
    head.graph_min_tokens = FL.MultimodalTransformer.graph_min_tokens
    head(*xs)
    assert head._engine is None                                           # chunks too small for launch overhead to matter
    head.graph_min_tokens, head.stash_fraction = 0, 1e-9
    head(*xs)
    assert head._engine is None                                           # the stash may not stay resident: eager with recompute
    head.stash_fraction = FL.MultimodalTransformer.stash_fraction
    step = pkg.GraphedTrainStep(head, xs, lambda out: (out["fused_features"].float() ** 2).mean())
    assert head._engine is None and step.kernel_launches > 100            # inside an outer capture the chunks are part of that graph
