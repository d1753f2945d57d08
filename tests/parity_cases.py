"""The parity-test matrix, shared by tests/test_parity_gpu.py and oracle/make_bf16_floor.py.
id -> dict(kind, cfg kwargs, B, lens, flag, mask_seed, chunk)."""
N3 = (None, None, None)


def _c(kind, B, lens=N3, flag=False, mask_seed=None, chunk=None, fp32_only=False, **cfg):
    return dict(kind=kind, B=B, lens=lens, flag=flag, mask_seed=mask_seed, chunk=chunk, fp32_only=fp32_only, cfg=cfg)


CASES = {
    "early_b1": _c("early", 1), "early_b16": _c("early", 16), "early_b257": _c("early", 257),
    "late_b33": _c("late", 33),
    "graph_b2": _c("graph", 2), "graph_b130": _c("graph", 130),
    "graph_1layer_256": _c("graph", 9, graph_hidden=256, graph_layers=1),
    "contrastive_b2": _c("contrastive", 2, flag=True), "contrastive_b64": _c("contrastive", 64, flag=True),
    "contrastive_b1000": _c("contrastive", 1000, flag=True), "contrastive_noloss": _c("contrastive", 8),
    "adaptive_b1": _c("adaptive", 1), "adaptive_b70": _c("adaptive", 70),
    # narrow rows (head dim 8): the shapes of the golden fixtures; fp32 only -- no bf16 floor was measured at these widths
    "adaptive_h64": _c("adaptive", 5, fp32_only=True, H=64, graph_hidden=64),
    "hier_2d_h32": _c("hierarchical", 4, flag=True, fp32_only=True, H=32, heads=4, graph_hidden=32),
    "mult_2d": _c("mult", 6),
    "mult_3d_111": _c("mult", 3, (1, 1, 1)), "mult_3d_64_64_30": _c("mult", 3, (64, 64, 30)),
    "mult_3d_33_65_7": _c("mult", 3, (33, 65, 7)), "mult_chunked": _c("mult", 5, (40, 24, 30), chunk=2),
    "hier_2d": _c("hierarchical", 12, flag=True),
    "hier_3d_mask": _c("hierarchical", 6, (48, 32, 30), flag=True, mask_seed=4321, chunk=4),
}
