"""Model dimensions of the text-to-semantic golden fixtures (shared by make_golden.py, which needs /root/reference, and the tests,
which must not import it)."""
T2S_CONFIGS = {
    # small: fast CPU fixture; base: configuration.py defaults (hidden 512, 16 heads of 32); train: train_config.yaml (hidden 384,
    # 8 heads of 48, depth 12)
    "small": dict(hidden=128, heads=4, depth=2, lp_heads=4, lp_depth=1),
    "base": dict(hidden=512, heads=16, depth=8, lp_heads=16, lp_depth=4),
    "train": dict(hidden=384, heads=8, depth=12, lp_heads=8, lp_depth=4),
}
