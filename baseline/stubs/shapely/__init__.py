"""Import-only stand-in for `shapely` (not installed here): the reference imports it at module level in
postprocess.py / datasets/util.py, which are OFF the training hot path. Nothing in it is ever called by the
paths this repo runs (synthetic tensors, no polygon post-processing); calling it raises."""
