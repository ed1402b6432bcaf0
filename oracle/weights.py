"""Name-keyed deterministic weights: moved to realcamnet_b200/synthetic.py (shared with bench.py and the tools, which must not import
the oracle package); re-exported here for the fixtures and tests."""
from realcamnet_b200.synthetic import _SKIP, _is_fixed_dwt, _rng, checksum, fill_, make_tensor  # noqa: F401
