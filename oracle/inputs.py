"""Seeded synthetic inputs: moved to realcamnet_b200/synthetic.py (shared with bench.py and the tools, which must not import the
oracle package); re-exported here for the fixtures and tests."""
from realcamnet_b200.synthetic import coord_map, make_inputs  # noqa: F401
