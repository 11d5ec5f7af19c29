"""Deterministic synthetic weights: re-exported from tqdne_b200.synthetic_weights (the generator is not a restatement of
reference arithmetic, and bench.py / the tools need it without importing the oracle package)."""
from tqdne_b200.synthetic_weights import seeded_state_dict, shapes_of  # noqa: F401
