"""Measurement scripts that use the oracle as their checker (not collected by pytest: no test_ prefix).  They live under
tests/ because only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import oracle/."""
