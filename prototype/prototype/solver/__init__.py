"""Evaluation command lines of the hot path (same module paths and flags as the reference's
prototype/prototype/solver/*.py), backed by robustart_b200.solver."""
