"""Fast-path stand-in for the reference's src/validate_blip.py compute_* entry points (Mode B): identical
signatures and return values (validate_blip.py:24-57,232-285), fused scan/top-k underneath."""
from sprc_b200.retrieval import compute_cirr_val_metrics, compute_fiq_val_metrics  # noqa: F401
