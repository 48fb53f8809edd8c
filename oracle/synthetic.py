"""Seeded synthetic inputs: kept here as the import point the tests and golden scripts use; the
generators themselves live in synthetic_case.py at the repository root (shared with bench.py, which must not import
the oracle)."""
from synthetic_case import case_volume, label_pair, label_volume, mri_volumes  # noqa: F401
