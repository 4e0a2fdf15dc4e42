from coda_b200.base import ModelSelector  # noqa: F401  (reference coda/base.py)
