from coda_b200.datasets import Dataset  # noqa: F401  (reference coda/datasets.py)
