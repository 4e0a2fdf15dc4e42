from coda_b200.selector import CODA  # noqa: F401  (reference coda/coda.py:171)
