from coda_b200.oracle import Oracle  # noqa: F401  (reference coda/oracle.py)
