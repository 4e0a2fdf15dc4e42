"""Empty stand-in: coda/util.py:2 of the reference imports matplotlib.pyplot for a debug plot nobody calls."""
