"""Empty stand-in so the reference's `from matplotlib import pyplot` (radae/dsp.py:35) imports in a
container without matplotlib.  Used only by tools/ scripts that import /root/reference."""
