"""Stand-in: the reference does `from matplotlib.pyplot import get` (M.py:5) and never calls it."""
