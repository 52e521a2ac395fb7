"""Stand-in for the two names the reference imports from timm (M.py:22)."""
