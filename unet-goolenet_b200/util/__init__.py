"""Counterpart of the reference's `util` package for the hot path (分类/util/roi.py)."""
