"""Minimal stand-in for Shapely 1.8 (test infrastructure; see ../README.md)."""
