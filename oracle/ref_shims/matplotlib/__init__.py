"""empty stand-in; plotting is never called on the hot path."""
