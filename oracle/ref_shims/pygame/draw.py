"""no-op stand-in; never called on the hot path."""
