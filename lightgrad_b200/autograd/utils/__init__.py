"""gradcheck and profiler helpers (import the sub-modules explicitly, as with the reference)."""
