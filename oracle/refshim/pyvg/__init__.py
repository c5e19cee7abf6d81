"""TEST INFRASTRUCTURE ONLY -- empty stand-in for pyvg (I/O, off the hot path)."""
