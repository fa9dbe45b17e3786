"""Test infrastructure: CPU oracle for the MIRROR hot path (see mirror_oracle.py)."""
