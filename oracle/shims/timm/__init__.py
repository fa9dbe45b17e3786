"""Stand-in for the nine ``timm~=1.0.15`` symbols used by models/mirror.py:29-39.
TEST INFRASTRUCTURE ONLY (see ../README.md)."""
