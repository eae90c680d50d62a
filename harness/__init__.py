"""Synthetic-data harness for tests/ and bench.py (not product code; never imported by avatar_b200)."""
