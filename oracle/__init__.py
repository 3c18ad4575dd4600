"""Test infrastructure only -- see oracle/rrdb_oracle.py."""
