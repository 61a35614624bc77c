"""Stub package: test infrastructure only (see oracle/README.md)."""
