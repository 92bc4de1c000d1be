"""Helpers shared by the fa1 / fa2 / fa3 entry points (mirrors the reference's ``src/common`` for the hot path only)."""
