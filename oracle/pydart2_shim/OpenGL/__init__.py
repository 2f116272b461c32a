"""Stub: lets the unmodified reference env modules import without GL (test infrastructure)."""
