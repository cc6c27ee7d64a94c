"""Stub of the absent third-party dev-tool package `limedev` (only its CLI shim is imported by the reference, never on the integration path)."""
