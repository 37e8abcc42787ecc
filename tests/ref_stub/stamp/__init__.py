"""Stand-in for the STAMP package (tests/ref_stub/README.md)."""
