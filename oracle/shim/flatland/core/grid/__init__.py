"""oracle shim package (see oracle/shim/README.md)."""
