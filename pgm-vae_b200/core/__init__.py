"""Reference-facing Python surface (mirrors the reference's ``core`` package)."""
