"""Empty stand-in for xformers.ops (import-only in the reference)."""
