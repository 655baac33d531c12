"""Drop-in replacements for the reference's `source` package (hot path only)."""
