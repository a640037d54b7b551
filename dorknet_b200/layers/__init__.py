"""B200 mirror of the reference's `layers` package (same module and class names)."""
