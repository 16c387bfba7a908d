"""Drop-in for the reference's ``utils`` package (same module and symbol names)."""
