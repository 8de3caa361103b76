"""Mirror of the reference's `models` package (models/tadgan.py)."""
