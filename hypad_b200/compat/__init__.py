"""Import-name compatibility for reference checkpoints (whole-module pickles)."""
