"""Mirror of the reference's `utils` package for the scoring path (anomaly_detection_utils, dataloader)."""
