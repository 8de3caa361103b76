"""Mirror of the reference's `hyperspace` package for the scoring path (hyrnn_nets, poincare_distance)."""
