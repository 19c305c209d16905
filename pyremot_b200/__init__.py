"""pyremot_b200 — B200-native batched solver for PyREMOT's pseudo-homogeneous packed-bed path (N1, N2)."""
