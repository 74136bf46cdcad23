"""Mirror of the reference's ``codes/models`` import surface for the hot path only."""
