"""Mirror of the reference's ``shared`` package interface (custom layers + depth rules)."""
