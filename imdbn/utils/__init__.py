"""``imdbn.utils``: only the hot-path pieces are provided (conditional_steps, energy_utils)."""
