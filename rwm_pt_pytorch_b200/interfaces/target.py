"""`TargetDistribution` base (reference: interfaces/target.py:1-14)."""


class TargetDistribution:
    """General interface for (NumPy-style) target distributions."""

    def __init__(self, dimension):
        self.dim = dimension

    def density(self, x):
        raise NotImplementedError("Subclasses must implement the density method.")

    def draw_sample(self, beta=1.0):
        raise NotImplementedError("Subclasses must implement the draw_sample method.")
