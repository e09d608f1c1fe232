"""`MHAlgorithm` base class (reference: interfaces/metropolis.py:7-99).

Holds the chain list and implements the reference's initial-state rule, extended to batches of chains:
the rule is applied per chain, drawing from NumPy's global RNG exactly like the reference does (before the
harness seeds it, interfaces/simulation_gpu.py:143-148)."""
from __future__ import annotations

import numpy as np


def _target_name(target_dist) -> str:
    name = getattr(target_dist, 'name', None)
    if isinstance(name, str):
        return name
    get = getattr(target_dist, 'get_name', None)
    if callable(get):
        n = get()
        if isinstance(n, str):
            return n
    return ""


def initial_states(target_dist, dim: int, n: int = 1) -> np.ndarray:
    """(n, dim) initial points by target *name* (interfaces/metropolis.py:21-64): Beta -> U(0.2, 0.8) float32;
    Gamma -> 5 + 0.01 N(0,1); RoughCarpet / ThreeMixture -> zeros; anything else -> 1e-8 N(0,1)."""
    name = _target_name(target_dist)
    if "Beta" in name:
        return np.random.uniform(0.2, 0.8, size=(n, dim)).astype(np.float32)
    if "Gamma" in name:
        return 5 + 0.01 * np.random.randn(n, dim)
    if "RoughCarpet" in name or "ThreeMixture" in name:
        return np.zeros((n, dim))
    return 0.00000001 * np.random.randn(n, dim)


class MHAlgorithm:
    """General purpose Metropolis-Hastings interface; `chain[-1]` is the current state of chain 0."""

    def __init__(self, dim, var, target_dist=None, symmetric=True, num_chains: int = 1):
        self.dim = dim
        self.var = var
        self.target_dist = target_dist
        self._x0 = initial_states(target_dist, dim, max(1, int(num_chains)))
        self.chain = [self._x0[0]]
        self.symmetric = symmetric
        self.num_acceptances = 0
        self.acceptance_rate = 0
        self.target_density = getattr(target_dist, 'density', None) if target_dist is not None else None

    def reset(self):
        self.chain = [self.chain[0]]

    def step(self):
        raise NotImplementedError("Step method must be implemented in subclass")

    def get_curr_state(self):
        return self.chain[-1]

    def set_curr_state(self, state):
        self.chain[-1] = state

    def get_name(self):
        raise NotImplementedError("Subclasses must implement the get_name method.")
