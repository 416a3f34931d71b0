"""Adapters that give the CPU oracles the trait-shaped surface of the product's innermost index."""

from oracle import brute_force_oracle as bfo
from oracle.c_oracle import CBruteForceKNNIndex
from pathway_b200.external_index import MissingKey


class OracleInner:
    def __init__(self, dimensions, reserved_space, auxiliary_space, metric, impl="numpy"):
        m = metric.value if hasattr(metric, "value") else int(metric)
        cls = bfo.BruteForceKNNIndex if impl == "numpy" else CBruteForceKNNIndex
        self.ix = cls(dimensions, reserved_space, auxiliary_space, m)
        self.dimensions = dimensions

    @property
    def current_size(self):
        return self.ix.current_size

    @property
    def current_allocated(self):
        return self.ix.current_allocated

    def add(self, add_data):
        return [(k, None) for k, _ in self.ix.add([(k, list(v)) for k, v in add_data])] if add_data else []

    def remove(self, keys):
        return [(k, None if e is None else MissingKey(k)) for k, e in self.ix.remove(list(keys))] if keys else []

    def search(self, queries):
        return self.ix.search([(k, list(v), lim) for k, v, lim in queries]) if queries else []
