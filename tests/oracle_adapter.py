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


class OracleInnerWithFilters(OracleInner):
    """CPU stand-in offering the push-down surface of the B200 index (filter_create / filter_set / search_filtered),
    so the host logic of DerivedFilteredSearchIndex(pushdown=True) can be tested without a GPU: a filtered query is
    answered by searching everything and keeping the first `limit` rows whose bit is set."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.filters: dict[int, dict[int, bool]] = {}
        self.created = 0

    def filter_create(self):
        self.created += 1
        self.filters[self.created] = {}
        return self.created

    def filter_destroy(self, fid):
        del self.filters[fid]

    def filter_set(self, fid, keys, passes):
        self.filters[fid].update(zip(keys, passes))

    def remove(self, keys):
        for f in self.filters.values():
            for k in keys:
                f.pop(k, None)
        return super().remove(keys)

    def search_filtered(self, queries):
        out = []
        for key, vec, lim, fid in queries:
            (_k, res), = self.search([(key, vec, self.current_size)])
            if fid is not None:
                res = [(m, s) for m, s in res if self.filters[fid].get(m, False)]
            out.append((key, res[:lim]))
        return out
