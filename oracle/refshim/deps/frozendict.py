"""Stand-in for frozendict: a hashable dict."""


class frozendict(dict):
    def __hash__(self):
        return hash(tuple(sorted(self.items(), key=repr)))

    def _ro(self, *a, **k):
        raise TypeError("frozendict is immutable")

    __setitem__ = __delitem__ = _ro
