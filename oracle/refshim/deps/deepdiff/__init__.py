"""Stand-in for deepdiff."""


class DeepDiff(dict):
    def __init__(self, a=None, b=None, **k):
        super().__init__()
        if a != b:
            self["values_changed"] = (a, b)
