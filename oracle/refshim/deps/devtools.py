"""Stand-in for devtools."""
import pprint


def debug(*a, **k):
    return a[0] if a else None


def pformat(obj, **k):
    return pprint.pformat(obj)
