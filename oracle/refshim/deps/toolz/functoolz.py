import functools
import inspect


def identity(x):
    return x


def complement(f):
    return lambda *a, **k: not f(*a, **k)


def compose(*funcs):
    def composed(*a, **k):
        fs = list(reversed(funcs))
        out = fs[0](*a, **k)
        for f in fs[1:]:
            out = f(out)
        return out

    return composed


class curry:
    def __init__(self, func, *args, **kwargs):
        self.func = func
        self.args = args
        self.keywords = kwargs
        functools.update_wrapper(self, func, updated=())

    def __call__(self, *args, **kwargs):
        a = self.args + args
        k = dict(self.keywords, **kwargs)
        try:
            sig = inspect.signature(self.func)
            sig.bind(*a, **k)
        except TypeError:
            return curry(self.func, *a, **k)
        except ValueError:
            pass
        return self.func(*a, **k)

    def __get__(self, instance, owner):
        if instance is None:
            return self
        return curry(self, instance)
