class classproperty:
    def __init__(self, fn):
        self.fn = fn

    def __get__(self, obj, owner):
        return self.fn(owner)


def make_sentinel(name="_MISSING", var_name=None):
    class Sentinel:
        def __init__(self):
            self.name = name
            self.var_name = var_name

        def __repr__(self):
            return self.var_name or "%s(%r)" % (type(self).__name__, self.name)

        def __bool__(self):
            return False

        def __copy__(self):
            return self

        def __deepcopy__(self, memo):
            return self

    return Sentinel()


def issubclass(subclass, baseclass):  # noqa: A001
    import builtins

    try:
        return builtins.issubclass(subclass, baseclass)
    except TypeError:
        return False


def get_all_subclasses(cls):
    out, todo = [], list(cls.__subclasses__())
    while todo:
        c = todo.pop()
        if c not in out:
            out.append(c)
            todo.extend(c.__subclasses__())
    return out
