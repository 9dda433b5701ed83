def is_iterable(obj):
    try:
        iter(obj)
    except TypeError:
        return False
    return True


def is_scalar(obj):
    return not is_iterable(obj) or isinstance(obj, (str, bytes))


def is_collection(obj):
    return is_iterable(obj) and not isinstance(obj, (str, bytes))


def flatten_iter(iterable):
    for item in iterable:
        if isinstance(item, (list, tuple, set, frozenset)) or (
            is_iterable(item) and not isinstance(item, (str, bytes, dict))
        ):
            yield from flatten_iter(item)
        else:
            yield item


def flatten(iterable):
    return list(flatten_iter(iterable))
