import itertools

no_default = "__no__default__"
no_pad = "__no__pad__"


def unique(seq, key=None):
    seen = set()
    for item in seq:
        val = item if key is None else key(item)
        if val not in seen:
            seen.add(val)
            yield item


def _getter(index):
    if isinstance(index, list):
        return lambda x: tuple(x[i] for i in index)
    return lambda x: x[index]


def pluck(ind, seqs, default=no_default):
    if default == no_default:
        get = _getter(ind)
        return map(get, seqs)
    if isinstance(ind, list):
        return (tuple(_get(item, seq, default) for item in ind) for seq in seqs)
    return (_get(ind, seq, default) for seq in seqs)


def _get(ind, seq, default):
    try:
        return seq[ind]
    except (KeyError, IndexError):
        return default


def partition(n, seq, pad=no_pad):
    args = [iter(seq)] * n
    if pad is no_pad:
        return zip(*args)
    return itertools.zip_longest(*args, fillvalue=pad)


def partition_all(n, seq):
    it = iter(seq)
    while True:
        chunk = tuple(itertools.islice(it, n))
        if not chunk:
            return
        yield chunk


def groupby(key, seq):
    if not callable(key):
        key = _getter(key)
    out = {}
    for item in seq:
        out.setdefault(key(item), []).append(item)
    return out


def reduceby(key, binop, seq, init=no_default):
    if not callable(key):
        key = _getter(key)
    out = {}
    for item in seq:
        k = key(item)
        if k not in out:
            if init == no_default:
                out[k] = item
                continue
            out[k] = init() if callable(init) else init
        out[k] = binop(out[k], item)
    return out


def take_nth(n, seq):
    return itertools.islice(seq, 0, None, n)


def diff(*seqs, **kwargs):
    N = len(seqs)
    if N == 1 and isinstance(seqs[0], list):
        seqs = seqs[0]
        N = len(seqs)
    default = kwargs.get("default", no_default)
    key = kwargs.get("key", None)
    if default == no_default:
        iters = zip(*seqs)
    else:
        iters = itertools.zip_longest(*seqs, fillvalue=default)
    for items in iters:
        vals = items if key is None else tuple(map(key, items))
        if vals.count(vals[0]) != N:
            yield items
