import re
import textwrap
import unicodedata


def a10n(string):
    if len(string) < 3:
        return string
    return "%s%s%s" % (string[0], len(string[1:-1]), string[-1])


def asciify(text, ignore=False):
    if isinstance(text, bytes):
        text = text.decode("utf-8", "ignore")
    return unicodedata.normalize("NFKD", text).encode("ascii", "ignore" if ignore else "replace")


def slugify(text, delim="_", lower=True, ascii=False):  # noqa: A002
    parts = re.split(r"[^\w]+", text)
    out = delim.join(p for p in parts if p)
    return out.lower() if lower else out


def camel2under(s):
    return re.sub(r"((?<=[a-z0-9])[A-Z]|(?!^)[A-Z](?=[a-z]))", r"_\1", s).lower()


def under2camel(s):
    return "".join(w.capitalize() or "_" for w in s.split("_"))


def unwrap_text(text, ending="\n\n"):
    paras = re.split(r"\n\s*\n", textwrap.dedent(text).strip())
    out = [" ".join(line.strip() for line in p.splitlines()) for p in paras]
    return ending.join(out) if ending is not None else out


def iter_splitlines(text):
    yield from text.splitlines()


def format_int_list(int_list, delim=",", range_delim="-", delim_space=False):
    return (delim + (" " if delim_space else "")).join(str(i) for i in int_list)


def parse_int_list(range_string, delim=",", range_delim="-"):
    out = []
    for part in range_string.split(delim):
        part = part.strip()
        if not part:
            continue
        if range_delim in part[1:]:
            a, b = part.split(range_delim, 1)
            out.extend(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return sorted(out)
