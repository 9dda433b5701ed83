"""Stand-in for mako: constructible, never rendered by the numpy backend."""


class Template:
    def __init__(self, *a, **k):
        pass

    def render(self, *a, **k):
        raise RuntimeError("mako stand-in cannot render")
