"""Stand-in for ``more_itertools.collapse`` (documented behaviour: flatten
nested iterables depth-first; ``str``/``bytes`` and ``base_type`` instances
are treated as atoms)."""


def collapse(iterable, base_type=None, levels=None):
    def walk(node, level):
        if (
            ((levels is not None) and (level > levels))
            or isinstance(node, (str, bytes))
            or ((base_type is not None) and isinstance(node, base_type))
        ):
            yield node
            return
        try:
            tree = iter(node)
        except TypeError:
            yield node
            return
        for child in tree:
            yield from walk(child, level + 1)

    yield from walk(iterable, 0)
