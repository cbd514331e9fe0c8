"""Small helpers shared by the storage layer (reference numpywren/utils.py:7-46)."""


def convert_to_slice(l):
    """None | int | (stop,) | (start, stop) | (start, stop, step) → slice  (block-index slicing)."""
    if l is None:
        return slice(None, None, None)
    if isinstance(l, int):
        return slice(l, l + 1, 1)
    if isinstance(l, slice):
        raise ValueError("Could not convert to slice.")
    parts = list(l)
    if len(parts) == 1:
        return slice(None, parts[0], None)
    if len(parts) == 2:
        return slice(parts[0], parts[1], None)
    if len(parts) == 3:
        return slice(parts[0], parts[1], parts[2])
    raise ValueError("Expected slices of length 1 to 3.")


def remove_duplicates(items):
    out = []
    for x in items:
        if x not in out:
            out.append(x)
    return out


def chunk(l, n):
    """Split ``l`` into consecutive chunks of at most ``n`` items."""
    if n == 0:
        return []
    return [l[i:i + n] for i in range(0, len(l), n)]
