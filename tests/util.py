import numpy as np


def bits_equal(a, b):
    """Bit-for-bit equality of two float32 arrays, NaN == NaN of any payload."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    if a.shape != b.shape:
        return False
    same = a.view(np.uint32) == b.view(np.uint32)
    same |= np.isnan(a) & np.isnan(b)
    return bool(same.all())


def diff_report(a, b):
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.abs(a - b)
    d[both_nan] = 0
    d = np.nan_to_num(d, nan=np.inf)
    ne = (a.view(np.uint32) != b.view(np.uint32)) & ~both_nan
    return "max|d|=%g  n(bits differ)=%d of %d  n(>1e-4)=%d" % (d.max() if d.size else 0, int(ne.sum()), a.size,
                                                              int((d > 1e-4).sum()))
