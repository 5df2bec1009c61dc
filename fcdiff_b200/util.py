"""
Index arithmetic between regions and connections (edges).

Same names, arguments and conventions as the reference's ``fcdiff/util.py``:
edges are in lower-triangular row-major order, ``c = n(n-1)/2 + m`` with
``m < n`` (fcdiff/util.py:40-84, pinned by test_fcdiff/test_util.py:14-36).
Unlike the Python-2 reference these return exact integers;
``C_to_N`` returns a float like the reference so that ``N % 1`` can flag a
non-triangular ``C`` (fcdiff/fit.py:62-65).  The device kernels use the same
maps (``c_to_nm`` in csrc/fcd_common.cuh).
"""
import math


def N_to_C(N):
    """Number of connections of a network with N regions (fcdiff/util.py:7-21)."""
    return N * (N - 1) // 2


def C_to_N(C):
    """Number of regions of a network with C connections (fcdiff/util.py:23-38)."""
    return (math.sqrt(8 * C + 1) - 1) / 2 + 1


def nm_to_c(n, m):
    """Connection index of the region pair (n, m) (fcdiff/util.py:40-60)."""
    return N_to_C(n) + m


def c_to_nm(c):
    """Region pair (n, m), m < n, of connection c (fcdiff/util.py:62-84)."""
    c = int(c)
    n = (math.isqrt(8 * c + 1) - 1) // 2 + 1
    m = c - N_to_C(n)
    return (n, m)
