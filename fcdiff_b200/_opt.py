"""
The bounded minimiser of the (eta, epsilon) sub-problem.

``scipy.optimize.minimize(method="L-BFGS-B")`` is what the reference calls
(fcdiff/fit.py:239-241) and what the oracle uses.  Its Python front end
(``ScalarFunction``, bounds standardisation, memoisation) costs ~90 us per
objective evaluation -- as much as the CUDA kernel that computes the objective.
``minimize_lbfgsb`` drives the same compiled routine (``_lbfgsb.setulb``) with
the same parameters as ``_minimize_lbfgsb`` (scipy/optimize/_lbfgsb_py.py), so
the iterates are bit-identical, and falls back to ``scipy.optimize.minimize``
if this SciPy does not have the expected private interface.
"""
import numpy as np
import scipy.optimize

try:
    from scipy.optimize import _lbfgsb
    from scipy.optimize._lbfgsb_py import HAS_ILP64 as _ILP64
except Exception:                       # pragma: no cover - depends on the SciPy build
    _lbfgsb = None
    _ILP64 = False

_direct_ok = _lbfgsb is not None and hasattr(_lbfgsb, "setulb")


class Result(object):
    def __init__(self, x, fun, nfev, nit):
        self.x = x
        self.fun = fun
        self.nfev = nfev
        self.nit = nit


def _minimize_direct(fun, x0, lower, upper, m=10, ftol=2.2204460492503131e-09, gtol=1e-5, maxfun=15000,
                     maxiter=15000, maxls=20):
    """scipy/optimize/_lbfgsb_py.py:_minimize_lbfgsb without the wrappers."""
    n = len(x0)
    int_dtype = np.int64 if _ILP64 else np.int32
    factr = ftol / np.finfo(float).eps
    nbd = np.full(n, 2, dtype=int_dtype)                      # both bounds finite
    low_bnd = np.array(lower, dtype=np.float64)
    upper_bnd = np.array(upper, dtype=np.float64)
    x = np.clip(np.array(x0, dtype=np.float64), low_bnd, upper_bnd)
    f = np.array(0.0, dtype=np.float64)
    g = np.zeros((n,), dtype=np.float64)
    wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m, np.float64)
    iwa = np.zeros(3 * n, dtype=int_dtype)
    task = np.zeros(2, dtype=int_dtype)
    ln_task = np.zeros(2, dtype=int_dtype)
    lsave = np.zeros(4, dtype=int_dtype)
    isave = np.zeros(44, dtype=int_dtype)
    dsave = np.zeros(29, dtype=np.float64)
    nfev = 0
    nit = 0
    while True:
        _lbfgsb.setulb(m, x, low_bnd, upper_bnd, nbd, f, g, factr, gtol, wa, iwa, task, lsave, isave, dsave,
                       maxls, ln_task)
        if task[0] == 3:
            (fv, gv) = fun(x)
            nfev += 1
            f[...] = fv
            g[...] = gv
        elif task[0] == 1:
            nit += 1
            if nit >= maxiter:
                task[0] = 5
                task[1] = 504
            elif nfev > maxfun:
                task[0] = 5
                task[1] = 502
        else:
            break
    return Result(x, float(f), nfev, nit)


def _probe_direct():
    """One solve of a tiny bounded quadratic through the private interface, compared with the public
    ``scipy.optimize.minimize``: any exception (another ``setulb`` signature, dtype or shape checks of
    another SciPy generation) or a different minimiser selects the public path for the whole process.
    Developed against SciPy 1.15-1.18 (``setulb(m, x, l, u, nbd, f, g, factr, pgtol, wa, iwa, task,
    lsave, isave, dsave, maxls, ln_task)``)."""
    if not _direct_ok:
        return False

    def quad(x):
        d = np.asarray(x) - np.array([0.3, 2.0])
        return float(d @ d), 2.0 * d

    try:
        r = _minimize_direct(quad, [0.9, 0.1], [0.0, 0.0], [1.0, 1.0])
        ref = scipy.optimize.minimize(quad, np.array([0.9, 0.1]), jac=True, method="L-BFGS-B",
                                      bounds=[(0.0, 1.0), (0.0, 1.0)])
        return bool(np.array_equal(r.x, ref.x) and r.nfev == ref.nfev)
    except Exception:
        return False


_direct_ok = _probe_direct()


def minimize_lbfgsb(fun, x0, lower, upper):
    """Minimises ``fun(x) -> (f, grad)`` on the box [lower, upper]."""
    if _direct_ok:
        return _minimize_direct(fun, x0, lower, upper)
    nfev = [0]

    def wrapped(x):
        nfev[0] += 1
        return fun(x)

    r = scipy.optimize.minimize(wrapped, np.asarray(x0, dtype=np.float64), jac=True, method="L-BFGS-B",
                                bounds=list(zip(lower, upper)))
    return Result(r.x, float(r.fun), nfev[0], int(r.nit))
