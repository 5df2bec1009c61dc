"""
Probabilistic models that generate correlations.

Mirror of the reference's ``fcdiff/model.py``: ``UnsharedRegionModel`` keeps its
attributes, method names, argument shapes and return types
(fcdiff/model.py:31-236).  Sampling runs on the GPU with a counter-based
Philox4x32-10 generator (``csrc/fcd_sample.cu``); NumPy's MT19937 stream is not
reproduced, so parity with the reference is distributional
(test_fcdiff/test_model.py:37-245).  ``self.rng`` is only used to draw the
64-bit Philox key of each ``sample*`` call, so re-seeding ``rng`` re-seeds the
sampler and successive calls give fresh draws.

Edge order: ``util`` order (``c = n(n-1)/2 + m``, ``m < n``) everywhere.  The
reference's ``sample_T`` enumerates edges upper-triangular row-major
(fcdiff/model.py:132-142), which is inconsistent with ``fcdiff.fit``
(SURVEY 0.3); this implementation uses the order ``fit`` assumes.
"""
import textwrap

import numpy as np
import torch

from . import _dev, _lib
from .util import N_to_C


class UnsharedRegionModel(object):
    """
    The individual anomalous region (IAR) model.

    Attributes
    ----------
    rng : :class:`numpy.random.RandomState`
        Random number generator (seeds the device-side Philox streams).
    pi : 0 <= float <= 1
        Probability of an anomalous region.
    eta : 0 <= float <= 1
        Probability of an anomalous connection btw a typical and anomalous region.
    gamma : :class:`numpy.ndarray`, (3,), 0 <= float <= 1
        Probability of each template connection type (array sums to 1).
    epsilon : 0 <= float <= 1
        Probability that a typical connection differs from the template.
    mu : :class:`numpy.ndarray`, (3,), -1 <= float <= 1
        Mean correlation of each connection type.
    sigma : :class:`numpy.ndarray`, (3,), 0 < float
        Standard deviation of the correlation of each connection type.
    """

    def __init__(self):
        # defaults of fcdiff/model.py:31-38
        self.rng = np.random.RandomState(0)
        self.pi = 0.05
        self.eta = 0.3
        self.gamma = np.array([0.1, 0.8, 0.1])
        self.epsilon = 0.03
        self.mu = np.array([-0.15, 0, 0.3])
        self.sigma = np.array([0.025, 0.035, 0.05])

    def __str__(self):
        # fcdiff/model.py:40-50
        return textwrap.dedent('''\
            fcdiff.models.UnsharedRegionModel
                rng = %s
                pi = %g
                eta = %g
                gamma = %s
                epsilon = %g
                mu = %s
                sigma = %s''' % (self.rng, self.pi, self.eta, self.gamma,
                                 self.epsilon, self.mu, self.sigma))

    # ------------------------------------------------------------------ keys
    def _next_key(self):
        """64-bit Philox key + 56-bit stream offset drawn from ``self.rng``."""
        w = self.rng.randint(0, 2 ** 32, size=4, dtype=np.uint64)
        seed = int(w[0]) | (int(w[1]) << 32)
        offset = int(w[2]) | ((int(w[3]) & 0xFFFFFF) << 32)
        return seed, offset

    # ------------------------------------------------------------------ device stages
    def _dev_R(self, key, N, U):
        lib = _lib.load()
        r = _dev.empty((N, U), torch.uint8)
        _lib.check(lib.fcd_sample_R(key[0], key[1], N, U, float(self.pi), _dev.ptr(r), _dev.stream()),
                   "fcd_sample_R")
        return r

    def _dev_T(self, key, r, c0=0, C=None):
        lib = _lib.load()
        (N, U) = r.shape
        if C is None:
            C = N_to_C(N)
        t = _dev.empty((C, U), torch.uint8)
        _lib.check(lib.fcd_sample_T(key[0], key[1], _dev.ptr(r), N, U, float(self.eta), c0, C,
                                    _dev.ptr(t), _dev.stream()), "fcd_sample_T")
        return t

    def _dev_F(self, key, C, c0=0):
        lib = _lib.load()
        f = _dev.empty((C, 3), torch.uint8)
        _lib.check(lib.fcd_sample_F(key[0], key[1], c0, C, _lib.d3(self.gamma), _dev.ptr(f), _dev.stream()),
                   "fcd_sample_F")
        return f

    def _dev_F_tilde(self, key, f, t, c0=0):
        lib = _lib.load()
        (C, U) = t.shape
        ft = _dev.empty((C, U, 3), torch.uint8)
        _lib.check(lib.fcd_sample_F_tilde(key[0], key[1], _dev.ptr(f), _dev.ptr(t), c0, C, U,
                                          float(self.epsilon), _dev.ptr(ft), _dev.stream()),
                   "fcd_sample_F_tilde")
        return ft

    def _dev_B(self, key, f, H, c0=0):
        lib = _lib.load()
        C = f.shape[0]
        b = _dev.empty((C, H))
        _lib.check(lib.fcd_sample_B(key[0], key[1], _dev.ptr(f), c0, C, H, _lib.d3(self.mu),
                                    _lib.d3(self.sigma), _dev.ptr(b), _dev.stream()), "fcd_sample_B")
        return b

    def _dev_B_tilde(self, key, ft, c0=0):
        lib = _lib.load()
        (C, U) = ft.shape[0:2]
        bt = _dev.empty((C, U))
        _lib.check(lib.fcd_sample_B_tilde(key[0], key[1], _dev.ptr(ft), c0, C, U, _lib.d3(self.mu),
                                          _lib.d3(self.sigma), _dev.ptr(bt), _dev.stream()),
                   "fcd_sample_B_tilde")
        return bt

    def sample_device(self, N, H, U, c0=0, C=None):
        """All variables as device tensors (uint8 / float64); ``[c0, c0+C)`` is
        the edge shard this rank generates (``r`` is generated in full on every
        rank from the same key)."""
        key = self._next_key()
        if C is None:
            C = N_to_C(N) - c0
        r = self._dev_R(key, N, U)
        t = self._dev_T(key, r, c0, C)
        f = self._dev_F(key, C, c0)
        ft = self._dev_F_tilde(key, f, t, c0)
        b = self._dev_B(key, f, H, c0)
        bt = self._dev_B_tilde(key, ft, c0)
        return (r, t, f, ft, b, bt)

    # ------------------------------------------------------------------ public API
    def sample(self, N, H, U):
        """
        Samples all random variables from the model (fcdiff/model.py:52-90).

        Returns
        -------
        r : (N, U) bool, t : (C, U) bool, f : (C, 3) bool, f_tilde : (C, U, 3) bool,
        b : (C, H) float64 in [-1, 1], b_tilde : (C, U) float64 in [-1, 1]
        """
        (r, t, f, ft, b, bt) = self.sample_device(N, H, U)
        return (_dev.download(r) > 0, _dev.download(t) > 0, _dev.download(f) > 0,
                _dev.download(ft) > 0, _dev.download(b), _dev.download(bt))

    def sample_R(self, N, U):
        """Anomalous regions of unhealthy patients, (N, U) bool (fcdiff/model.py:92-109)."""
        return _dev.download(self._dev_R(self._next_key(), N, U)) > 0

    def sample_T(self, r):
        """Anomalous connections given anomalous regions, (C, U) bool
        (fcdiff/model.py:111-143)."""
        r_dev = _dev.upload(np.asarray(r) != 0, np.uint8)
        return _dev.download(self._dev_T(self._next_key(), r_dev)) > 0

    def sample_F(self, N):
        """Connection template of healthy subjects, (C, 3) bool (fcdiff/model.py:145-160)."""
        return _dev.download(self._dev_F(self._next_key(), N_to_C(N))) > 0

    def sample_F_tilde(self, f, t):
        """Patients' connections given a template and anomalous connections,
        (C, U, 3) bool (fcdiff/model.py:162-189)."""
        f_dev = _dev.upload(np.asarray(f) != 0, np.uint8)
        t_dev = _dev.upload(np.asarray(t) != 0, np.uint8)
        return _dev.download(self._dev_F_tilde(self._next_key(), f_dev, t_dev)) > 0

    def sample_B(self, f, H):
        """Correlations of healthy subjects given a template, (C, H) float64
        clipped to [-1, 1] (fcdiff/model.py:191-213)."""
        f_dev = _dev.upload(np.asarray(f) != 0, np.uint8)
        return _dev.download(self._dev_B(self._next_key(), f_dev, H))

    def sample_B_tilde(self, f_tilde):
        """Correlations of unhealthy patients given their connections, (C, U)
        float64 clipped to [-1, 1] (fcdiff/model.py:215-236)."""
        ft_dev = _dev.upload(np.asarray(f_tilde) != 0, np.uint8)
        return _dev.download(self._dev_B_tilde(self._next_key(), ft_dev))
