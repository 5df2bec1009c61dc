"""
fcdiff_b200 -- B200-native (sm_100a CUDA) drop-in for the hot path of
andy-sweet/fcdiff: variational-EM inference in the individual-anomalous-region
model of population functional-connectivity differences.

The package surface mirrors the reference's ``fcdiff/__init__.py:1-5``::

    fcdiff_b200.UnsharedRegionModel, fcdiff_b200.fit, fcdiff_b200.N_to_C,
    fcdiff_b200.nm_to_c, fcdiff_b200.c_to_nm

plus ``fcdiff_b200.corr`` (time series -> Fisher-z correlations, new),
``fcdiff_b200.dist`` (edge sharding over the GPUs of one box) and
``fcdiff_b200.sweep`` (label-permutation / random-restart replicas).  All arithmetic
of the path runs in ``libfcdiff_b200.so`` (hand-written CUDA behind a C-ABI,
``include/fcdiff_b200.h``); there is no CPU fallback.
"""
from .model import UnsharedRegionModel

from . import fit
from . import model
from . import util

from .util import N_to_C, nm_to_c, c_to_nm

__all__ = ["UnsharedRegionModel", "fit", "model", "util", "N_to_C", "nm_to_c", "c_to_nm"]
