"""
Drop-in alias: ``import fcdiff`` resolves to the B200-native implementation
(``fcdiff_b200``) with the reference's package surface (fcdiff/__init__.py:1-5),
so the reference's own scripts and tests run unchanged.
"""
import sys as _sys

import fcdiff_b200 as _impl
from fcdiff_b200 import UnsharedRegionModel, fit, model, util, N_to_C, nm_to_c, c_to_nm  # noqa: F401

_sys.modules[__name__ + ".fit"] = _impl.fit
_sys.modules[__name__ + ".model"] = _impl.model
_sys.modules[__name__ + ".util"] = _impl.util
