"""acf_b200 -- B200 (sm_100a) engine for the ACF chnsPyramid + acfDetect hot path of elucideye/acf.

The CUDA library (acf_b200/libacf_b200.so, C ABI in include/acf_b200.h) is the product; this
package is the Python host mirror of the reference's detector interface used by tests and bench.py.
"""
from ._capi import AcfError, lib  # noqa: F401
from .detector import Detector, Model, Pyramid, get_scales  # noqa: F401
from . import synth  # noqa: F401
