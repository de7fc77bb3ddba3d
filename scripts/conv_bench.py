"""Warm-L2 timing of the bare conv kernels through mz_debug_conv3x3 (MZ_DEBUG_CONV_REPS)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy
os.environ["MZ_DEBUG_CONV_REPS"] = "50"
from muzero_general_b200.engine import debug_conv3x3
rs = numpy.random.RandomState(0)
for n in (1024, 4096):
    x = rs.standard_normal((n, 64, 6, 7)).astype(numpy.float32)
    w = (rs.standard_normal((64, 64, 3, 3)) / 24).astype(numpy.float32)
    b = rs.standard_normal(64).astype(numpy.float32)
    for tc in (True, False):
        debug_conv3x3(x, w, b, None, True, tensor_cores=tc)
        debug_conv3x3(x, w, b, x, True, tensor_cores=tc)
