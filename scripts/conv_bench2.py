"""Bottleneck hunt for conv3x3_tc: same launch with parts disabled (MZ_TC_DEBUG_SKIP bitmask)."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = '''
import os, sys, numpy
sys.path.insert(0, %r)
os.environ["MZ_DEBUG_CONV_REPS"] = "50"
from muzero_general_b200.engine import debug_conv3x3
rs = numpy.random.RandomState(0)
for n in (1024, 4096):
    x = rs.standard_normal((n, 64, 6, 7)).astype(numpy.float32)
    w = (rs.standard_normal((64, 64, 3, 3)) / 24).astype(numpy.float32)
    debug_conv3x3(x, w, None, x, True, tensor_cores=True)
''' % root
for skip in (0, 1, 2, 4, 8, 3, 6, 7, 15):
    env = dict(os.environ, MZ_TC_DEBUG_SKIP=str(skip))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    for line in out.stderr.splitlines():
        if "mz_debug" in line:
            print(f"skip={skip:2d} {line}")
