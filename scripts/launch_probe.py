"""python scripts/launch_probe.py -- floor of a host-buffer call (pf_debug_launch_probe): launch + one word back + poll."""
import ctypes as C, sys
sys.path.insert(0, '.')
import numpy as np
from fetching_b200 import _lib
out = np.zeros(4)
_lib.check(_lib.load().pf_debug_launch_probe(0, out.ctypes.data_as(C.POINTER(C.c_double))), "launch probe")
print("host-buffer call floor (us): 24 B params, 1 CTA %.2f | 3.6 KB params, 1 CTA %.2f | 24 B, 296 CTAs + ticket %.2f | "
      "3.6 KB, 296 CTAs + ticket %.2f" % tuple(out))
