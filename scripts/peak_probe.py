import ctypes, sys
sys.path.insert(0, "/root/repo")
from paramol_b200 import _lib
L = _lib.lib()
for iters in (1000, 4000, 16000, 64000):
    fl, ms = ctypes.c_double(), ctypes.c_double()
    _lib.check(L.pm_measure_fma_peak(0, 1, iters, ctypes.byref(fl), ctypes.byref(ms)))
    print("iters", iters, "TFLOP/s %.2f" % (fl.value / 1e12), "ms %.3f" % ms.value)
