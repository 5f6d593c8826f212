"""GPU box: properties test at BASELINE's full size + peak device memory."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from util import *
from omp_amg_b200 import matrices as M
import test_gpu_parity as T
L = api.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
t = time.time()
T.test_size_independent_properties(L, "poisson7", n)
print("properties at poisson7 %d^3: OK (%.0f s incl. host SpGEMM checks)" % (n, time.time() - t))
print("peak device bytes: %.2f GB" % (L.amgb_peak_device_bytes() / 1e9))
