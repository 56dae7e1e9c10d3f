import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roitr_b200.synthetic import synthetic_pair
from tests import parity
from tests.helpers import CONFIG_3D, weights
for n, i in [(1024, 0), (4096, 1)] + ([(20000, 0)] if "--full" in sys.argv else []):
    t = time.time()
    try:
        rows, out, ref = parity.run(synthetic_pair(i, n), CONFIG_3D, weights(1))
        print("==== N=%d (%.1fs)" % (n, time.time() - t)); print(parity.format_rows(rows)); print("FAILURES:", parity.failures(rows))
    except Exception as e:
        import traceback; traceback.print_exc()
