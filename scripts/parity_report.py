"""Prints the stage-by-stage parity table (tests/parity.py) of the CUDA forward against the oracle for the BASELINE.json
configurations. Usage: python scripts/parity_report.py [small] [20k] [30k] [4d8k] [unscaled]"""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from roitr_b200.synthetic import synthetic_pair  # noqa: E402
from tests import parity  # noqa: E402
from tests.helpers import CONFIG_3D, CONFIG_4D, baseline_pair, weights  # noqa: E402

CASES = {
    "small": [("3DMatch N=1024", lambda: synthetic_pair(0, 1024), CONFIG_3D, 1, {}),
              ("3DMatch N=4096", lambda: synthetic_pair(1, 4096), CONFIG_3D, 1, {}),
              ("4DMatch N=2048", lambda: synthetic_pair(3, 2048, deform=True), CONFIG_4D, 2, {})],
    "20k": [("3DMatch N=20000", lambda: synthetic_pair(0, 20000), CONFIG_3D, 1, {})],
    "30k": [("3DMatch 30000/28000 (BASELINE config 3)", lambda: baseline_pair("3dmatch_30k"), CONFIG_3D, 1, {})],
    "4d8k": [("4DMatch 8000/7000 factor 2 (BASELINE config 5)", lambda: baseline_pair("4dmatch_8k"), CONFIG_4D, 2, {})],
    "unscaled": [("3DMatch N=4096, fine_proj NOT scaled x8", lambda: synthetic_pair(1, 4096), CONFIG_3D, 1, dict(fine_scale=1.0))],
}

if __name__ == "__main__":
    torch.set_num_threads(min(os.cpu_count() or 1, 32))
    which = [a for a in sys.argv[1:] if a in CASES] or ["small"]
    for w in which:
        for name, mk, cfg, factor, kw in CASES[w]:
            t = time.time()
            try:
                rows, out, ref = parity.run(mk(), cfg, weights(factor, **kw))
                print("==== %s (%.1fs)  P=%d correspondences=%d" % (name, time.time() - t, out["matching_scores"].shape[0],
                                                                    out["corr_scores"].shape[0]))
                print(parity.format_rows(rows))
                print("FAILURES:", parity.failures(rows))
            except Exception:
                traceback.print_exc()
            sys.stdout.flush()
