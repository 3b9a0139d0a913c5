"""Golden pickle of the rollout writer from the REFERENCE'S OWN save_rollout_results (motion/crowd_ppo/utils.py:10-50) -
build container only. The function is pure Python / torch; it is run on a deterministic synthetic rollout
(tests/test_host_logic.py::synthetic_rollout) and its pickle is stored verbatim.

Run:  python tests/golden/gen_rollout_golden.py   ->  tests/golden/rollout_golden.pkl
"""
import glob
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, os.path.join(HERE, ".."))
REF = os.environ.get("EGOGEN_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "motion"))
from crowd_ppo import utils as ref_utils                     # noqa: E402
from test_host_logic import synthetic_rollout                # noqa: E402

scene, outmps = synthetic_rollout()
with tempfile.TemporaryDirectory() as tmp:
    ref_utils.save_rollout_results(scene, outmps, os.path.join(tmp, "out"), man_id="golden7")
    files = glob.glob(os.path.join(tmp, "out", "*"))
    assert [os.path.basename(f) for f in files] == ["motion_golden7.pkl"], files
    shutil.copy(files[0], os.path.join(HERE, "rollout_golden.pkl"))
print("wrote rollout_golden.pkl", os.path.getsize(os.path.join(HERE, "rollout_golden.pkl")), "bytes")
