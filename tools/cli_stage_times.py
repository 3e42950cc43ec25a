import os, sys, subprocess, tempfile, time
sys.path.insert(0, os.getcwd())
import numpy as np
from arap_flow_b200 import driver, flowio, synth
sys.path.insert(0, "tools")
import cli_throughput as T
d = tempfile.mkdtemp()
items = T.write_items(d, "C3", 8)
lst = os.path.join(d, "l.txt"); driver.write_list_file(lst, items)
for b in (8, 3, 4):
    env = dict(os.environ, ARAP_PLAN=driver.PLAN, ARAP_TIMING="1", ARAP_BATCH=str(b))
    t0 = time.time(); r = subprocess.run([driver.ARAP_BIN, lst], env=env, capture_output=True, text=True); dt = time.time() - t0
    print("batch", b, "wall %.2f" % dt, r.stderr.strip().splitlines()[-1])
