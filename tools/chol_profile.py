import os, sys
sys.path.insert(0, "/root/repo")
os.environ["ADB_CHOL_PROFILE"] = "1"
import numpy as np
from airdos_b200 import ba
for n, cl in ((294, 8), (294, 16), (1226, 16), (1226, 8)):
    rng = np.random.default_rng(n); m = rng.normal(size=(n, n)); a = m @ m.T + n * np.eye(n); b = rng.normal(size=n)
    x, info, ms = ba.dense_solve(a, b, cluster=cl, reps=5)
    print(n, cl, ms, file=sys.stderr)
