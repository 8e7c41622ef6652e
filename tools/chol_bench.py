"""Reduced-solve timing on the GPU box: the in-cluster DMMA Cholesky (adb_dense_solve, CUDA events around the kernel) against
cuSOLVER potrf + potrs as torch.linalg calls them (torch.linalg.cholesky_ex + torch.cholesky_solve, FP64, CUDA events, warm),
at the orders of BASELINE configs 4 and 5.  Measurement only: torch / cuSOLVER are never on the product path."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from airdos_b200 import ba  # noqa: E402


def spd(n, seed):
    rng = np.random.default_rng(seed)
    m = rng.normal(size=(n, n))
    return m @ m.T + n * np.eye(n), rng.normal(size=n)


def cusolver_ms(a, b, reps=50):
    import torch
    A = torch.tensor(a, device="cuda"); B = torch.tensor(b, device="cuda")[:, None]
    for _ in range(5):
        L, info = torch.linalg.cholesky_ex(A); x = torch.cholesky_solve(B, L)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        L, info = torch.linalg.cholesky_ex(A); x = torch.cholesky_solve(B, L)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, x[:, 0].cpu().numpy()


def main():
    out = []
    for n in (96, 294, 474, 1226, 2048):
        a, b = spd(n, n)
        ref = np.linalg.solve(a, b)
        row = {"n": n}
        for cl in (4, 8, 16):
            try:
                x, info, ms = ba.dense_solve(a, b, cluster=cl, reps=30)
                row[f"adb_cluster{cl}_ms"] = round(ms, 4)
                row[f"adb_cluster{cl}_relerr"] = float(np.abs(x - ref).max() / np.abs(ref).max())
            except Exception as e:   # noqa: BLE001
                row[f"adb_cluster{cl}_error"] = repr(e)
        try:
            ms, x = cusolver_ms(a, b)
            row["cusolver_potrf_potrs_ms"] = round(ms, 4)
            row["cusolver_relerr"] = float(np.abs(x - ref).max() / np.abs(ref).max())
        except Exception as e:   # noqa: BLE001
            row["cusolver_error"] = repr(e)
        row["gflops_adb_best"] = round(n ** 3 / 3 / (min(v for k, v in row.items() if k.endswith("_ms") and k.startswith("adb")) * 1e-3) / 1e9, 1)
        out.append(row)
        print(json.dumps(row), flush=True)
    return out


if __name__ == "__main__":
    main()
