import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from src.algorithms.velocity_solver_improved import ImprovedVelocitySolver
g = np.load("/root/repo/tests/golden/interframe_de.npz")
for name in ("slow", "dense"):
    hit = g[f"{name}_match_cur"]; az = g[f"{name}_cur_az"][hit]; y = g[f"{name}_y"]
    s = ImprovedVelocitySolver(fc=77e9, lambda_c=3e8/77e9, num_antennas=4)
    k = 4*np.pi*0.1/s.lambda_c
    s._global_search(np.cos(az), np.sin(az), y, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); v, f, npts = s._global_search(np.cos(az), np.sin(az), y, k); e1.record(); torch.cuda.synchronize()
    print(name, "N", len(az), "points", npts, "ms (kernel + topk + host polish)", e0.elapsed_time(e1), "v", v, "cost", f)
