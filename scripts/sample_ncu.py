"""Launch pinn_sample_points a few times on 16 Mi x 3 rows (for an ncu capture of the sampler; scripts/gpu_round.sh)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pinnacle_b200 import capi  # noqa: E402
from pinnacle_b200.resample import Region  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "regions.npz"))
m = json.loads(bytes(z["meta"]).decode())["Heat2D_ComplexGeometry"]
x = torch.empty((1 << 24, 3), device="cuda:0")
for reg in (Region(m["lo"], m["hi"]), Region(m["lo"], m["hi"], [tuple(h) for h in m["holes"]])):
    for i in range(3):
        capi.sample_points(x, reg.to_c(), 1, i, 0)
torch.cuda.synchronize()
print("ok", float(x[:, 2].mean()))
