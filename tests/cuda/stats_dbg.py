import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from fastproject_b200 import _engine as E
rng = np.random.RandomState(1)
for n in [1, 2, 7, 8, 9, 127, 128, 129, 255, 500, 1000, 4099, 20011]:
    rmed = rng.normal(500, 4, size=n)
    order = np.arange(n, dtype=np.int32); ptr = np.array([0, n], dtype=np.int32); grp = np.zeros(3, dtype=np.int32)
    med = rng.normal(500, 4, size=3)
    t = lambda a, dt: torch.from_numpy(np.asarray(a)).to("cuda", dt)
    try:
        p, mu, sd = E.background_pvalues(t(med, torch.float64), t(grp, torch.int32), t(rmed, torch.float64), t(order, torch.int32), t(ptr, torch.int32))
        torch.cuda.synchronize()
        print(n, "ok", float(mu[0]) == rmed.mean(), float(sd[0]) == rmed.std())
    except Exception as e:
        print(n, "FAIL", str(e)[:80]); break
