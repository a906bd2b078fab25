"""How much tensor work could be skipped if the cells were sorted along a space-filling curve and
the MMA kernel skipped every (cell tile x 128-cell stage) product whose weight digit plane is all
zero?    Exact count on the host: for a projection of the synthetic workload
the fixed-point weights W_ij = round(w'_ij * (2^32 - 1)), w'_ij = exp(-(d_ij^2 - m_i) / sigma^2),
are formed for every (tile of 128 cells i) x (all j), and a digit plane a (0 = most significant
byte) of a 128 x 128 block can be non-zero only if some w' in the block is >= 2^-8(a+1).
Products per stage: (weight digit a, value digit b) with a + b <= 3 -> [2, 2, 2, 1] per weight
digit with two value digits, [3, 3, 2, 1] with three.  CPU only; results in
profiles/r02_tile_skipping_estimate.md."""
import numpy as np, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fastproject_b200 import synth
from scipy.spatial import cKDTree
def hilbert_d(x, y, order=16):
    # vectorised xy -> hilbert index
    x = x.astype(np.int64).copy(); y = y.astype(np.int64).copy()
    d = np.zeros_like(x)
    s = 1 << (order-1)
    while s > 0:
        rx = ((x & s) > 0).astype(np.int64); ry = ((y & s) > 0).astype(np.int64)
        d += s * s * ((3 * rx) ^ ry)
        # rotate
        m = (ry == 0)
        flip = m & (rx == 1)
        x[flip] = s - 1 - x[flip]; y[flip] = s - 1 - y[flip]
        xs = x.copy(); x[m] = y[m]; y[m] = xs[m]
        s >>= 1
    return d
def morton(ix, iy):
    def part(v):
        v = v.astype(np.uint64)
        v = (v | (v << 16)) & 0x0000FFFF0000FFFF
        v = (v | (v << 8)) & 0x00FF00FF00FF00FF
        v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0F
        v = (v | (v << 2)) & 0x3333333333333333
        v = (v | (v << 1)) & 0x5555555555555555
        return v
    return part(ix) | (part(iy) << 1)
def est(coords, sr, label, curve):
    n = coords.shape[1]
    pts = coords.T
    lo, hi = pts.min(0), pts.max(0)
    # clip to 3 sigma box like nn grid? use quantiles
    q = np.clip(((pts - lo) / (hi - lo) * 65535), 0, 65535).astype(np.int64)
    key = hilbert_d(q[:,0], q[:,1]) if curve=="hilbert" else morton(q[:,0], q[:,1])
    order = np.argsort(key) if curve!="none" else np.arange(n)
    p = pts[order]
    tree = cKDTree(p); d, _ = tree.query(p, k=2); m = d[:,1]**2
    T = 128
    nt = (n + T - 1)//T
    cnt = np.zeros(5)
    tot = 0
    for t in range(nt):
        pi = p[t*T:(t+1)*T]; mi = m[t*T:(t+1)*T]
        d2 = ((pi[:,None,:]-p[None,:,:])**2).sum(-1)          # T x n
        lw = -(d2 - mi[:,None])/0.33**2/np.log(2)
        # per j-stage max over tile rows and 128 cols
        pad = (-n) % T
        mx = np.pad(lw.max(0), (0,pad), constant_values=-1e9).reshape(-1,T).max(1)
        for a in range(4):
            cnt[a] += (mx >= -8*(a+1)).sum()
        tot += mx.size
    frac = cnt[:4]/tot
    prods = {2:[2,2,2,1], 3:[3,3,2,1]}[sr]
    avg = sum(f*c for f,c in zip(frac, prods)); full=sum(prods)
    print(label, curve, "n=%d nonzero fracs"%n, np.round(frac,3), "MMAs %.2f/%d = %.3f"%(avg, full, avg/full), flush=True)
proj,_ = synth.make_projections(20000, 6, seed=1335607829)
for k in ["Proj00","Proj03"]:
    for curve in ("none","morton","hilbert"):
        est(proj[k], 2, k, curve)
