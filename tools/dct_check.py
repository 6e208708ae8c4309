"""GPU bring-up check of the tcgen05 DCT against the oracle (prints per-tile error maps)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rtvqa_b200
from rtvqa_b200 import _native as N
from oracle import np_oracle as NO

ctx = N.Context(0)
sizes = [(64, 64), (128, 128), (37, 100), (128, 96), (200, 264), (270, 480), (1080, 1920)]
if len(sys.argv) > 1:
    sizes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
ok = True
for h, w in sizes:
    g = np.random.default_rng(h * 7 + w).integers(0, 256, (h, w), dtype=np.uint8)
    want = NO.dct2(g)
    t = time.time()
    c = ctx.debug_dct(g, 0).astype(np.float64)
    dt = time.time() - t
    err = np.abs(c - want)
    scale = np.abs(want).max()
    e_rel = abs(np.sum(c * c) - np.sum(want * want)) / np.sum(want * want)
    l1 = abs(np.abs(c).sum() - np.abs(want).sum()) / np.abs(want).sum()
    good = err.max() <= 2e-6 * scale + 2e-2 and e_rel < 1e-5
    ok &= good
    print(f"{h}x{w}: max|err| {err.max():.4g} (DC {scale:.4g})  energy rel {e_rel:.3g}  L1 rel {l1:.3g}  {dt*1e3:.1f} ms  {'OK' if good else 'BAD'}")
    if not good:
        th, tw = -(-h // 128), -(-w // 128)
        m = np.zeros((th, tw))
        for i in range(th):
            for j in range(tw):
                m[i, j] = err[i * 128:(i + 1) * 128, j * 128:(j + 1) * 128].max()
        np.set_printoptions(linewidth=200, precision=3, suppress=True)
        print("per-tile max err:\n", m)
        print("got[:4,:6]\n", c[:4, :6], "\nwant[:4,:6]\n", want[:4, :6])
print("ALL OK" if ok else "FAILED")
