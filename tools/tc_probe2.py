import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tc_probe_lib import *

W = 16384
lo = (np.arange(W) % 2048).astype(np.float32)
hi = (np.arange(W) // 2048).astype(np.float32)
for layout in [2, 1, 4, 6]:
    for lbo, sbo in [(8192, 1024), (1024, 8192)]:
        d1 = run(kmajor_identity(128), lo, desc(2048, 128), desc(lbo, sbo, layout), idesc(128, 64, False, True), 64)
        d2 = run(kmajor_identity(128), hi, desc(2048, 128), desc(lbo, sbo, layout), idesc(128, 64, False, True), 64)
        idx = (d2[:8] * 2048 + d1[:8]).astype(np.int64)
        print("B MN-major layout_type %d lbo %d sbo %d: byte offsets of B[n][k]" % (layout, lbo, sbo))
        for n in [0, 1, 2, 3, 4, 5, 7, 8, 12, 16, 28, 31, 32, 33, 63]:
            print("   n=%2d " % n, " ".join("%6d" % (4 * idx[k][n]) for k in range(8)))
# K-major SW128 reference view of the same kind of tile (rows = N, K = 8 columns starting at column 0 and at column 8)
for layout in [2]:
    for start_words in [0, 8]:
        img = np.roll(lo, -start_words)   # emulate advancing the start address by 32 B: not exact, so use descriptor start instead
    d1 = run(kmajor_identity(128), lo, desc(2048, 128), desc(16, 1024, 2), idesc(128, 64, False, False), 64)
    idx = d1[:8].astype(np.int64)
    print("B K-major SW128 (sbo 1024): byte offsets of B[n][k]")
    for n in [0, 1, 2, 7, 8, 9, 63]:
        print("   n=%2d " % n, " ".join("%6d" % (4 * idx[k][n]) for k in range(8)))
