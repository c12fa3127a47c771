"""Where do the rows of an M = 64 accumulator land in TMEM (cta_group::1, kind::tf32)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tc_probe_lib import *

def kmajor(R, vals):   # vals [R][8] -> image
    img = np.zeros(max(R * 8, 256), np.float32)
    for r in range(R):
        for c in range(8):
            off = (c // 4) * (R * 16) + (r // 8) * 128 + (r % 8) * 16 + (c % 4) * 4
            img[off // 4] = vals[r][c]
    return img

for M in (64, 128):
    A = np.zeros((M, 8), np.float32); A[:, 0] = np.arange(M) + 1
    B = np.zeros((16, 8), np.float32); B[:, 0] = 1.0; B[3, 0] = 2.0
    # poison TMEM first with a full M=128 product of -1 rows, then run the M-row product
    P = np.zeros((128, 8), np.float32); P[:, 0] = -1
    run(kmajor(128, P), kmajor(16, B), desc(128 * 16, 128), desc(16 * 16, 128), idesc(128, 16, False, False), 16)
    D = run(kmajor(M, A), kmajor(16, B), desc(M * 16, 128), desc(16 * 16, 128), idesc(M, 16, False, False), 16)
    print("M = %d: value in column 0 / column 3 of every TMEM lane (row id + 1, x2 in column 3)" % M)
    for l0 in range(0, 128, 16):
        print("  lanes %3d..%3d:" % (l0, l0 + 15), " ".join("%4d" % int(D[l][0]) for l in range(l0, l0 + 16)), "|", " ".join("%4d" % int(D[l][3]) for l in range(l0, l0 + 4)))
