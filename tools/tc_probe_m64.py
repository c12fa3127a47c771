"""Bring-up tool: which TMEM lanes hold the rows of a cta_group::1 tcgen05.mma accumulator with M = 64?  Runs on the GPU box.
A = K-major [64 x 8] with A[r][0] = r + 1, B = K-major 16 x 8 with B[n][n] = 1 (n < 8)  ->  D[r][0] = r + 1.
The probe kernel returns all 128 TMEM lanes, so the lane of every row can be read off column 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tc_probe_lib import desc, idesc, run


def kmajor(R, vals):
    """vals: dict (r, c) -> value, c < 8; canonical no-swizzle K-major image of an R x 8 tile."""
    img = np.zeros(max(R * 8, 256), np.float32)
    for (r, c), v in vals.items():
        off = (c // 4) * (R * 16) + (r // 8) * 128 + (r % 8) * 16 + (c % 4) * 4
        img[off // 4] = v
    return img


if __name__ == "__main__":
    for M in (64, 128):
        A = kmajor(M, {(r, 0): float(r + 1) for r in range(M)})
        B = kmajor(16, {(n, n): 1.0 for n in range(8)})
        D = run(A, B, desc(M * 16, 128), desc(16 * 16, 128), idesc(M, 16, False, False), 16)
        col = D[:, 0]
        lanes = {}
        for lane in range(128):
            v = col[lane]
            if v == int(v) and 1 <= v <= M:
                lanes.setdefault(int(v) - 1, []).append(lane)
        print("M =", M, "row -> lane:", [(r, lanes.get(r)) for r in (0, 1, 15, 16, 17, 31, 32, 47, 48, 63)])
        rows_found = sorted(lanes)
        print("   rows found:", len(rows_found), "lanes used:", sorted(l for v in lanes.values() for l in v)[:70])
