"""Bring-up tool: how does tcgen05.mma (kind::tf32) walk a NO-SWIZZLE MN-major operand?  Runs on the GPU box.
One operand is a K-major 8x8 identity (layout known to work: tests/test_gpu_tc.py::test_gemm_tf32x3), the other is an
image whose words hold their own index, so the accumulator displays which shared-memory word was read for every (row, k)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from seggroup_b200 import _lib, ops


def desc(lbo, sbo, layout=0):
    return ((lbo >> 4) & 0x3fff) << 16 | ((sbo >> 4) & 0x3fff) << 32 | 1 << 46 | layout << 61


def idesc(M, N, a_mn, b_mn):
    return (1 << 4) | (2 << 7) | (2 << 10) | (int(a_mn) << 15) | (int(b_mn) << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def kmajor_identity(R):
    img = np.zeros(max(R * 8, 256), np.float32)
    for m in range(8):
        c = m
        off = (c // 4) * (R * 16) + (m // 8) * 128 + (m % 8) * 16 + (c % 4) * 4
        img[off // 4] = 1.0
    return img


def run(imgA, imgB, dA, dB, ide, N):
    D = torch.empty(128, N, device="cuda")
    a = torch.as_tensor(imgA).cuda(); b = torch.as_tensor(imgB).cuda()
    _lib.call("sgb_tc_probe", a, a.numel(), b, b.numel(), dA, dB, ide, N, D, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return D.cpu().numpy()


W = 16384
lo = (np.arange(W) % 2048).astype(np.float32)
hi = (np.arange(W) // 2048).astype(np.float32)
print("=== sanity: B operand K-major (N = 64) through the same probe")
d1 = run(kmajor_identity(128), lo, desc(2048, 128), desc(1024, 128), idesc(128, 64, False, False), 64)
idx = d1[:8].astype(np.int64)
for n in [0, 1, 7, 8, 9, 63]:
    print("   n=%2d " % n, " ".join("%6d" % (4 * idx[k][n]) for k in range(8)))
print("=== B operand MN-major (N = 64), A = K-major identity: D[k][n] = word read for B[n][k]")
for lbo, sbo in [(128, 2048), (2048, 128), (128, 1024), (1024, 128), (256, 2048), (128, 128)]:
    try:
        d1 = run(kmajor_identity(128), lo, desc(2048, 128), desc(lbo, sbo), idesc(128, 64, False, True), 64)
        d2 = run(kmajor_identity(128), hi, desc(2048, 128), desc(lbo, sbo), idesc(128, 64, False, True), 64)
        idx = (d2[:8] * 2048 + d1[:8]).astype(np.int64)       # [k][n]
        print("lbo %5d sbo %5d : byte offsets of B[n][k], rows n = 0..9, 16, 32, 63 ; cols k = 0..7" % (lbo, sbo))
        for n in list(range(10)) + [16, 32, 63]:
            print("   n=%2d " % n, " ".join("%6d" % (4 * idx[k][n]) for k in range(8)))
    except Exception as e:
        print("lbo %d sbo %d failed: %s" % (lbo, sbo, e))
print("=== A operand MN-major (M = 128), B = K-major identity (N = 16): D[m][k] = word read for A[m][k]")
for lbo, sbo in [(128, 2048), (2048, 128)]:
    d1 = run(lo, kmajor_identity(16), desc(lbo, sbo), desc(256, 128), idesc(128, 16, True, False), 16)
    d2 = run(hi, kmajor_identity(16), desc(lbo, sbo), desc(256, 128), idesc(128, 16, True, False), 16)
    idx = (d2[:, :8] * 2048 + d1[:, :8]).astype(np.int64)     # [m][k]
    print("lbo %5d sbo %5d : byte offsets of A[m][k], rows m = 0..9, 16, 64, 127" % (lbo, sbo))
    for m in list(range(10)) + [16, 64, 127]:
        print("   m=%3d " % m, " ".join("%6d" % (4 * idx[m][k]) for k in range(8)))
