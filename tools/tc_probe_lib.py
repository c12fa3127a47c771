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


