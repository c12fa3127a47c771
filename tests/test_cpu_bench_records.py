"""bench.py bookkeeping that needs no GPU: the roofline block takes `traffic` / `ncu` from the newest committed ncu summary under
profiles/ (nothing typed in), and the kernel it names must be one the shipped library still contains."""
import importlib.util
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("sgb_bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(m)
    finally:
        sys.argv = argv
    return m


def test_roofline_record_is_the_current_forward_kernel():
    b = _bench()
    rec = b.ncu_record("ec2_tc1_kernel<1, 1>") or b.ncu_record("ec2_tc1_kernel<true, true>")
    assert rec is not None, "no committed ncu summary names the dominant kernel: roofline.traffic would be null"
    assert rec["source"] >= "profiles/r03" and rec["dram_bytes"] > 0 and rec["points"] == 150000      # this round's capture or a later one
    # algorithmic bytes of the kernel (436 B per point) against the measured DRAM traffic: no wasted re-reads
    assert rec["dram_bytes"] < 1.1 * 436 * rec["points"]
    for key in ("ec2_bwd_tc_kernel", "segment_pool_staged_kernel", "knn_sweep_kernel", "centralize_kernel", "export_labels_kernel"):
        assert b.ncu_record(key) is not None, key


def test_named_kernels_exist_in_the_library():
    lib = os.path.join(ROOT, "seggroup_b200", "lib", "libseggroup_b200.so")
    if not os.path.isfile(lib):
        import pytest
        pytest.skip("library not built")
    syms = subprocess.run(["nm", "-C", lib], capture_output=True, text=True).stdout
    for k in ("ec2_tc1_kernel", "ec2_bwd_tc_kernel", "segment_pool_staged_kernel", "knn_sweep_kernel"):
        assert k in syms, k
    assert "ec2_tc_kernel<" not in syms          # the round-2 forward kernel is gone, not shipped next to its successor
