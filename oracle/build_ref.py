"""TEST INFRASTRUCTURE — compiles the UNMODIFIED reference C++ cores of the KPConv operator set into
oracle/_ref/libkpconv_ref.so (git-ignored, travels to the GPU box), from the sources where they lie under
/root/reference (never copied).  Flags follow the reference's own build (tf_custom_ops/compile_op.sh:8-13:
-std=c++11 -O2, no -march=native / -ffast-math, so x86 fp32 arithmetic stays non-fused).

    python -m oracle.build_ref
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("KPCONV_REFERENCE", "/root/reference/kpconv")
OUT_DIR = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT_DIR, "libkpconv_ref.so")
SOURCES = ["tf_custom_ops/tf_neighbors/neighbors/neighbors.cpp",
           "tf_custom_ops/tf_subsampling/grid_subsampling/grid_subsampling.cpp",
           "tf_custom_ops/cpp_utils/cloud/cloud.cpp"]
# the cpp_wrappers variant defines `grid_subsampling` with one more parameter (C++ overload) and its own
# SampledData class -> compiled as a separate object with hidden class symbols renamed via a namespace-free
# trick: its SampledData differs, so it is built with -DSampledData=SampledDataW.
WRAPPER_SOURCE = "cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp"


def available() -> bool:
    return os.path.isfile(LIB)


def build(force: bool = False) -> str | None:
    if not os.path.isdir(REF):
        return LIB if available() else None       # GPU box: use the prebuilt file
    if available() and not force:
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    objs = []
    flags = ["-std=c++11", "-O2", "-fPIC", "-w", "-I", REF]
    for i, src in enumerate(SOURCES):
        obj = os.path.join(OUT_DIR, "ref_%d.o" % i)
        subprocess.check_call(["g++"] + flags + ["-c", os.path.join(REF, src), "-o", obj])
        objs.append(obj)
    obj = os.path.join(OUT_DIR, "ref_wrapper.o")
    subprocess.check_call(["g++"] + flags + ["-DSampledData=SampledDataW", "-c", os.path.join(REF, WRAPPER_SOURCE), "-o", obj])
    objs.append(obj)
    obj = os.path.join(OUT_DIR, "ref_driver.o")
    subprocess.check_call(["g++"] + flags + ["-c", os.path.join(HERE, "ref_src", "ref_driver.cpp"), "-o", obj])
    objs.append(obj)
    subprocess.check_call(["g++", "-shared", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
