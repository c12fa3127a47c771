"""Import-path shim: `import cpp_wrappers.cpp_subsampling.grid_subsampling as cpp_subsampling` (kpconv/datasets/
common.py:29, Scannet.py:46) resolves here when `<repo>/seggroup_b200/compat` is on sys.path."""
from seggroup_b200.kpconv_ops import grid_subsampling as _m

compute = _m.compute
