"""Test infrastructure: CPU restatement of the reference path (oracle.py / pn2_oracle.c), its module-level
composition (modules_ref.py) and a loader for the reference's own kernels compiled verbatim (ref_cuda.py).
Never imported by the product package."""
