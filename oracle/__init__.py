"""CPU oracle (TEST INFRASTRUCTURE ONLY — see oracle/tf_oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  Nothing under texturefusion_b200/ does.
"""
from .oracle import (OracleMap, build_oracle, build_ref, have_patch_ref, have_ref, oracle_lib_path, patch_texcoords,  # noqa: F401
                     ref_lib_path, set_dot3_order)
