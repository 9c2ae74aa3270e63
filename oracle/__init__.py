"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement (oracle/spmv_oracle.c -> liboracle.so) of the reference's SpMV hot path and,
when built, the unmodified reference itself (oracle/_ref/libspmv_ref.so).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product package (spmv_openmp_cuda_b200) never does.
"""
from .oracle import *  # noqa: F401,F403
