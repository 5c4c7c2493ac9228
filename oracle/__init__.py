"""CPU oracle for the MPI-Flow hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this package.  The product (mpi-flow_b200/) never does: it fails loudly without its CUDA library.
"""
