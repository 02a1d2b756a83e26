"""CPU oracle for the RADE hot path — TEST INFRASTRUCTURE, NOT THE PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
this package.  The product (radae_b200/, libradae_b200.so) never does, and fails loudly without its CUDA
library instead of falling back here.
"""
