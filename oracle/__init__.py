"""oracle/ — TEST INFRASTRUCTURE.  CPU restatement of the reference algorithm for the hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import anything from here.  The product (cumind_b200/) never does and fails loudly without its
CUDA library.
"""
