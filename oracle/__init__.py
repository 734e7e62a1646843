"""oracle/ -- TEST INFRASTRUCTURE ONLY (the checker, never the product).

`oracle.port`  : ctypes view of oracle/_ref/liboracle.so  = oracle.c, the plain-C restatement of the path.
`oracle.ref`   : ctypes view of oracle/_ref/libmisha_ref.so = the reference's own unmodified hot-path objects
                 (compiled in place from /root/reference/src by oracle/Makefile) behind ref_harness.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  misha_b200/ must never import it (tests/test_no_oracle_in_product.py enforces that).
"""
from . import port, ref  # noqa: F401
