"""CPU oracle for the MS-CNN forward path -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (mscnn_b200/) never does.

  oracle.ref      -- ctypes view of oracle/_ref/libmscnn_ref.so: the reference's own layer code
                     compiled verbatim (oracle/build_ref.py).  kind = "reference".
  oracle.port     -- numpy / C restatement of the same algorithms, each function citing the
                     reference file:line it follows; pinned against oracle.ref and the
                     reference's golden vectors (tests/test_oracle.py).  kind = "port".
"""
