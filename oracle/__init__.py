"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's encode-and-contrast algorithm (msclip_oracle.py), the shim that
imports the real reference in the authoring container (ref_shim.py) and the script that pins the
former against the latter (make_golden.py).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product (msclip_b200/) never does.
"""
