"""CPU oracle package.  TEST INFRASTRUCTURE ONLY - see oracle/README.md.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm.
The product package ``distgcn_b200`` never imports it (tests/test_no_oracle_in_product.py checks).
"""
