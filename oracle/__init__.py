"""TEST INFRASTRUCTURE.  CPU oracle (plain-C restatement, oracle/aidax_oracle.c) and wrappers around the
unmodified reference compiled into oracle/_ref.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; the product never does."""
