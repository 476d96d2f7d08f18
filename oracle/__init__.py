"""ORACLE -- test infrastructure only.

CPU restatement of the reference's NLP-callback hot path. Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; the product package never does.
"""
