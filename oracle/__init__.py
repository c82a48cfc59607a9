"""CPU ORACLE — test infrastructure, not product code.

A restatement of the reference's algorithm for the observation hot path, used
only as the checker by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package
(isaacgyminsertion_b200/) never imports anything from here.
"""
