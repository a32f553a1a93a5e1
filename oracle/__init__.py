"""CPU oracle for the SequenceAligning hot path -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only.
The product package (sequencealigning_b200) must never import this.
"""
