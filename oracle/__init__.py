"""ORACLE - test infrastructure only (see the headers of reference_chain.py, agc.py, spectrum.py and
dither_quant.py for what each restates and how it is pinned to the reference's own outputs).

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package never imports it.
"""
