"""ORACLE - test infrastructure only (see reference_chain.py / dither_quant.py headers).

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package never imports it.
"""
