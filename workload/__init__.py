"""Synthetic workloads shared by the parity tests and both arms of bench.py.

Pure torch / numpy: scenes, cameras, loss weights (`synthetic`) and the flat parameter + gradient
bucket of the view-sharded step (`params`).  Nothing here imports a native module, so the
reference arm of bench.py can use it without loading the product's libraries.
"""
