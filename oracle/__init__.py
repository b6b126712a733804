"""TEST INFRASTRUCTURE ONLY -- CPU checkers for the NNUE hot path.

Two independent checkers live here:

* ``oracle/_ref/libsp_ref_<isa>.so`` -- the unmodified reference sources compiled with g++
  (recipe: ``oracle/Makefile``), reached through ``oracle/ref_shim.cpp``.
* ``oracle/_build/libsp_oracle.so``  -- a plain-C restatement of the algorithm
  (``oracle/nnue_oracle.c``), every function citing the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this package.  The product (``stormphrax_b200``) must never do so.
"""
