"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the MapNet training hot path.

Nothing under ``geomapnet_b200/`` (the product) may import this package.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg use it, and only as the checker / the timed CPU arm.

Contents
  ref_loader.py     executes the *reference's own files* from /root/reference
                    (this container only) with Python-2 shims; used to pin the
                    restatement and to generate tests/golden/*.npz.
  weights.py        deterministic, seed-driven weights shared by both sides.
  mapnet_oracle.py  the CPU restatement (torch fp32/fp64 functional code) of
                    the reference algorithm; every function cites file:line.
  make_goldens.py   regenerates tests/golden/ from the reference (committed
                    script, per the parity contract).

Parity status: the reference ships NO golden vectors or known-answer tests for
this path, so the oracle is pinned against outputs of
the reference itself executed here (ref_loader) -- see tests/test_oracle_*.py.
"""
