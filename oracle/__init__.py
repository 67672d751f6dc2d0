"""CPU oracle for the RecAD victim-model hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``recad_b200/`` may import this
package; the only legal callers are ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.

Every function is a plain numpy / scipy / torch-CPU restatement of the
reference algorithm and cites the reference file:line it follows
(paths relative to the reference checkout, e.g. ``recad/dataset/implicit.py``).

Pinning: the reference holds no golden vectors of its own (its test file is
empty), so the oracle is pinned against outputs of the *live* reference,
imported in the build container by ``tests/golden/make_golden.py`` and
committed as fixtures under ``tests/golden/``.  ``tests/test_oracle_golden.py``
replays every fixture through this package.  Recall/NDCG@20 has no reference
implementation at all and is marked "parity unpinned" in
``oracle/evaluate.py``.
"""
