"""CPU oracle — TEST INFRASTRUCTURE ONLY.

A restatement, in plain torch/numpy on the CPU, of what the reference (JBanks/GraphGym) computes on
the message-passing hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package; nothing under ``graphgym_b200/``
does (the product path has no CPU fallback).

Parity status: the reference ships no tests, golden vectors or fixtures for this path,
and its third-party arithmetic (torch_geometric, torch_scatter, tf_geometric) is neither vendored
nor installed.  The oracle is therefore pinned two ways:
  1. ``tests/golden/make_golden.py`` imports the reference's OWN ``graphgym/contrib/layer/idconv.py``,
     ``graphgym/contrib/transform/identity.py`` and ``graphgym/models/transform.py`` from
     /root/reference — with the absent third-party helpers (``MessagePassing``, ``scatter_add``,
     ``add_remaining_self_loops`` ...) supplied by ``oracle.pyg_shim`` — runs them on seeded inputs
     and commits the outputs under ``tests/golden/``; the oracle must reproduce them.
  2. the closed-form known-answer vectors of SURVEY §8(c) (K3 / P3 / C4).
The PyG helper semantics themselves are restated from the PyG 1.x documentation ("parity unpinned"
for those helpers: no PyG build is available to run them against).
"""
