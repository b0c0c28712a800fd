"""CPU oracle for the DeSCo hot path (canonical partition -> SHMP typing -> SHMP counting -> gossip).

TEST INFRASTRUCTURE ONLY.  Nothing under ``desco_b200/`` imports this package.  The only legal callers are
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, where it
is the checker or the CPU baseline, never the thing measured as the product.

Each function cites the reference ``file:line`` it restates (paths relative to fuvty/DeSCo @ 4508f7a).

Parity pin status (see DESIGN.md "Oracle"):
  * partition  : PINNED  - checked against the reference's own ``k_neigh`` / ``get_neigh_hetero`` /
                 ``get_neigh_canonical`` executed from ``/root/reference/subgraph_counting/data.py`` (tests run the
                 comparison live when the reference tree is present; fixtures in tests/golden/ otherwise).
  * SHMP typing: PINNED to the literal ``A*A@A + A > 1`` sparse formulation of ``transforms.py:201-225`` run in
                 torch here; the reference module itself needs torch_geometric to import.
  * SHMP forward / gossip: the reference ships no tests, golden vectors or checkpoints for these and cannot be
                 imported here (torch_geometric / pytorch_lightning absent).  ``oracle/ref_shim`` runs the reference's
                 OWN ``gnn_model.py`` classes (SAGEConv, GossipConv, BaseGNNCore, BaseGNN) on a ~100-line stand-in for
                 the PyG primitives they call; fixtures generated that way are committed in tests/golden/.  The
                 ``to_hetero`` fx rewrite itself cannot be run, so the hetero expansion is a restatement:
                 "parity unpinned" for that step, stated here and in DESIGN.md.
"""
