"""CPU oracle for the AnalysisGNN heterogeneous message-passing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker (or
as the timed CPU baseline), never as the thing shipped.  The product package
``analysisgnn_b200`` must not import this package.

Pinning status:

* in-tree layers (``oracle/intree.py``): PINNED.  Checked against the
  reference's own ``analysisgnn/models/core/{gnn,hgnn}.py`` executed in the
  build container through ``oracle/ref_loader.py`` (``torch_scatter`` shim),
  and against the golden vectors that run produced (``tests/golden``).
* score-graph edges (``oracle/graph.py::score_graph_edges``): PINNED against
  the reference's ``utils/hgraph.py:214-300`` body executed standalone.
* PyG / graphmuse shaped operators (``oracle/pyg.py``): **parity unpinned** —
  torch_geometric, torch_scatter, pyg-lib and graphmuse are not installable in
  this environment and the reference's tests hold no golden vectors, so these
  follow the published operator semantics (PyG >= 2.3) restated in SURVEY.md
  §8c / Appendix A.  Every report says "oracle = this repo's restatement".
* samplers (``oracle/graph.py``): defined by this repo (counter-based RNG);
  pyg-lib's RNG stream cannot be reproduced.
"""
