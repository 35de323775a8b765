"""Oracle for the kNN similarity graph the GNN / graph-transformer services build over clip embeddings
(SURVEY.md section 8(f) "next" #1).  numpy restatement of
services/gnn-pipeline/app/main.py:55-100 ``GraphBuilder.compute_knn_edges``:
L2-normalise (+1e-8), full cosine matrix, self excluded with -inf, the k largest per row in ASCENDING similarity order
(``np.argsort(sim_i)[-k:]``), k shrunk to max(1, N - 1) when N <= k.  Tie order is numpy-introsort dependent in the
reference; here ties are DEFINED like the re-ID path: (score desc, index asc) before the ascending flip.
Test infrastructure only.
"""
from __future__ import annotations

import numpy as np


def compute_knn_edges(embeddings: np.ndarray, k: int = 5):
    n = len(embeddings)
    if n <= k:
        k = max(1, n - 1)
    e = embeddings / (np.linalg.norm(embeddings, axis=1, keepdims=True) + 1e-8)
    sim = e @ e.T
    src, dst, w = [], [], []
    for i in range(n):
        s = sim[i].copy()
        s[i] = -np.inf
        order = np.lexsort((np.arange(n), -s))[:k][::-1]      # k best, ascending similarity like argsort()[-k:]
        for j in order:
            if s[j] > -np.inf:
                src.append(i)
                dst.append(int(j))
                w.append(s[j])
    return np.array([src, dst]), np.array(w)


def reference_graph_builder():
    """The reference's OWN GraphBuilder class, extracted from services/gnn-pipeline/app/main.py by AST (the module itself
    cannot be imported here: torch_geometric is absent) and executed unmodified.  Only where /root/reference exists."""
    import ast
    from pathlib import Path
    from typing import Dict, List, Optional, Tuple

    src = Path("/root/reference/services/gnn-pipeline/app/main.py").read_text()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "GraphBuilder")
    ns = {"np": np, "Tuple": Tuple, "List": List, "Optional": Optional, "Dict": Dict}
    exec(compile(ast.Module(body=[node], type_ignores=[]), "gnn_main_GraphBuilder", "exec"), ns)
    return ns["GraphBuilder"]
