#!/usr/bin/env python
"""Write a synthetic graph of a named benchmark shape as a scipy-style .npz (float32 CSR), the
input format of the benchmark drivers: python tools/make_dataset.py <c1|c2|c3|c4|c5|tiny> <out.npz> [scale]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graphlily_b200 import datasets, io  # noqa: E402

SHAPES = {
    "c1": lambda s, dev: datasets.uniform_csr(10000, 10000, 10, seed=1),
    "c2": lambda s, dev: datasets.powerlaw_csr(int(4_194_304 * s), int(4_194_304 * s), int(134_217_728 * s), seed=42, device=dev),
    "c3": lambda s, dev: datasets.c3_gplus(s, device=dev),
    "c4": lambda s, dev: datasets.c4_ogbn_products(s, device=dev),
    "c5": lambda s, dev: datasets.c5_orkut(s, device=dev),
    "tiny": lambda s, dev: datasets.powerlaw_graph(4096, 60000, seed=2),
}

if __name__ == "__main__":
    import torch
    name, out = sys.argv[1], sys.argv[2]
    scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    m = SHAPES[name](scale, dev)
    io.save_csr_matrix_to_npz(out, m, compressed=False)
    print(f"{name}: {m.num_rows} x {m.num_cols}, nnz {m.nnz} -> {out}")
