"""Seeded synthetic graphs of the shapes BASELINE.json names.

The reference's datasets (gplus, ogbn-products, orkut ... ``benchmark/run_spmv.sh:12-17``)
are not shipped, so every bench / parity input is generated:

* ``uniform_csr``     C1: n x n, exactly k distinct random columns per row
                      (mirrors ``uniform_10K_10`` of tests/test_module_spmv_spmspv.cpp:174)
* ``powerlaw_csr``    C2: power-law row degrees (Pareto, alpha 2.1, 1 .. 2^20, rescaled to the
                      requested nnz) and Zipf-skewed column popularity (s = 0.9) under a
                      random column permutation -- the hard, locality-free case
* ``social_graph``    C3 / C4 / C5: symmetric graph with a truncated power-law degree sequence
                      (configuration model), hub degree capped at what the real datasets show
                      (gplus ~2e4, ogbn-products 17 481, orkut 33 313); ``c3_gplus`` /
                      ``c4_ogbn_products`` / ``c5_orkut`` are the named shapes every bench and test uses
* ``powerlaw_graph``  symmetric graph with one Zipf-drawn endpoint: hubs of 10^5-10^6 edges, far
                      beyond any of the named datasets -- kept as a labelled STRESS case (giant
                      rows: chunk-spanning fix-ups, fp32 summation order), optional full diagonal

Column indices are sorted and distinct inside each row.  Generation runs in torch on the
given device (CPU here, CUDA on the GPU box for the 128 M-nnz cases) because it is only
input plumbing; it is seeded per device type, so a CPU-generated and a GPU-generated matrix
of the same seed differ -- every consumer takes the matrix itself as the input of record.
"""
import numpy as np
import torch

from .io import CSRMatrix


def _gen(seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def _finish(keys, n_rows, n_cols, value, device):
    """Sorted unique int64 keys (row * n_cols + col) -> CSRMatrix on the host."""
    rows = torch.div(keys, n_cols, rounding_mode="floor")
    cols = keys - rows * n_cols
    counts = torch.bincount(rows, minlength=n_rows)
    indptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=keys.device)
    indptr[1:] = torch.cumsum(counts, 0)
    nnz = int(keys.numel())
    assert nnz < 2**32
    return CSRMatrix(n_rows, n_cols, np.full(nnz, value, np.float32), cols.to(torch.int32).cpu().numpy().view(np.uint32),
                     indptr.to(torch.int32).cpu().numpy().view(np.uint32) if nnz < 2**31 else
                     indptr.cpu().numpy().astype(np.uint32))


def uniform_csr(n_rows, n_cols, nnz_per_row, seed=0, value=None):
    """Exactly ``nnz_per_row`` distinct uniformly random columns in every row."""
    rng = np.random.default_rng(seed)
    k = nnz_per_row
    cols = np.empty((n_rows, k), np.int64)
    # rejection-free: sample k distinct of n_cols per row via sorted random keys
    for start in range(0, n_rows, 65536):
        stop = min(n_rows, start + 65536)
        if n_cols <= 4096:
            cols[start:stop] = np.argsort(rng.random((stop - start, n_cols)), axis=1)[:, :k]
        else:
            c = rng.integers(0, n_cols, size=(stop - start, k))
            while True:
                c.sort(axis=1)
                dup = np.zeros_like(c, bool)
                dup[:, 1:] = c[:, 1:] == c[:, :-1]
                if not dup.any():
                    break
                c[dup] = rng.integers(0, n_cols, size=int(dup.sum()))
            cols[start:stop] = c
    cols.sort(axis=1)
    indptr = (np.arange(n_rows + 1, dtype=np.uint64) * k).astype(np.uint32)
    val = np.float32(1.0 / n_rows) if value is None else np.float32(value)
    return CSRMatrix(n_rows, n_cols, np.full(n_rows * k, val, np.float32), cols.reshape(-1).astype(np.uint32), indptr)


def _zipf_ranks(n, count, s, gen, device):
    """`count` samples of a continuous power law P(rank) ~ rank^-s on [1, n], as 0-based ints."""
    u = torch.rand(count, generator=gen, device=device, dtype=torch.float64)
    a = 1.0 - s
    r = torch.pow(u * (float(n) ** a - 1.0) + 1.0, 1.0 / a)
    return torch.clamp(r.to(torch.int64) - 1, 0, n - 1)


def powerlaw_csr(n_rows, n_cols, nnz, seed=42, alpha=2.1, max_degree=1 << 20, col_skew=0.9, value=None,
                 device="cpu"):
    """C2-shaped matrix: power-law row degrees, Zipf column popularity, random column labels.

    Exactly ``nnz`` non-zeros after de-duplication (short of pathological requests)."""
    device = torch.device(device)
    gen = _gen(seed, device)
    # row degrees: Pareto(alpha) >= 1, truncated, rescaled to the requested total
    u = torch.rand(n_rows, generator=gen, device=device, dtype=torch.float64)
    deg = torch.clamp(torch.pow(1.0 - u, -1.0 / (alpha - 1.0)), max=float(max_degree))
    col_perm = torch.randperm(n_cols, generator=gen, device=device)
    keys = torch.empty(0, dtype=torch.int64, device=device)
    want = nnz
    for _ in range(8):
        scale = want * 1.08 / float(deg.sum())
        d = torch.clamp(torch.round(deg * scale), min=1, max=min(max_degree, n_cols)).to(torch.int64)
        rows = torch.repeat_interleave(torch.arange(n_rows, device=device), d)
        cols = col_perm[_zipf_ranks(n_cols, rows.numel(), col_skew, gen, device)]
        keys = torch.unique(torch.cat([keys, rows * n_cols + cols]))  # sorted, distinct
        if keys.numel() >= nnz:
            break
        want = max(nnz - keys.numel(), 1024)
    if keys.numel() > nnz:  # drop a uniform random surplus
        keep = torch.randperm(keys.numel(), generator=gen, device=device)[:nnz]
        keys = keys[torch.sort(keep).values]
    val = np.float32(1.0 / n_rows) if value is None else np.float32(value)
    return _finish(keys, n_rows, n_cols, val, device)


def powerlaw_graph(n, nnz, seed=42, skew=0.9, diagonal=False, value=1.0, device="cpu"):
    """Symmetric power-law graph (C3 gplus-, C4 ogbn-products-, C5 orkut-shaped).

    Both endpoints of every edge are drawn from a Zipf popularity under a random vertex
    relabelling, the pattern is symmetrised and de-duplicated; ``nnz`` is approximate
    (within a few percent).  ``diagonal=True`` adds every self-loop (SSSP, sssp.h:16-62)."""
    device = torch.device(device)
    gen = _gen(seed, device)
    perm = torch.randperm(n, generator=gen, device=device)
    keys = torch.empty(0, dtype=torch.int64, device=device)
    want = nnz // 2
    for _ in range(8):
        m = int(want * 1.05)
        u = perm[_zipf_ranks(n, m, skew, gen, device)]
        v = torch.randint(0, n, (m,), generator=gen, device=device)
        ok = u != v
        u, v = u[ok], v[ok]
        keys = torch.unique(torch.cat([keys, u * n + v, v * n + u]))
        if keys.numel() >= nnz * 0.98:
            break
        want = max((nnz - keys.numel()) // 2, 1024)
    if diagonal:
        d = torch.arange(n, device=device, dtype=torch.int64)
        keys = torch.unique(torch.cat([keys, d * n + d]))
    return _finish(keys, n, n, np.float32(value), device)


def social_graph(n, nnz, max_degree, seed=42, alpha=2.1, diagonal=False, value=1.0, device="cpu"):
    """Symmetric graph whose degrees follow a Pareto(alpha) law truncated at ``max_degree``.

    Configuration model: vertex u gets d[u] stubs (d ~ Pareto, rescaled so that the stubs sum to
    nnz / 2, capped at max_degree / 2, at least 1), the stub list is paired with a random
    permutation of itself, self-pairs are dropped, the pattern is symmetrised and de-duplicated --
    so a vertex ends up with about 2 d[u] <= max_degree neighbours and ``nnz`` is met within a few
    percent.  Degrees are independent of the vertex id (no relabelling needed)."""
    device = torch.device(device)
    gen = _gen(seed, device)
    u = torch.rand(n, generator=gen, device=device, dtype=torch.float64)
    pareto = torch.pow(1.0 - u, -1.0 / (alpha - 1.0))
    cap = max(int(max_degree) // 2, 1)
    target = nnz / 2 * 1.02
    scale = target / float(pareto.sum())
    for _ in range(6):   # the cap removes mass from the tail: re-fit the scale
        d = torch.clamp(torch.round(pareto * scale), min=1, max=cap)
        scale *= target / float(d.sum())
    d = torch.clamp(torch.round(pareto * scale), min=1, max=cap).to(torch.int64)
    stubs = torch.repeat_interleave(torch.arange(n, device=device), d)
    partner = stubs[torch.randperm(stubs.numel(), generator=gen, device=device)]
    ok = stubs != partner
    a, b = stubs[ok], partner[ok]
    keys = torch.unique(torch.cat([a * n + b, b * n + a]))
    if diagonal:
        dg = torch.arange(n, device=device, dtype=torch.int64)
        keys = torch.unique(torch.cat([keys, dg * n + dg]))
    return _finish(keys, n, n, np.float32(value), device)


def _pad128(n):
    return (int(n) + 127) // 128 * 128


def c3_gplus(scale=1.0, device="cpu"):
    """bench_bfs shape: 107 648 vertices, ~13 M nnz, hub degree <= 20 000."""
    # (the stub target is above 13 M: with hubs of 10^4 on 10^5 vertices one pairing in eight is a duplicate)
    return social_graph(_pad128(107_648 * scale), int(14_700_000 * scale), 20_000, seed=3, device=device)


def c4_ogbn_products(scale=1.0, device="cpu"):
    """bench_pagerank shape (C4): 2 449 024 vertices, ~124 M nnz, hub degree <= 17 481 (the real graph's maximum)."""
    return social_graph(_pad128(2_449_024 * scale), int(124_000_000 * scale), 17_481, seed=4, device=device)


def c5_orkut(scale=1.0, device="cpu"):
    """bench_sssp shape (C5): 3 072 512 vertices, ~117 M nnz + the full diagonal, hub degree <= 33 313."""
    return social_graph(_pad128(3_072_512 * scale), int(117_000_000 * scale), 33_313, seed=5, diagonal=True, device=device)


def line_graph(n):
    """tests/test_data/line_8_csr_float32.npz shape: sub-diagonal chain, A[i, i-1] = 1."""
    indptr = np.concatenate([[0], np.arange(0, n)]).astype(np.uint32)
    return CSRMatrix(n, n, np.ones(n - 1, np.float32), np.arange(n - 1, dtype=np.uint32), indptr)


def eye(n):
    """tests/test_data/eye_10_csr_float32.npz shape."""
    return CSRMatrix(n, n, np.ones(n, np.float32), np.arange(n, dtype=np.uint32), np.arange(n + 1, dtype=np.uint32))
