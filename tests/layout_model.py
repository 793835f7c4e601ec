"""Sequential numpy model of the warp-segment SpMV schedule (test helper, CPU only).

Executes exactly what ``spmv_ws_kernel`` + ``spmv_fixup_kernel`` (graphlily_b200/csrc/spmv.cu)
do with the arrays ``glb_csr_format_host`` produces -- per-chunk running segments, head /
tail carries, fix-up list, empty rows -- but one non-zero at a time, so the formatter and
the carry protocol can be validated without a GPU.
"""
import numpy as np

FLAG = np.uint32(0x80000000)


def run_model(layout, data, x, op, zero, n_rows_total, shard_sb=0):
    add = {0: lambda a, b: np.float32(a + b), 1: lambda a, b: np.float32(1.0 if (a != 0 or b != 0) else 0.0),
           2: lambda a, b: b if b < a else a}[op]
    mul = {0: lambda a, b: np.float32(a * b), 1: lambda a, b: np.float32(1.0 if (a != 0 and b != 0) else 0.0),
           2: lambda a, b: np.float32(a + b)}[op]
    ident = np.float32(np.inf) if op == 2 else np.float32(0)
    with_zero = {0: lambda z, t: np.float32(z + t), 1: lambda z, t: np.float32(1.0 if (z != 0 or t != 0) else 0.0),
                 2: lambda z, t: t if t < z else z}[op]
    chunk, nnz = int(layout["chunk"]), layout["nnz"]
    cols, nz_rows, cf = layout["cols"], layout["nz_rows"], layout["chunk_first"]
    y = np.full(n_rows_total, np.nan, np.float32)
    written = np.zeros(n_rows_total, np.int32)
    head = np.full(layout["n_chunks"], np.nan, np.float32)
    tail = np.full(layout["n_chunks"], np.nan, np.float32)

    def finish(row, total):
        y[row] = with_zero(np.float32(zero), total)
        written[row] += 1

    for c in range(layout["n_chunks"]):
        ord0, fresh = int(cf[c] & ~FLAG), bool(cf[c] & FLAG)
        ordn, acc = ord0, ident
        for p in range(c * chunk, min((c + 1) * chunk, nnz)):
            w = cols[p]
            if w & FLAG:
                assert p % chunk != 0, "flag at a chunk start"
                if ordn == ord0 and not fresh:
                    head[c] = acc
                else:
                    finish(int(nz_rows[ordn]), acc)
                ordn += 1
                acc = ident
            acc = add(acc, mul(np.float32(data[shard_sb + p]), np.float32(x[int(w & ~FLAG)])))
        tail[c] = acc
    for row, cb, ce in layout["fixups"]:
        has_head = bool(ce & FLAG)
        ce = int(ce & ~FLAG)
        t = ident
        for c in range(int(cb), ce if has_head else ce + 1):
            assert not np.isnan(tail[c])
            t = add(t, tail[c])
        if has_head:
            assert not np.isnan(head[ce])
            t = add(t, head[ce])
        finish(int(row), t)
    for r in layout["empty_rows"]:
        finish(int(r), ident)
    return y, written
