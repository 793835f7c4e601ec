"""Sequential numpy model of the lane-segment SpMV schedule (test helper, CPU only).

Executes exactly what ``spmv_lane_kernel`` + ``spmv_fixup_kernel`` (graphlily_b200/csrc/spmv.cu)
do with the arrays ``glb_csr_format_host`` produces -- per-lane serial runs, staging by row
ordinal, the cross-lane segmented scan, head / tail carries, fix-up list, empty rows -- but one
lane at a time, so the formatter and the carry protocol can be validated without a GPU.
"""
import numpy as np

FLAG = np.uint32(0x80000000)


def semiring(op):
    add = {0: lambda a, b: np.float32(a + b), 1: lambda a, b: np.float32(1.0 if (a != 0 or b != 0) else 0.0),
           2: lambda a, b: b if b < a else a}[op]
    mul = {0: lambda a, b: np.float32(a * b), 1: lambda a, b: np.float32(1.0 if (a != 0 and b != 0) else 0.0),
           2: lambda a, b: np.float32(a + b)}[op]
    ident = np.float32(np.inf) if op == 2 else np.float32(0)
    with_zero = {0: lambda z, t: np.float32(z + t), 1: lambda z, t: np.float32(1.0 if (z != 0 or t != 0) else 0.0),
                 2: lambda z, t: t if t < z else z}[op]
    return add, mul, ident, with_zero


def gather(layout, x, word):
    """hot words index hot_x = x[hot_cols] (or x itself under the identity numbering), cold ones x."""
    tile_k = int(layout["tile_k"])
    if word < tile_k:
        return x[int(layout["hot_cols"][word])] if len(layout["hot_cols"]) else x[word]
    return x[word - tile_k]


def run_model(layout, x, op, zero, n_rows_total):
    add, mul, ident, with_zero = semiring(op)
    stream, flags, goff, cf = layout["stream"], layout["flags"], layout["chunk_goff"], layout["chunk_first"]
    vals = stream.view(np.float32)
    nz_rows = layout["nz_rows"]
    y = np.full(n_rows_total, np.nan, np.float32)
    written = np.zeros(n_rows_total, np.int32)
    head = np.full(layout["n_chunks"], np.nan, np.float32)
    tail = np.full(layout["n_chunks"], np.nan, np.float32)

    def finish(row, total):
        y[row] = with_zero(np.float32(zero), total)
        written[row] += 1

    for c in range(layout["n_chunks"]):
        g0, n = int(goff[c]), int(goff[c + 1] - goff[c])
        assert 1 <= n <= layout["max_groups"]
        ord0, fresh = int(cf[c] & ~FLAG), bool(cf[c] & FLAG)
        cnt = [bin(int(w)).count("1") for w in flags[c]]
        total = sum(cnt)
        assert total <= layout["row_cap"]
        stage = [None] * total
        tails, first_slot = [], []
        k = 0
        for lane in range(32):
            fw, acc = int(flags[c, lane]), ident
            assert fw >> (4 * n) == 0, "flag beyond the lane's run"
            first_slot.append(k if fw else None)
            for r in range(4 * n):
                w = (g0 + (r >> 2)) * 256 + lane * 4 + (r & 3)
                prod = mul(vals[w + 128], np.float32(gather(layout, x, int(stream[w]))))
                if fw >> r & 1:
                    stage[k] = acc
                    k += 1
                    acc = ident
                acc = add(acc, prod)
            tails.append(acc)
        # segmented scan of the lane tails, then each lane's first row end takes the carry from below
        carry = ident
        for lane in range(32):
            if first_slot[lane] is not None:
                stage[first_slot[lane]] = add(carry, stage[first_slot[lane]])
                carry = tails[lane]
            else:
                carry = add(carry, tails[lane])
        tail[c] = carry
        for k2 in range(total):
            if k2 == 0 and not fresh:
                head[c] = stage[0]
            else:
                finish(int(nz_rows[ord0 + k2]), stage[k2])
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
