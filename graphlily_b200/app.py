"""Graph algorithms: the Python mirror of ``graphlily::app`` (BFS, PageRank, SSSP).

Follows ``/root/reference/graphlily/app/{module_collection,bfs,pagerank,sssp}.h``: same class
and method names, same pre-processing, same buffer aliasing, same results.  Two execution
modes per pull loop:

* ``fused=False`` -- the reference's launch sequence literally (``SpMV.run(); eWiseAdd.run();
  DenseAssign.run()`` per iteration, ``bfs.h:117-124``);
* ``fused=True`` (default) -- one ``glb_spmv_fused`` launch per iteration: the eWiseAdd and the
  dense assign ride in the SpMV write-back and the results / vector buffers ping-pong instead
  of being copied.

The push->pull switch of ``pull_push`` converts the sparse frontier to a dense vector on the
device (``glb_sparse_to_dense``) instead of the reference's host round trip (``bfs.h:195-201``).
``load_and_format_matrix`` takes an ``.npz`` path like the reference, or a ``CSRMatrix``.
"""
import copy

import numpy as np

from . import capi, io
from .capi import Epilogue, SpmspvEpilogue, SpmspvNext
from .module import (AssignVectorDenseModule, AssignVectorSparseModule, SpMSpVModule, SpMVModule, eWiseAddModule)

ArithmeticSemiring = (capi.OP_MUL_ADD, 1.0, 0.0)
LogicalSemiring = (capi.OP_LOGICAL_AND_OR, 1.0, 0.0)
TropicalSemiring = (capi.OP_ADD_MIN, 0.0, 255.0)
PACK = 8  # graphlily::pack_size, global.h:57


class ModuleCollection:
    """module_collection.h:15-114"""

    def __init__(self):
        self.modules_ = []
        self.target_ = "hw"
        self.ctx = None

    def add_module(self, module):
        self.modules_.append(module)

    def set_target(self, target):
        assert target in ("sw_emu", "hw_emu", "hw")
        self.target_ = target

    def set_up_runtime(self, xclbin_file_path=None, device=0, ctx=None):
        """One context (device + stream) shared by all modules; the bitstream path is ignored."""
        self.ctx = ctx or capi.Context(device)
        for m in self.modules_:
            m.set_context(self.ctx)

    def set_pinned_results(self, on=True):
        """Result vectors are returned as the modules' page-locked host mirrors (valid until the next
        run) instead of fresh copies: see BaseModule.pinned_results."""
        for m in self.modules_:
            m.pinned_results = bool(on)

    # -- multi-GPU: row-range sharding of the pull (SpMV) direction, SURVEY.md 8e -------------
    rank_, world_ = 0, 1

    exchange_ = None

    def set_sharding(self, rank, world, exchange=None):
        """This process owns rows [rank, rank + 1) * N / world of the CSR.  How the slices of y meet
        before the next iteration:
        * ``exchange`` = a connected ``capi.Exchange`` (>= 3 vectors of N floats): the SpMV
          write-back stores every row into all ranks' copies over NVLink (``glb_spmv_exchange``);
        * else the context must hold an NCCL communicator of ``world`` ranks (``ctx.comm_init``) and
          one in-place allgather follows every SpMV."""
        assert 0 <= rank < world
        self.rank_, self.world_, self.exchange_ = rank, world, exchange

    def _cuts(self, n):
        """Row cuts of the ranks.  Over an exchange the slices need not be equal, so they are cut where
        the nnz prefix of the CSR crosses r / world (32-row aligned): every rank streams the same
        share of the matrix.  The NCCL allgather needs equal slots."""
        cuts = self.__dict__.get("cuts_")
        if cuts is not None and cuts[-1] == n and len(cuts) == self.world_ + 1:
            return cuts
        w = self.world_
        if w == 1:
            cuts = [0, n]
        elif self.exchange_ is None:
            assert n % w == 0, "padded dimension must divide by the number of ranks"
            cuts = [r * (n // w) for r in range(w + 1)]
        else:
            ip = np.asarray(self.csr_matrix_.indptr, dtype=np.int64)
            cuts = [int(np.searchsorted(ip, ip[-1] * r // w)) // 32 * 32 for r in range(w + 1)]
            cuts[0], cuts[-1] = 0, n
            for r in range(1, w + 1):
                cuts[r] = max(cuts[r], cuts[r - 1])
        self.cuts_ = cuts
        return cuts

    def _row_range(self, n):
        cuts = self._cuts(n)
        return cuts[self.rank_], cuts[self.rank_ + 1]

    def _bind_exchange(self):
        """After the matrix upload: vector / results / mask of the SpMV module become exchange vectors."""
        xc = self.exchange_
        if xc is None:
            return
        assert xc.n == self.matrix_num_rows_ and xc.n_vectors >= 3, "exchange must hold 3 vectors of the padded dimension"
        self.SpMV_.exchange = xc
        self.SpMV_.vector_buf, self.SpMV_.results_buf, self.SpMV_.mask_buf = xc.buffer(0), xc.buffer(1), xc.buffer(2)
        if hasattr(self, "SpMSpV_"):   # the distance vector both directions update lives in the exchange too
            self.SpMSpV_.mask_buf = xc.buffer(2)

    def _matrix_changed(self):
        """A new matrix went to the device: recorded launch sequences of the old one are void."""
        self.__dict__.pop("graphs_", None)
        self.__dict__.pop("cuts_", None)

    def _begin_run(self):
        """No rank starts writing into the peers' vectors before every rank has finished reading the
        previous run's results out of them."""
        if self.exchange_ is not None:
            self.exchange_.barrier()

    def _exchange(self, buf, n):
        """After every SpMV.  With peer-mapped vectors the exchange already happened in the write-back."""
        if self.world_ > 1 and self.exchange_ is None:
            self.ctx.allgather_f32(buf, n // self.world_)

    def _gather(self, buf, n):
        """Completes a vector whose rows were updated shard-locally (BFS distance, at the end)."""
        if self.world_ == 1:
            return
        if self.exchange_ is not None:
            rb, re = self._row_range(n)
            self.exchange_.allgather(buf.tag[1], rb, re - rb)
        else:
            self.ctx.allgather_f32(buf, n // self.world_)

    def _push_begin(self):
        """Row-sharded push: the dense vectors the frontier travels in are the SpMV module's vector and
        results buffers (both idle while the push direction runs)."""
        if self.world_ > 1:
            n = self.matrix_num_rows_
            self.SpMV_.home_buffers()
            if self.SpMV_.vector_buf is None or self.SpMV_.vector_buf.nbytes != 4 * n:
                self.SpMV_.vector_buf = self.ctx.alloc(4 * n)
            if self.SpMV_.results_buf is None or self.SpMV_.results_buf.nbytes != 4 * n:
                self.SpMV_.results_buf = self.ctx.alloc(4 * n)

    def _exchange_frontier(self, list_buf, zero):
        """Row-sharded push: every rank's SpMSpV listed only the rows it owns.  The
        lists meet as ONE dense vector -- own slice reset and scattered, slices exchanged, the full
        vector listed again on every rank (into the same list buffer) -- so what follows (the sparse
        assign on the replicated distance vector, the next SpMSpV) sees the whole frontier.

        The dense vector ALTERNATES between two buffers: over a peer / multicast exchange the
        signal / wait step only says "every rank has written its slice of step k", not "every rank
        has finished reading step k", so a fast rank's slice of step k + 1 must not land in the
        vector a slow rank is still listing.  With two vectors a rank can only be overwritten two
        steps later, which its own signal of step k + 1 (issued after its listing of step k, in
        stream order) gates.  After the call the fresh frontier is ``SpMV_.vector_buf``."""
        if self.world_ == 1:
            return
        n = self.matrix_num_rows_
        dense = self.SpMV_.results_buf
        rb, re = self._row_range(n)
        capi.sparse_to_dense_rows(self.ctx, list_buf, dense, rb, re, zero)
        self._gather(dense, n)
        capi.dense_to_sparse(self.ctx, dense, n, zero, list_buf)
        self.SpMV_.vector_buf, self.SpMV_.results_buf = dense, self.SpMV_.vector_buf

    # -- push direction on one GPU: a level is ONE launch (glb_spmspv_fused) -----------------------
    def _home_lists(self):
        """Canonical roles of the two sparse-list buffers at the start of a run (lower address = vector),
        so that the recorded sequence of an earlier run with the same arguments is found again."""
        m = self.SpMSpV_
        if m.vector_buf.ptr > m.results_buf.ptr:
            m.vector_buf, m.results_buf = m.results_buf, m.vector_buf

    def _dense_pair(self, fill):
        """The two dense vectors the pull levels ping-pong through (the SpMV module's vector / results),
        in canonical order, filled with ``fill``: the push level that stops pushing builds the input of
        the first pull level in one of them."""
        n, sp = self.matrix_num_rows_, self.SpMV_
        for name in ("vector_buf", "results_buf"):
            b = getattr(sp, name)
            if b is None or not b.ptr or b.nbytes != 4 * n:
                setattr(sp, name, self.ctx.alloc(4 * n))
        sp.home_buffers()
        for b in (sp.vector_buf, sp.results_buf):
            capi.check(capi.lib.glb_buffer_fill_f32(self.ctx.handle, b.ptr, float(fill), n))
        return [sp.vector_buf, sp.results_buf]

    @property
    def push_iterations_(self):
        """Levels the last pull_push run spent pushing.  With the decision taken on the device it is read
        back on demand (a blocking 16-byte copy), not on the critical path of the run."""
        v = self.__dict__.get("push_iterations_host_")
        if v is not None:
            return v
        levels = self.SpMSpV_.matrix.push_state()[1]
        return levels if self.__dict__.get("last_num_iterations_", 2) >= 2 else 1

    # -- launch replay: the iteration loop of an app is a fixed launch sequence ------------------
    use_graphs_ = True
    _GRAPH_CACHE = 8

    def _replay(self, key, launches):
        """Run ``launches()`` -- a loop of glb_* launches over fixed buffers and scalars -- as one
        recorded CUDA graph (recorded at the first call with this ``key``, replayed afterwards): a
        7-iteration BFS is 21 launches of 5-50 us kernels, which the host cannot enqueue fast enough
        one by one.  Row-sharded runs over an exchange are recorded too (the exchange's epoch lives in
        device memory); with the NCCL allgather the launches are issued one by one."""
        if not self.use_graphs_ or (self.world_ > 1 and self.exchange_ is None):
            launches()
            return
        # a recorded sequence holds the device addresses of the matrix it was recorded with
        key = key + (self.SpMV_.matrix.handle.value,)
        cache = self.__dict__.setdefault("graphs_", {})
        g = cache.get(key)
        if g is None:
            if len(cache) >= self._GRAPH_CACHE:
                cache.pop(next(iter(cache)))
            g = cache[key] = self.ctx.record(launches)
        g.launch()


def _load(path_or_csr):
    if isinstance(path_or_csr, str):
        return io.load_csr_matrix_from_float_npz(path_or_csr)
    return copy.deepcopy(path_or_csr)


class BFS(ModuleCollection):
    """bfs.h:20-360"""

    def __init__(self, num_channels=16, spmv_out_buf_len=0, spmspv_out_buf_len=0, vec_buf_len=0):
        super().__init__()
        self.num_channels_ = num_channels
        self.semiring_ = LogicalSemiring
        self.SpMV_ = SpMVModule(num_channels, spmv_out_buf_len, vec_buf_len)
        self.SpMV_.set_semiring(self.semiring_)
        self.SpMV_.set_mask_type(capi.MASK_WRITE_TO_ZERO)
        self.add_module(self.SpMV_)
        self.DenseAssign_ = AssignVectorDenseModule()
        self.DenseAssign_.set_mask_type(capi.MASK_WRITE_TO_ONE)
        self.add_module(self.DenseAssign_)
        self.SpMSpV_ = SpMSpVModule(spmspv_out_buf_len)
        self.SpMSpV_.set_semiring(self.semiring_)
        self.SpMSpV_.set_mask_type(capi.MASK_WRITE_TO_ZERO)
        self.add_module(self.SpMSpV_)
        self.SparseAssign_ = AssignVectorSparseModule(False)
        self.add_module(self.SparseAssign_)
        self.eWiseAdd_ = eWiseAddModule()
        self.add_module(self.eWiseAdd_)

    def get_nnz(self):
        return self.SpMV_.get_nnz()

    def load_and_format_matrix(self, csr_float_npz_path, skip_empty_rows=True):
        """bfs.h:84-97: round dims to 128, all values 1, CSC for the push direction."""
        m = _load(csr_float_npz_path)
        io.util_round_csr_matrix_dim(m, self.num_channels_ * PACK, self.num_channels_ * PACK)
        m.data = np.ones(m.nnz, np.float32)
        self.csr_matrix_ = m
        self.SpMV_.load_and_format_matrix(m, skip_empty_rows)
        self.SpMSpV_.load_and_format_matrix(io.csr2csc(m))
        self.matrix_num_rows_, self.matrix_num_cols_ = m.num_rows, m.num_cols
        assert m.num_rows == m.num_cols

    def send_matrix_host_to_device(self):
        self._matrix_changed()
        self.SpMV_.send_matrix_host_to_device(*self._row_range(self.matrix_num_rows_))
        self._bind_exchange()
        self.SpMSpV_.send_matrix_host_to_device(*self._row_range(self.matrix_num_rows_))

    # -- pull ------------------------------------------------------------------------
    def _pull_loop(self, first_iter, num_iterations, fused):
        n = self.matrix_num_rows_
        if fused:
            vec, res, mask = self.SpMV_.vector_buf, self.SpMV_.results_buf, self.SpMV_.mask_buf
            iters = range(first_iter, num_iterations + 1)

            def launches():
                eps = [Epilogue(0, 0.0, mask.ptr, float(it + 1), capi.MASK_WRITE_TO_ONE) for it in iters]
                if self.exchange_ is not None:   # the whole loop in one call: no acquire launch between steps
                    self.SpMV_.iterate_with(vec, mask, res, eps)
                    return
                v, r = vec, res
                for ep in eps:
                    self.SpMV_.run_with(v, mask, r, ep)
                    self._exchange(r, n)   # distance stays row-local until the end
                    v, r = r, v

            self._replay(("bfs", first_iter, num_iterations, vec.ptr, res.ptr, mask.ptr), launches)
            if len(iters) % 2:
                self.SpMV_.vector_buf, self.SpMV_.results_buf = res, vec
        else:
            assert self.world_ == 1, "the unfused launch sequence is single-GPU"
            self.DenseAssign_.bind_mask_buf(self.SpMV_.vector_buf)
            self.DenseAssign_.bind_inout_buf(self.SpMV_.mask_buf)
            self.eWiseAdd_.bind_in_buf(self.SpMV_.results_buf)
            self.eWiseAdd_.bind_out_buf(self.SpMV_.vector_buf)
            for it in range(first_iter, num_iterations + 1):
                self.SpMV_.run()
                self.eWiseAdd_.run(n, 0.0)
                self.DenseAssign_.run(n, float(it + 1))

    def pull(self, source, num_iterations, fused=True):
        """bfs.h:106-126"""
        n = self.matrix_num_rows_
        self._begin_run()
        # input = zero but input[source] = 1; distance = 0 but distance[source] = 1 (bfs.h:108-112),
        # built on the device instead of uploaded
        self.SpMV_.set_vector_constant(self.semiring_[2], source, 1.0)
        self.SpMV_.set_mask_constant(0.0, source, 1.0)
        self._pull_loop(1, num_iterations, fused)
        self._gather(self.SpMV_.mask_buf, n)
        return self.SpMV_.send_mask_device_to_host()

    # -- push ------------------------------------------------------------------------
    def _push_setup(self, source):
        self._begin_run()
        self._push_begin()
        self._home_lists()
        self.SpMSpV_.set_vector_single(source, 1.0)
        self.SpMSpV_.set_mask_constant(0.0, source, 1.0)      # distance, bfs.h:138-141
        self.SparseAssign_.bind_inout_buf(self.SpMSpV_.mask_buf)

    def _push_step(self, it):
        # SpMSpV.run(); results -> vector (the reference copies 1+nnz entries, bfs.h:147-151;
        # here the two list buffers swap roles); SparseAssign.run(iter + 1)
        self.SpMSpV_.run()
        self._exchange_frontier(self.SpMSpV_.results_buf, self.semiring_[2])
        self.SpMSpV_.vector_buf, self.SpMSpV_.results_buf = self.SpMSpV_.results_buf, self.SpMSpV_.vector_buf
        self.SparseAssign_.bind_mask_buf(self.SpMSpV_.vector_buf)
        self.SparseAssign_.run(float(it + 1))

    def _push_level_fused(self, lists, level, nxt=None):
        """One push level in one launch: SpMSpV on lists[(level - 1) & 1] -> lists[level & 1] with the
        sparse assign distance[row] = level + 1 (bfs.h:147-151) fused into the kernel."""
        dist = self.SpMSpV_.mask_buf
        ep = SpmspvEpilogue(capi.SPMSPV_EP_ASSIGN, dist.ptr, float(level + 1), None)
        self.SpMSpV_.run_with(lists[(level - 1) & 1], lists[level & 1], ep, nxt)

    def push(self, source, num_iterations, fused=True):
        """bfs.h:129-157"""
        self._push_setup(source)
        if not fused or self.world_ > 1:
            for it in range(1, num_iterations + 1):
                self._push_step(it)
            return self.SpMSpV_.send_mask_device_to_host()
        lists = [self.SpMSpV_.vector_buf, self.SpMSpV_.results_buf]

        def launches():
            for level in range(1, num_iterations + 1):
                self._push_level_fused(lists, level)

        self._replay(("bfs_push", num_iterations, lists[0].ptr, lists[1].ptr, self.SpMSpV_.mask_buf.ptr), launches)
        self.SpMSpV_.vector_buf, self.SpMSpV_.results_buf = lists[num_iterations & 1], lists[(num_iterations + 1) & 1]
        return self.SpMSpV_.send_mask_device_to_host()

    def pull_push(self, source, num_iterations, threshold=0.05, fused=True):
        """bfs.h:160-219: push while the frontier is sparse, then pull.

        On one GPU (fused, launch replay on) the whole run is ONE recorded sequence with no host
        synchronisation between the start vectors and the read-back of the result: every push level is
        one launch that also takes the reference's decision (bfs.h:186-190) on the device and feeds it
        to the IF / ELSE node of the next level; the level that stops pushing scatters its frontier
        into the dense input of the first pull level.  Otherwise (row-sharded, unfused, replay off)
        the host reads the frontier size after every push level like the reference does."""
        self.last_num_iterations_ = num_iterations
        if fused and self.world_ == 1 and self.use_graphs_ and num_iterations >= 2:
            return self._pull_push_device(source, num_iterations, threshold)
        n = self.matrix_num_rows_
        self._push_setup(source)
        it = 1
        while True:
            self._push_step(it)
            vector_nnz = capi.sparse_count(self.ctx, self.SpMSpV_.vector_buf)
            it += 1
            if not (it < num_iterations and float(vector_nnz) / n < threshold):
                break
        self.push_iterations_host_ = it - 1
        # switch: the last frontier becomes the dense SpMV input, on the device
        self.SpMV_.bind_mask_buf(self.SpMSpV_.mask_buf)
        if self.world_ == 1:   # (sharded: the frontier already sits in vector_buf, it travelled as a dense vector)
            self.SpMV_.home_buffers()
            if self.SpMV_.vector_buf is None or self.SpMV_.vector_buf.nbytes < 4 * n:
                self.SpMV_.vector_buf = self.ctx.zeros_f32(n)
            capi.sparse_to_dense(self.ctx, self.SpMSpV_.vector_buf, self.SpMV_.vector_buf, n, LogicalSemiring[2])
        self._pull_loop(it, num_iterations, fused)
        self._gather(self.SpMSpV_.mask_buf, n)    # the pull levels updated the distance shard by shard
        return self.SpMSpV_.send_mask_device_to_host()

    def _pull_push_device(self, source, num_iterations, threshold):
        n, ctx = self.matrix_num_rows_, self.ctx
        self.push_iterations_host_ = None
        self._push_setup(source)
        dist = self.SpMSpV_.mask_buf
        self.SpMV_.bind_mask_buf(dist)
        dense = self._dense_pair(LogicalSemiring[2])
        lists = [self.SpMSpV_.vector_buf, self.SpMSpV_.results_buf]
        self.SpMSpV_.matrix.reset_levels()

        def launches():
            conds = {level: ctx.cond_create() for level in range(2, num_iterations + 1)}

            def push_level(level):
                nxt = None
                if level < num_iterations:   # a level follows: this launch decides its direction
                    nxt = SpmspvNext(int(level + 1 >= num_iterations), float(threshold), n, conds[level + 1],
                                     capi.SPMSPV_DENSE_SCATTER, dense[(level + 1) & 1].ptr, None, n)
                self._push_level_fused(lists, level, nxt)

            def pull_level(level):
                ep = Epilogue(0, 0.0, dist.ptr, float(level + 1), capi.MASK_WRITE_TO_ONE)
                self.SpMV_.run_with(dense[level & 1], dist, dense[(level + 1) & 1], ep)

            push_level(1)
            for level in range(2, num_iterations + 1):
                ctx.branch(conds[level], lambda lv=level: push_level(lv), lambda lv=level: pull_level(lv))

        self._replay(("bfs_pull_push", num_iterations, float(threshold), lists[0].ptr, lists[1].ptr, dist.ptr,
                      dense[0].ptr, dense[1].ptr), launches)
        return self.SpMSpV_.send_mask_device_to_host()


class PageRank(ModuleCollection):
    """pagerank.h:17-159"""

    def __init__(self, num_channels=16, spmv_out_buf_len=0, vec_buf_len=0):
        super().__init__()
        self.num_channels_ = num_channels
        self.semiring_ = ArithmeticSemiring
        self.SpMV_ = SpMVModule(num_channels, spmv_out_buf_len, vec_buf_len)
        self.SpMV_.set_semiring(self.semiring_)
        self.SpMV_.set_mask_type(capi.MASK_NONE)
        self.add_module(self.SpMV_)
        self.eWiseAdd_ = eWiseAddModule()
        self.add_module(self.eWiseAdd_)

    def get_nnz(self):
        return self.SpMV_.get_nnz()

    def load_and_format_matrix(self, csr_float_npz_path, damping, skip_empty_rows=True):
        """pagerank.h:60-73: round dims, 1/colcount (double divide -> float), * damping in float."""
        m = _load(csr_float_npz_path)
        io.util_round_csr_matrix_dim(m, self.num_channels_ * PACK, self.num_channels_ * PACK)
        io.util_normalize_csr_matrix_by_outdegree(m)
        m.data = (m.data * np.float32(damping)).astype(np.float32)
        self.csr_matrix_ = m
        self.SpMV_.load_and_format_matrix(m, skip_empty_rows)
        self.matrix_num_rows_, self.matrix_num_cols_ = m.num_rows, m.num_cols
        assert m.num_rows == m.num_cols

    def send_matrix_host_to_device(self):
        self._matrix_changed()
        self.SpMV_.send_matrix_host_to_device(*self._row_range(self.matrix_num_rows_))
        self._bind_exchange()

    def pull(self, damping, num_iterations, fused=True):
        """pagerank.h:80-90"""
        n = self.matrix_num_rows_
        self._begin_run()
        teleport = float((np.float32(1) - np.float32(damping)) / np.float32(n))
        self.SpMV_.set_vector_constant(float(np.float32(1.0 / n)))   # rank0 = 1 / N, pagerank.h:81-82
        if fused:
            vec, res = self.SpMV_.vector_buf, self.SpMV_.results_buf

            def launches():
                if self.exchange_ is not None:
                    self.SpMV_.iterate_with(vec, None, res, [Epilogue(1, teleport, None, 0.0, 0)] * num_iterations)
                    return
                v, r = vec, res
                for _ in range(num_iterations):
                    self.SpMV_.run_with(v, None, r, Epilogue(1, teleport, None, 0.0, 0))
                    self._exchange(r, n)
                    v, r = r, v

            self._replay(("pagerank", teleport, num_iterations, vec.ptr, res.ptr), launches)
            if num_iterations % 2:
                self.SpMV_.vector_buf, self.SpMV_.results_buf = res, vec
        else:
            assert self.world_ == 1, "the unfused launch sequence is single-GPU"
            self.eWiseAdd_.bind_in_buf(self.SpMV_.results_buf)
            self.eWiseAdd_.bind_out_buf(self.SpMV_.vector_buf)
            for _ in range(num_iterations):
                self.SpMV_.run()
                self.eWiseAdd_.run(n, teleport)
        return self.SpMV_.send_vector_device_to_host()


class SSSP(ModuleCollection):
    """sssp.h:68-253"""

    def __init__(self, num_channels=16, spmv_out_buf_len=0, spmspv_out_buf_len=0, vec_buf_len=0):
        super().__init__()
        self.num_channels_ = num_channels
        self.semiring_ = TropicalSemiring
        self.SpMV_ = SpMVModule(num_channels, spmv_out_buf_len, vec_buf_len)
        self.SpMV_.set_semiring(self.semiring_)
        self.SpMV_.set_mask_type(capi.MASK_NONE)
        self.add_module(self.SpMV_)
        self.SpMSpV_ = SpMSpVModule(spmspv_out_buf_len)
        self.SpMSpV_.set_semiring(self.semiring_)
        self.SpMSpV_.set_mask_type(capi.MASK_NONE)
        self.add_module(self.SpMSpV_)
        self.SparseAssign_ = AssignVectorSparseModule(True)
        self.add_module(self.SparseAssign_)
        self.eWiseAdd_ = eWiseAddModule()
        self.add_module(self.eWiseAdd_)

    def get_nnz(self):
        return self.SpMV_.get_nnz()

    def load_and_format_matrix(self, csr_float_npz_path, skip_empty_rows=True):
        """sssp.h:128-143: _preprocess (weights 1, zero diagonal) THEN round dims, CSC for push."""
        m = _load(csr_float_npz_path)
        io.sssp_preprocess(m)
        io.util_round_csr_matrix_dim(m, self.num_channels_ * PACK, self.num_channels_ * PACK)
        self.csr_matrix_ = m
        self.SpMV_.load_and_format_matrix(m, skip_empty_rows)
        self.SpMSpV_.load_and_format_matrix(io.csr2csc(m))
        self.matrix_num_rows_, self.matrix_num_cols_ = m.num_rows, m.num_cols
        assert m.num_rows == m.num_cols

    def send_matrix_host_to_device(self):
        self._matrix_changed()
        self.SpMV_.send_matrix_host_to_device(*self._row_range(self.matrix_num_rows_))
        self._bind_exchange()
        self.SpMSpV_.send_matrix_host_to_device(*self._row_range(self.matrix_num_rows_))

    def _pull_loop(self, first_iter, num_iterations, fused):
        n = self.matrix_num_rows_
        if fused:
            vec, res = self.SpMV_.vector_buf, self.SpMV_.results_buf
            iters = range(first_iter, num_iterations + 1)

            def launches():
                if self.exchange_ is not None:
                    self.SpMV_.iterate_with(vec, None, res, None, len(iters))
                    return
                v, r = vec, res
                for _ in iters:
                    self.SpMV_.run_with(v, None, r)
                    self._exchange(r, n)
                    v, r = r, v

            self._replay(("sssp", len(iters), vec.ptr, res.ptr), launches)
            if len(iters) % 2:
                self.SpMV_.vector_buf, self.SpMV_.results_buf = res, vec
        else:
            assert self.world_ == 1, "the unfused launch sequence is single-GPU"
            self.eWiseAdd_.bind_in_buf(self.SpMV_.results_buf)
            self.eWiseAdd_.bind_out_buf(self.SpMV_.vector_buf)
            for _ in range(first_iter, num_iterations + 1):
                self.SpMV_.run()
                self.eWiseAdd_.run(n, 0.0)

    def pull(self, source, num_iterations, fused=True):
        """sssp.h:152-166"""
        self._begin_run()
        self.SpMV_.set_vector_constant(self.semiring_[2], source, 0.0)   # sssp.h:153-156
        self._pull_loop(1, num_iterations, fused)
        return self.SpMV_.send_vector_device_to_host()

    def _push_setup(self, source):
        self._begin_run()
        self._push_begin()
        self._home_lists()
        self.SpMSpV_.set_vector_single(source, 0.0)
        self.SpMSpV_.set_mask_constant(self.semiring_[2], source, 0.0)   # distance, sssp.h:172-176
        self.SparseAssign_.bind_mask_buf(self.SpMSpV_.results_buf)
        self.SparseAssign_.bind_inout_buf(self.SpMSpV_.mask_buf)
        self.SparseAssign_.bind_new_frontier_buf(self.SpMSpV_.vector_buf)

    def _frontier_lists(self):
        """The two lists the frontier alternates between in the fused push levels (the kernel reads the
        old frontier while it appends the new one, so they cannot be one buffer as in sssp.h:185-187)."""
        cap = self.SpMSpV_.get_num_cols() + 1
        extra = self.__dict__.get("frontier2_buf_")
        if extra is None or not extra.ptr or extra.nbytes != 8 * cap:
            extra = self.frontier2_buf_ = self.ctx.alloc(8 * cap)
        return [self.SpMSpV_.vector_buf, extra]

    def _push_level_fused(self, lists, level, nxt=None):
        """One push level in one launch: SpMSpV on lists[(level - 1) & 1] -> results, with the relax of
        the distance vector and the new frontier -> lists[level & 1] (sssp.h:178-190) fused into the kernel."""
        ep = SpmspvEpilogue(capi.SPMSPV_EP_RELAX, self.SpMSpV_.mask_buf.ptr, 0.0, lists[level & 1].ptr)
        self.SpMSpV_.run_with(lists[(level - 1) & 1], self.SpMSpV_.results_buf, ep, nxt)

    def push(self, source, num_iterations, fused=True):
        """sssp.h:169-194"""
        self._push_setup(source)
        if not fused or self.world_ > 1:
            for _ in range(num_iterations):
                self.SpMSpV_.run()
                self._exchange_frontier(self.SpMSpV_.results_buf, self.semiring_[2])
                self.SparseAssign_.run()
            return self.SpMSpV_.send_mask_device_to_host()
        lists = self._frontier_lists()

        def launches():
            for level in range(1, num_iterations + 1):
                self._push_level_fused(lists, level)

        self._replay(("sssp_push", num_iterations, lists[0].ptr, lists[1].ptr, self.SpMSpV_.results_buf.ptr,
                      self.SpMSpV_.mask_buf.ptr), launches)
        return self.SpMSpV_.send_mask_device_to_host()

    def pull_push(self, source, num_iterations, threshold=0.05, fused=True):
        """sssp.h:197-243.  On one GPU the run is one recorded sequence with the direction decided on the
        device (see BFS.pull_push); the level that stops pushing copies the distance vector into the
        input of the first pull level (sssp.h:219-222)."""
        self.last_num_iterations_ = num_iterations
        if fused and self.world_ == 1 and self.use_graphs_ and num_iterations >= 2:
            return self._pull_push_device(source, num_iterations, threshold)
        n = self.matrix_num_rows_
        self._push_setup(source)
        it = 1
        while True:
            self.SpMSpV_.run()
            self._exchange_frontier(self.SpMSpV_.results_buf, self.semiring_[2])
            self.SparseAssign_.run()
            vector_nnz = self.SpMSpV_.get_results_nnz()
            it += 1
            if not (it < num_iterations and float(vector_nnz) / n < threshold):
                break
        self.push_iterations_host_ = it - 1
        # switch: the distance vector becomes the SpMV input (device copy, no host round trip)
        self.SpMV_.home_buffers()
        if self.SpMV_.vector_buf is None or self.SpMV_.vector_buf.nbytes < 4 * n:
            self.SpMV_.vector_buf = self.ctx.zeros_f32(n)
        capi.d2d(self.ctx, self.SpMV_.vector_buf, self.SpMSpV_.mask_buf, 4 * n)
        self._pull_loop(it, num_iterations, fused)
        return self.SpMV_.send_vector_device_to_host()

    def _pull_push_device(self, source, num_iterations, threshold):
        n, ctx = self.matrix_num_rows_, self.ctx
        self.push_iterations_host_ = None
        self._push_setup(source)
        dist = self.SpMSpV_.mask_buf
        dense = self._dense_pair(self.semiring_[2])
        lists = self._frontier_lists()
        self.SpMSpV_.matrix.reset_levels()

        def launches():
            conds = {level: ctx.cond_create() for level in range(2, num_iterations + 1)}

            def push_level(level):
                nxt = None
                if level < num_iterations:
                    nxt = SpmspvNext(int(level + 1 >= num_iterations), float(threshold), n, conds[level + 1],
                                     capi.SPMSPV_DENSE_COPY, dense[(level + 1) & 1].ptr, dist.ptr, n)
                self._push_level_fused(lists, level, nxt)

            def pull_level(level):
                self.SpMV_.run_with(dense[level & 1], None, dense[(level + 1) & 1])

            push_level(1)
            for level in range(2, num_iterations + 1):
                ctx.branch(conds[level], lambda lv=level: push_level(lv), lambda lv=level: pull_level(lv))

        self._replay(("sssp_pull_push", num_iterations, float(threshold), lists[0].ptr, lists[1].ptr,
                      self.SpMSpV_.results_buf.ptr, dist.ptr, dense[0].ptr, dense[1].ptr), launches)
        # the last level is always a pull level (sssp.h:214: iter < num_iterations): its output is the result
        last = dense[(num_iterations + 1) & 1]
        other = dense[num_iterations & 1]
        self.SpMV_.vector_buf, self.SpMV_.results_buf = last, other
        return self.SpMV_.send_vector_device_to_host()
