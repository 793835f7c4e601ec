"""Operator modules: the Python mirror of ``graphlily::module`` over the C ABI.

Same class names, method names, argument meaning and error behaviour as
``/root/reference/graphlily/module/*.h`` (the C++ mirror with identical signatures is in
``include/graphlily/module``).  ``cl::Buffer`` members become :class:`capi.DeviceBuffer`
handles with the same public names (``vector_buf``, ``mask_buf``, ``results_buf`` ...), so the
apps alias them across modules exactly as the reference does (``bfs.h:113-116``).
``run()`` enqueues on the context's stream; every ``send_*_device_to_host`` synchronises.
Tuning arguments of the FPGA build (``num_channels``, ``*_buf_len``) are accepted and ignored.
"""
import numpy as np

from . import capi
from .io import CSRMatrix


class BaseModule:
    """base_module.h:10-103"""
    _shared_ctx = None

    def __init__(self, kernel_name="overlay"):
        self.kernel_name_ = kernel_name
        self.target_ = "hw"
        self.ctx = None

    def get_kernel_name(self):
        return self.kernel_name_

    def set_target(self, target):
        assert target in ("sw_emu", "hw_emu", "hw")   # base_module.h:75
        self.target_ = target

    def set_context(self, ctx):
        self.ctx = ctx

    def set_up_runtime(self, xclbin_file_path=None, device=0):
        """base_module.h:106-133; the bitstream path is ignored (kernels are in the library)."""
        if self.ctx is None:
            self.ctx = capi.Context(device)

    def copy_buffer_device_to_device(self, src, dst, nbytes):
        """base_module.h:82-85 (enqueueCopyBuffer + finish)"""
        capi.d2d(self.ctx, dst, src, nbytes)
        self.ctx.sync()

    # send_*_device_to_host returns a copy of the host mirror, as the reference does.  With
    # pinned_results = True the dense results land in a page-locked mirror owned by the module and
    # that array itself is returned (valid until the module's next read-back): no staging copy and
    # no first-touch page faults on a fresh 12 MB array per call.
    pinned_results = False

    def _read_dense(self, buf, n):
        if not self.pinned_results:
            return buf.read(np.float32, n)
        mirror = self.__dict__.get("host_mirror_")
        if mirror is None or mirror.array.size != n:
            mirror = self.host_mirror_ = capi.PinnedArray(n, np.float32)
        return buf.read(np.float32, n, out=mirror.array)

    def _dense_to_device(self, vec):
        return self.ctx.to_device(np.ascontiguousarray(vec, np.float32))

    def _constant_on_device(self, n, value, index=None, index_value=None, reuse=None):
        """A start vector that is constant but for one element, built on the device (a fill kernel
        and a 4-byte copy) instead of uploading n floats.  ``reuse``: an existing buffer of the
        right size is refilled in place, so the app loops see the same addresses call after call
        (their recorded launch sequences stay valid)."""
        buf = reuse if (reuse is not None and reuse.ptr and reuse.nbytes == 4 * n) else self.ctx.alloc(4 * n)
        if index is None:
            capi.check(capi.lib.glb_buffer_fill_f32(self.ctx.handle, buf.ptr, float(value), n))
        else:
            capi.check(capi.lib.glb_buffer_fill_one_f32(self.ctx.handle, buf.ptr, float(value), n, int(index), float(index_value)))
        return buf


class SpMVModule(BaseModule):
    """spmv_module.h:26-272"""

    def __init__(self, num_channels=16, out_buf_len=0, vec_buf_len=0):
        super().__init__()
        self.semiring_ = (capi.OP_MUL_ADD, 1.0, 0.0)
        self.mask_type_ = capi.MASK_NONE
        self.csr_matrix_float_ = None
        self.matrix = None
        self.vector_buf = self.mask_buf = self.results_buf = None
        self.exchange = None   # capi.Exchange of a row-sharded run: vector / results are its vectors

    def set_semiring(self, semiring):
        self.semiring_ = semiring

    def set_mask_type(self, mask_type):
        self.mask_type_ = mask_type

    def get_num_rows(self):
        return self.csr_matrix_float_.num_rows

    def get_num_cols(self):
        return self.csr_matrix_float_.num_cols

    def get_nnz(self):
        return self.csr_matrix_float_.nnz

    def load_and_format_matrix(self, csr_matrix_float, skip_empty_rows=True):
        """spmv_module.h:282-370; the device layout itself is built at upload time."""
        self.csr_matrix_float_ = csr_matrix_float

    def send_matrix_host_to_device(self, row_begin=0, row_end=None):
        """spmv_module.h:374-420 (+ allocation of the results buffer)."""
        self.matrix = capi.CsrMatrix(self.ctx, self.csr_matrix_float_, row_begin, row_end)
        self.results_buf = self.ctx.zeros_f32(self.get_num_rows())

    def send_vector_host_to_device(self, vector):
        self.vector_buf = self._dense_to_device(vector)      # spmv_module.h:424-440

    def send_mask_host_to_device(self, mask):
        self.mask_buf = self._dense_to_device(mask)          # spmv_module.h:444-459

    def home_buffers(self):
        """The app loops ping-pong vector / results and end with the roles swapped after an odd number
        of iterations; a new run starts from the same assignment every time (lower address = vector),
        so the recorded launch sequence of the previous run with these arguments is found again."""
        v, r = self.vector_buf, self.results_buf
        if v is not None and r is not None and v.ptr and r.ptr and v.nbytes == r.nbytes and v.ptr > r.ptr:
            self.vector_buf, self.results_buf = r, v

    def set_vector_constant(self, value, index=None, index_value=None):
        n = self.get_num_cols()
        if self.vector_buf is None or not self.vector_buf.ptr or self.vector_buf.nbytes != 4 * n:
            self.vector_buf = self.ctx.alloc(4 * n)   # allocate first, then settle the roles: the first run is canonical too
        self.home_buffers()
        self.vector_buf = self._constant_on_device(n, value, index, index_value, self.vector_buf)

    def set_mask_constant(self, value, index=None, index_value=None):
        self.mask_buf = self._constant_on_device(self.get_num_rows(), value, index, index_value, self.mask_buf)

    def bind_mask_buf(self, src_buf):
        self.mask_buf = src_buf                              # spmv_module.h:463-467

    def run(self, epilogue=None):
        """spmv_module.h:471-475"""
        self.run_with(self.vector_buf, self.mask_buf, self.results_buf, epilogue)

    def run_with(self, vector_buf, mask_buf, results_buf, epilogue=None):
        """run() on explicit buffers (the app loops ping-pong vector / results instead of copying)."""
        op, _one, zero = self.semiring_
        mask = mask_buf if self.mask_type_ != capi.MASK_NONE else None
        if self.exchange is not None:
            # row-sharded with peer-mapped vectors: the write-back stores the rows on every rank
            assert vector_buf.tag and results_buf.tag, "vector / results must be exchange vectors"
            self.exchange.spmv(self.matrix, op, zero, self.mask_type_, vector_buf.tag[1], results_buf.tag[1], mask, epilogue)
        else:
            self.matrix.spmv(op, zero, self.mask_type_, vector_buf, mask, results_buf, epilogue)

    def iterate_with(self, vector_buf, mask_buf, results_buf, epilogues=None, n_steps=None):
        """Row-sharded pull loop in one call (glb_spmv_exchange_iterate): vector -> results -> vector ..."""
        assert self.exchange is not None and vector_buf.tag and results_buf.tag
        op, _one, zero = self.semiring_
        mask = mask_buf if self.mask_type_ != capi.MASK_NONE else None
        n_steps = len(epilogues) if n_steps is None else n_steps
        self.exchange.spmv_iterate(self.matrix, op, zero, self.mask_type_, vector_buf.tag[1], results_buf.tag[1], n_steps,
                                   mask, epilogues)

    def send_vector_device_to_host(self):
        return self._read_dense(self.vector_buf, self.get_num_cols())

    def send_mask_device_to_host(self):
        return self._read_dense(self.mask_buf, self.get_num_rows())

    def send_results_device_to_host(self):
        return self._read_dense(self.results_buf, self.get_num_rows())


class SpMSpVModule(BaseModule):
    """spmspv_module.h:26-254"""

    def __init__(self, out_buf_len=0):
        super().__init__()
        self.semiring_ = (capi.OP_MUL_ADD, 1.0, 0.0)
        self.mask_type_ = capi.MASK_NONE
        self.csc_matrix_float_ = None
        self.matrix = None
        self.vector_buf = self.mask_buf = self.results_buf = None

    def set_semiring(self, semiring):
        self.semiring_ = semiring

    def set_mask_type(self, mask_type):
        self.mask_type_ = mask_type

    def get_num_rows(self):
        return self.csc_matrix_float_.num_rows

    def get_num_cols(self):
        return self.csc_matrix_float_.num_cols

    def get_nnz(self):
        return self.csc_matrix_float_.nnz

    def load_and_format_matrix(self, csc_matrix_float):
        self.csc_matrix_float_ = csc_matrix_float             # spmspv_module.h:264-286

    def send_matrix_host_to_device(self, row_begin=0, row_end=None):
        """spmspv_module.h:290-370: matrix upload + results (rows+1) and vector (cols+1) lists.
        A row range uploads the row shard of a multi-GPU run (the entries whose row lies inside)."""
        self.matrix = capi.CscMatrix(self.ctx, self.csc_matrix_float_, row_begin, row_end)
        self.results_buf = self.ctx.to_device(np.zeros(self.get_num_rows() + 1, capi.IDX_VAL))
        self.vector_buf = self.ctx.to_device(np.zeros(self.get_num_cols() + 1, capi.IDX_VAL))

    def send_vector_host_to_device(self, vector):
        """``vector``: idx_val_t array with the {nnz, -} head (spmspv_module.h:374-399)."""
        self.vector_buf.write(np.ascontiguousarray(vector, capi.IDX_VAL))

    def set_vector_single(self, index, val):
        """The one-entry start frontier, written on the device (no blocking upload)."""
        capi.check(capi.lib.glb_sparse_fill_one(self.ctx.handle, self.vector_buf.ptr, int(index), float(val)))

    def send_mask_host_to_device(self, mask):
        self.mask_buf = self._dense_to_device(mask)          # spmspv_module.h:403-433

    def set_mask_constant(self, value, index=None, index_value=None):
        self.mask_buf = self._constant_on_device(self.get_num_rows(), value, index, index_value, self.mask_buf)

    def bind_mask_buf(self, src_buf):
        self.mask_buf = src_buf

    def run(self):
        """spmspv_module.h:437-441"""
        self.run_with(self.vector_buf, self.results_buf)

    def run_with(self, vector_buf, results_buf, epilogue=None, nxt=None):
        """run() on explicit list buffers, optionally with the fused sparse assign of a push level and the
        device-side direction decision (glb_spmspv_fused)."""
        op, _one, zero = self.semiring_
        self.matrix.spmspv(op, zero, self.mask_type_, vector_buf,
                           self.mask_buf if self.mask_type_ != capi.MASK_NONE else None, results_buf, epilogue, nxt)

    def send_vector_device_to_host(self):
        n = capi.sparse_count(self.ctx, self.vector_buf)
        return self.vector_buf.read(capi.IDX_VAL, n + 1)

    def send_mask_device_to_host(self):
        return self._read_dense(self.mask_buf, self.get_num_rows())

    def send_results_device_to_host(self):
        n = capi.sparse_count(self.ctx, self.results_buf)
        return self.results_buf.read(capi.IDX_VAL, n + 1)

    def get_results_nnz(self):
        return capi.sparse_count(self.ctx, self.results_buf)  # spmspv_module.h:239-242


class eWiseAddModule(BaseModule):
    """add_scalar_vector_dense_module.h:17-138"""

    def __init__(self):
        super().__init__()
        self.in_buf = self.out_buf = None
        self._out_len = 0

    def send_in_host_to_device(self, vec):
        self.in_buf = self._dense_to_device(vec)
        self._in_len = len(vec)

    def allocate_out_buf(self, length):
        self.out_buf = self.ctx.zeros_f32(length)
        self._out_len = length

    def bind_in_buf(self, src_buf):
        self.in_buf = src_buf

    def bind_out_buf(self, src_buf):
        self.out_buf = src_buf

    def run(self, length, val):
        capi.ewise_add(self.ctx, self.in_buf, self.out_buf, length, val)

    def send_out_device_to_host(self):
        return self.out_buf.read(np.float32, self.out_buf.nbytes // 4)


class AssignVectorDenseModule(BaseModule):
    """assign_vector_dense_module.h:17-164"""

    def __init__(self):
        super().__init__()
        self.mask_type_ = None
        self.mask_buf = self.inout_buf = None

    def set_mask_type(self, mask_type):
        if mask_type == capi.MASK_NONE:   # assign_vector_dense_module.h:88-95: print + exit(EXIT_FAILURE)
            print("Please set the mask type")
            raise SystemExit(1)
        self.mask_type_ = mask_type

    def send_mask_host_to_device(self, mask):
        self.mask_buf = self._dense_to_device(mask)

    def send_inout_host_to_device(self, inout):
        self.inout_buf = self._dense_to_device(inout)

    def bind_mask_buf(self, src_buf):
        self.mask_buf = src_buf

    def bind_inout_buf(self, src_buf):
        self.inout_buf = src_buf

    def run(self, length, val):
        capi.assign_dense(self.ctx, self.mask_buf, self.inout_buf, length, val, self.mask_type_)

    def send_mask_device_to_host(self):
        return self.mask_buf.read(np.float32, self.mask_buf.nbytes // 4)

    def send_inout_device_to_host(self):
        return self.inout_buf.read(np.float32, self.inout_buf.nbytes // 4)


class AssignVectorSparseModule(BaseModule):
    """assign_vector_sparse_module.h:17-210"""

    def __init__(self, generate_new_frontier):
        super().__init__()
        self.generate_new_frontier_ = generate_new_frontier
        self.mask_buf = self.inout_buf = self.new_frontier_buf = None

    def send_mask_host_to_device(self, mask):
        mask = np.ascontiguousarray(mask, capi.IDX_VAL)
        self.mask_buf = self.ctx.to_device(mask)
        if self.generate_new_frontier_:   # assign_vector_sparse_module.h:232-247
            self.new_frontier_buf = self.ctx.to_device(np.zeros(len(mask), capi.IDX_VAL))

    def send_inout_host_to_device(self, inout):
        self.inout_buf = self._dense_to_device(inout)

    def bind_mask_buf(self, src_buf):
        self.mask_buf = src_buf

    def bind_inout_buf(self, src_buf):
        self.inout_buf = src_buf

    def bind_new_frontier_buf(self, src_buf):
        if not self.generate_new_frontier_:
            print("[ERROR]: this->generate_new_frontier_ should be true")
            raise SystemExit(1)
        self.new_frontier_buf = src_buf

    def run(self, val=None):
        if val is not None:
            if self.generate_new_frontier_:
                print("[ERROR]: this->generate_new_frontier_ should be false")
                raise SystemExit(1)
            capi.assign_sparse(self.ctx, self.mask_buf, self.inout_buf, val)
        else:
            if not self.generate_new_frontier_:
                print("[ERROR]: this->generate_new_frontier_ should be true")
                raise SystemExit(1)
            capi.assign_sparse_relax(self.ctx, self.mask_buf, self.inout_buf, self.new_frontier_buf)

    def send_mask_device_to_host(self):
        n = capi.sparse_count(self.ctx, self.mask_buf)
        return self.mask_buf.read(capi.IDX_VAL, n + 1)

    def send_inout_device_to_host(self):
        return self.inout_buf.read(np.float32, self.inout_buf.nbytes // 4)

    def send_new_frontier_device_to_host(self):
        if not self.generate_new_frontier_:
            print("[ERROR]: this->generate_new_frontier_ should be true")
            raise SystemExit(1)
        n = capi.sparse_count(self.ctx, self.new_frontier_buf)
        return self.new_frontier_buf.read(capi.IDX_VAL, n + 1)
