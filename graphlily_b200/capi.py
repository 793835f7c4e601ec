"""ctypes binding of ``libgraphlily_b200.so`` (the C ABI in ``include/graphlily_b200.h``).

This is the only route from Python to the CUDA kernels and it has no fallback: if the
shared library is missing the import fails, and without a CUDA device every compute call
raises :class:`GlbError` -- nothing here ever computes on the CPU.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GLB_LIB_PATH") or os.path.join(_HERE, "lib", "libgraphlily_b200.so")   # env: tuning builds only

OP_MUL_ADD, OP_LOGICAL_AND_OR, OP_ADD_MIN = 0, 1, 2
VAL_F32, VAL_U32, VAL_UFIXED = 0, 1, 2
MASK_NONE, MASK_WRITE_TO_ZERO, MASK_WRITE_TO_ONE = 0, 1, 2

IDX_VAL = np.dtype([("index", np.uint32), ("val", np.float32)])

_vp = C.c_void_p
_u32p = C.POINTER(C.c_uint32)
_f32p = C.POINTER(C.c_float)


class GlbError(RuntimeError):
    pass


class Epilogue(C.Structure):
    """glb_spmv_epilogue_t"""
    _fields_ = [("add_enable", C.c_int), ("add_val", C.c_float), ("assign_inout", _vp),
                ("assign_val", C.c_float), ("assign_mask_type", C.c_int)]


class SpmspvEpilogue(C.Structure):
    """glb_spmspv_epilogue_t"""
    _fields_ = [("mode", C.c_int), ("inout", _vp), ("val", C.c_float), ("new_frontier", _vp)]


class SpmspvNext(C.Structure):
    """glb_spmspv_next_t"""
    _fields_ = [("force_stop", C.c_int), ("threshold", C.c_float), ("num_vertices", C.c_uint32), ("cond_next", C.c_uint64),
                ("dense_mode", C.c_int), ("dense", _vp), ("dense_src", _vp), ("dense_len", C.c_uint32)]


SPMSPV_EP_NONE, SPMSPV_EP_ASSIGN, SPMSPV_EP_RELAX = 0, 1, 2
SPMSPV_DENSE_NONE, SPMSPV_DENSE_SCATTER, SPMSPV_DENSE_COPY = 0, 1, 2


class HostLayout(C.Structure):
    """glb_host_layout_t"""
    _fields_ = [("group", C.c_uint32), ("max_groups", C.c_uint32), ("row_cap", C.c_uint32), ("nnz", C.c_uint64),
                ("n_chunks", C.c_uint32), ("n_groups", C.c_uint32), ("n_nz_rows", C.c_uint32), ("n_empty", C.c_uint32),
                ("n_fixups", C.c_uint32), ("tile_k", C.c_uint32), ("n_hot", C.c_uint32),
                ("stream", _u32p), ("flags", _u32p), ("chunk_goff", _u32p), ("chunk_first", _u32p),
                ("nz_rows", _u32p), ("empty_rows", _u32p), ("fixups", _u32p), ("hot_cols", _u32p)]


# name -> (restype, argtypes); must list every symbol include/graphlily_b200.h declares.
SIGNATURES = {
    "glb_version": (C.c_int, []),
    "glb_last_error": (C.c_char_p, []),
    "glb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "glb_ctx_create": (C.c_int, [C.c_int, _vp, C.POINTER(_vp)]),
    "glb_ctx_destroy": (C.c_int, [_vp]),
    "glb_ctx_sync": (C.c_int, [_vp]),
    "glb_device_sync": (C.c_int, [_vp]),
    "glb_ctx_stream": (C.c_int, [_vp, C.POINTER(_vp)]),
    "glb_ctx_kernel_timing": (C.c_int, [_vp, C.c_int]),
    "glb_ctx_kernel_timing_read": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "glb_buffer_alloc": (C.c_int, [_vp, C.c_size_t, C.POINTER(_vp)]),
    "glb_buffer_free": (C.c_int, [_vp, _vp]),
    "glb_buffer_h2d": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "glb_buffer_d2h": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "glb_buffer_h2d_async": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "glb_buffer_d2h_async": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "glb_buffer_d2d": (C.c_int, [_vp, _vp, _vp, C.c_size_t]),
    "glb_buffer_fill_f32": (C.c_int, [_vp, _vp, C.c_float, C.c_size_t]),
    "glb_buffer_fill_one_f32": (C.c_int, [_vp, _vp, C.c_float, C.c_size_t, C.c_size_t, C.c_float]),
    "glb_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_vp)]),
    "glb_host_free": (C.c_int, [_vp]),
    "glb_csr_create": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp, C.c_uint32, C.c_uint32, C.POINTER(_vp)]),
    "glb_csr_destroy": (C.c_int, [_vp]),
    "glb_csr_info": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "glb_csr_format_host": (C.c_int, [C.c_uint32, C.c_uint32, _vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32,
                                      C.POINTER(HostLayout)]),
    "glb_host_layout_free": (C.c_int, [C.POINTER(HostLayout)]),
    "glb_csc_create": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp, C.POINTER(_vp)]),
    "glb_csc_create_rows": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp, C.c_uint32, C.c_uint32, C.POINTER(_vp)]),
    "glb_csc_destroy": (C.c_int, [_vp]),
    "glb_spmv": (C.c_int, [_vp, _vp, C.c_int, C.c_float, C.c_int, _vp, _vp, _vp]),
    "glb_spmv_fused": (C.c_int, [_vp, _vp, C.c_int, C.c_float, C.c_int, _vp, _vp, _vp, C.POINTER(Epilogue)]),
    "glb_spmv_host": (C.c_int, [_vp, _vp, C.c_int, C.c_float, C.c_int, _vp, _vp, _vp]),
    "glb_spmv_host_batch": (C.c_int, [_vp, _vp, C.c_int, C.c_float, C.c_int, C.c_int, _vp, _vp, _vp]),
    "glb_spmspv": (C.c_int, [_vp, _vp, C.c_int, C.c_float, C.c_int, _vp, _vp, _vp]),
    "glb_spmspv_fused": (C.c_int, [_vp, _vp, C.c_int, C.c_float, C.c_int, _vp, _vp, _vp, C.POINTER(SpmspvEpilogue),
                                   C.POINTER(SpmspvNext)]),
    "glb_spmspv_push_state": (C.c_int, [_vp, _vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "glb_spmspv_reset_levels": (C.c_int, [_vp, _vp]),
    "glb_sparse_fill_one": (C.c_int, [_vp, _vp, C.c_uint32, C.c_float]),
    "glb_sparse_count": (C.c_int, [_vp, _vp, C.POINTER(C.c_uint32)]),
    "glb_sparse_to_dense_rows": (C.c_int, [_vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_float]),
    "glb_dense_to_sparse": (C.c_int, [_vp, _vp, C.c_uint32, C.c_float, _vp]),
    "glb_sparse_to_dense": (C.c_int, [_vp, _vp, _vp, C.c_uint32, C.c_float]),
    "glb_ewise_add": (C.c_int, [_vp, _vp, _vp, C.c_uint32, C.c_float]),
    "glb_assign_dense": (C.c_int, [_vp, _vp, _vp, C.c_uint32, C.c_float, C.c_int]),
    "glb_assign_sparse": (C.c_int, [_vp, _vp, _vp, C.c_float]),
    "glb_assign_sparse_relax": (C.c_int, [_vp, _vp, _vp, _vp]),
    "glb_spmv_vt": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_uint32, C.c_int, _vp, _vp, _vp, C.POINTER(Epilogue)]),
    "glb_spmspv_vt": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_uint32, C.c_int, _vp, _vp, _vp]),
    "glb_ewise_add_vt": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_uint32, C.c_uint32]),
    "glb_assign_dense_vt": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_uint32, C.c_uint32, C.c_int]),
    "glb_assign_sparse_vt": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_uint32]),
    "glb_assign_sparse_relax_vt": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp]),
    "glb_nccl_available": (C.c_int, []),
    "glb_nccl_unique_id": (C.c_int, [_vp]),
    "glb_comm_init": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "glb_comm_destroy": (C.c_int, [_vp]),
    "glb_allgather_f32": (C.c_int, [_vp, _vp, C.c_size_t]),
    "glb_graph_begin": (C.c_int, [_vp]),
    "glb_graph_end": (C.c_int, [_vp, C.POINTER(_vp)]),
    "glb_graph_cond_create": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "glb_graph_branch_begin": (C.c_int, [_vp, C.c_uint64]),
    "glb_graph_branch_else": (C.c_int, [_vp]),
    "glb_graph_branch_end": (C.c_int, [_vp]),
    "glb_graph_launch": (C.c_int, [_vp, _vp]),
    "glb_graph_destroy": (C.c_int, [_vp]),
    "glb_xchg_create": (C.c_int, [_vp, C.c_uint32, C.c_int, C.POINTER(_vp)]),
    "glb_xchg_export": (C.c_int, [_vp, _vp]),
    "glb_xchg_connect": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "glb_xchg_block_bytes": (C.c_size_t, [C.c_uint32, C.c_int]),
    "glb_xchg_adopt": (C.c_int, [_vp, C.c_uint32, C.c_int, C.c_int, C.c_int, _vp, _vp, C.POINTER(_vp)]),
    "glb_xchg_has_multicast": (C.c_int, [_vp]),
    "glb_xchg_mc_supported": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "glb_xchg_mc_open": (C.c_int, [_vp, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp), C.POINTER(C.c_int)]),
    "glb_xchg_mc_bind": (C.c_int, [_vp]),
    "glb_xchg_vector": (C.c_int, [_vp, C.c_int, C.POINTER(_vp)]),
    "glb_xchg_allgather": (C.c_int, [_vp, _vp, C.c_int, C.c_size_t, C.c_size_t]),
    "glb_xchg_barrier": (C.c_int, [_vp, _vp]),
    "glb_xchg_status": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "glb_xchg_destroy": (C.c_int, [_vp]),
    "glb_spmv_host_batch_exchange": (C.c_int, [_vp, _vp, C.c_int, C.c_float, C.c_int, _vp, C.c_int, _vp, _vp, _vp]),
    "glb_spmv_exchange": (C.c_int, [_vp, _vp, C.c_int, C.c_float, C.c_int, _vp, C.c_int, C.c_int, _vp, C.POINTER(Epilogue)]),
    "glb_spmv_exchange_iterate": (C.c_int, [_vp, _vp, C.c_int, C.c_float, C.c_int, _vp, C.c_int, C.c_int, _vp,
                                            C.POINTER(Epilogue), C.c_int, C.POINTER(C.c_int)]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C graphlily_b200/csrc` "
            "(or __graft_entry__.build()); graphlily_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc):
    if rc != 0:
        raise GlbError(f"glb error {rc}: {lib.glb_last_error().decode(errors='replace')}")


def device_count():
    n = C.c_int(0)
    rc = lib.glb_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def _ptr(a):
    """Raw address of a numpy array / DeviceBuffer / int / None."""
    if a is None:
        return None
    if isinstance(a, DeviceBuffer):
        return a.ptr
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a)


class Context:
    """One CUDA device + one stream (glb_ctx_t)."""

    def __init__(self, device=0, stream=None):
        h = _vp()
        check(lib.glb_ctx_create(int(device), _vp(stream) if stream else None, C.byref(h)))
        self.handle = h
        self.device = device

    def sync(self):
        check(lib.glb_ctx_sync(self.handle))

    def close(self):
        if self.handle:
            lib.glb_ctx_destroy(self.handle)
            self.handle = None

    def kernel_timing(self, enable):
        check(lib.glb_ctx_kernel_timing(self.handle, int(enable)))

    def kernel_timing_read(self):
        """(ms in SpMV main kernel, ms in fix-up kernel, launches) since the last read."""
        out = (C.c_double * 3)()
        check(lib.glb_ctx_kernel_timing_read(self.handle, out))
        return out[0], out[1], int(out[2])

    # ---- multi-GPU (NCCL resolved with dlopen inside the library) -------------------------
    @staticmethod
    def nccl_unique_id():
        buf = (C.c_char * 128)()
        check(lib.glb_nccl_unique_id(buf))
        return bytes(buf)

    def comm_init(self, unique_id, rank, nranks):
        check(lib.glb_comm_init(self.handle, C.c_char_p(unique_id), rank, nranks))

    def allgather_f32(self, buf, count_per_rank):
        check(lib.glb_allgather_f32(self.handle, _ptr(buf), count_per_rank))

    # ---- launch replay (CUDA graph of a fixed launch sequence) -----------------------------
    def record(self, fn):
        """Record the glb_* launches ``fn()`` makes on this context (nothing executes) -> Graph."""
        check(lib.glb_graph_begin(self.handle))
        try:
            fn()
        except BaseException:
            h = _vp()
            lib.glb_graph_end(self.handle, C.byref(h))
            if h:
                lib.glb_graph_destroy(h)
            raise
        h = _vp()
        check(lib.glb_graph_end(self.handle, C.byref(h)))
        return Graph(self, h)

    def cond_create(self):
        """A device-side condition of the sequence being recorded (glb_graph_cond_create)."""
        c = C.c_uint64(0)
        check(lib.glb_graph_cond_create(self.handle, C.byref(c)))
        return c.value

    def branch(self, cond, if_arm, else_arm):
        """Record ``if_arm()`` / ``else_arm()`` as the two arms of an IF / ELSE node on ``cond``."""
        check(lib.glb_graph_branch_begin(self.handle, cond))
        try:
            if_arm()
            check(lib.glb_graph_branch_else(self.handle))
            else_arm()
        finally:
            check(lib.glb_graph_branch_end(self.handle))

    # ---- buffers ------------------------------------------------------------------
    def alloc(self, nbytes):
        return DeviceBuffer(self, nbytes)

    def to_device(self, array):
        array = np.ascontiguousarray(array)
        buf = DeviceBuffer(self, array.nbytes)
        buf.write(array)
        return buf

    def zeros_f32(self, n, value=0.0):
        buf = DeviceBuffer(self, 4 * n)
        check(lib.glb_buffer_fill_f32(self.handle, buf.ptr, float(value), n))
        return buf


class Graph:
    """A recorded launch sequence (glb_graph_t)."""

    def __init__(self, ctx, handle):
        self.ctx, self.handle = ctx, handle

    def launch(self):
        check(lib.glb_graph_launch(self.ctx.handle, self.handle))

    def __del__(self):
        try:
            if self.handle:
                lib.glb_graph_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class DeviceBuffer:
    """Owning handle over device memory (the role of cl::Buffer)."""

    def __init__(self, ctx, nbytes):
        p = _vp()
        check(lib.glb_buffer_alloc(ctx.handle, int(nbytes), C.byref(p)))
        self.ctx, self.ptr, self.nbytes = ctx, p.value, int(nbytes)

    owned, tag = True, None

    @classmethod
    def view(cls, ctx, ptr, nbytes, tag=None):
        """Non-owning handle over device memory that belongs to something else (an exchange vector)."""
        b = cls.__new__(cls)
        b.ctx, b.ptr, b.nbytes, b.owned, b.tag = ctx, int(ptr), int(nbytes), False, tag
        return b

    def write(self, array):
        array = np.ascontiguousarray(array)
        assert array.nbytes <= self.nbytes
        check(lib.glb_buffer_h2d(self.ctx.handle, self.ptr, array.ctypes.data, array.nbytes))

    def write_at(self, byte_offset, array):
        """Overwrite a slice of the buffer (e.g. the one non-constant element of a start vector)."""
        array = np.ascontiguousarray(array)
        assert byte_offset + array.nbytes <= self.nbytes
        check(lib.glb_buffer_h2d(self.ctx.handle, self.ptr + byte_offset, array.ctypes.data, array.nbytes))

    def read(self, dtype, count, out=None):
        """Blocking copy to the host: a fresh array, or ``out`` (e.g. a PinnedArray's view)."""
        if out is None:
            out = np.empty(count, dtype=dtype)
        assert out.dtype == np.dtype(dtype) and out.size >= count and out.nbytes <= max(self.nbytes, out.nbytes)
        out = out[:count]
        assert out.nbytes <= self.nbytes
        check(lib.glb_buffer_d2h(self.ctx.handle, out.ctypes.data, self.ptr, out.nbytes))
        return out

    def read_sparse(self):
        """Device sparse vector -> (indices, values), head slot stripped."""
        head = self.read(IDX_VAL, 1)
        n = int(head["index"][0])
        body = self.read(IDX_VAL, n + 1)[1:]
        return body["index"].copy(), body["val"].copy()

    def free(self):
        if self.owned and self.ptr and self.ctx.handle:
            lib.glb_buffer_free(self.ctx.handle, self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedArray:
    """Page-locked host array (glb_host_alloc): copies to / from it run at full PCIe speed and are
    not staged.  ``.array`` is a numpy view; the memory is released with the object."""

    def __init__(self, count, dtype=np.float32):
        dtype = np.dtype(dtype)
        p = _vp()
        check(lib.glb_host_alloc(max(1, count * dtype.itemsize), C.byref(p)))
        self.ptr = p.value
        self.array = np.frombuffer((C.c_char * (count * dtype.itemsize)).from_address(self.ptr), dtype=dtype, count=count)

    def __del__(self):
        try:
            if self.ptr:
                self.array = None
                lib.glb_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def sparse_to_numpy(idx, val, capacity=None):
    """(indices, values) -> idx_val_t array with the {nnz, 0} head slot."""
    idx = np.asarray(idx, np.uint32)
    n = len(idx)
    out = np.zeros(max(n + 1, capacity or 0), dtype=IDX_VAL)
    out["index"][0] = n
    out["index"][1:n + 1] = idx
    out["val"][1:n + 1] = np.asarray(val, np.float32)
    return out


class CsrMatrix:
    """Device-resident CSR row shard in lane-segment layout (glb_csr_t)."""

    def __init__(self, ctx, m, row_begin=0, row_end=None):
        row_end = int(m.num_rows) if row_end is None else row_end
        self._keep = (np.ascontiguousarray(m.indptr, np.uint32), np.ascontiguousarray(m.indices, np.uint32),
                      np.ascontiguousarray(m.data, np.float32))
        h = _vp()
        check(lib.glb_csr_create(ctx.handle, int(m.num_rows), int(m.num_cols), self._keep[0].ctypes.data,
                                 self._keep[1].ctypes.data, self._keep[2].ctypes.data, row_begin, row_end, C.byref(h)))
        self._keep = None  # the layout lives on the device now
        self.ctx, self.handle = ctx, h
        self.num_rows, self.num_cols = int(m.num_rows), int(m.num_cols)
        self.row_begin, self.row_end = row_begin, row_end

    def info(self):
        a = (C.c_uint64 * 8)()
        check(lib.glb_csr_info(self.handle, a))
        keys = ("rows", "cols", "nnz", "chunks", "fixups", "empty_rows", "device_bytes", "tile_k")
        return dict(zip(keys, (int(v) for v in a)))

    def spmv(self, op, zero, mask_type, x, mask, y, epilogue=None):
        if epilogue is None:
            check(lib.glb_spmv(self.ctx.handle, self.handle, op, zero, mask_type, _ptr(x), _ptr(mask), _ptr(y)))
        else:
            check(lib.glb_spmv_fused(self.ctx.handle, self.handle, op, zero, mask_type, _ptr(x), _ptr(mask), _ptr(y),
                                     C.byref(epilogue)))

    def spmv_host(self, op, zero, mask_type, x_host, mask_host, y_host):
        check(lib.glb_spmv_host(self.ctx.handle, self.handle, op, zero, mask_type, _ptr(x_host), _ptr(mask_host),
                                _ptr(y_host)))

    def spmv_host_batch(self, op, zero, mask_type, x_hosts, mask_hosts, y_hosts):
        """Pipelined glb_spmv_host over a sequence of host vectors (addresses or objects _ptr accepts)."""
        n = len(x_hosts)
        assert len(y_hosts) == n and (mask_hosts is None or len(mask_hosts) == n)
        arr = C.c_void_p * n
        xs = arr(*[_ptr(v) for v in x_hosts])
        ys = arr(*[_ptr(v) for v in y_hosts])
        ms = arr(*[_ptr(v) for v in mask_hosts]) if mask_hosts is not None else None
        check(lib.glb_spmv_host_batch(self.ctx.handle, self.handle, op, zero, mask_type, n, xs, ms, ys))

    def close(self):
        if self.handle:
            lib.glb_csr_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Exchange:
    """Peer-mapped vectors of a row-sharded run (glb_xchg_t): every rank's SpMV stores its rows
    into all ranks' copies, so no allgather follows the kernel.  ``all_gather_bytes(b)`` must
    return the list of every rank's ``b`` in rank order (e.g. torch.distributed.all_gather_object)."""

    def __init__(self, ctx, n_floats, rank, nranks, all_gather_bytes, n_vectors=2):
        self.ctx, self.handle, self.n, self.rank, self.nranks = ctx, None, int(n_floats), rank, nranks
        self.n_vectors = n_vectors
        # every rank reaches the handle exchange (a collective) whatever happened locally
        mine, err = b"", None
        try:
            h = _vp()
            check(lib.glb_xchg_create(ctx.handle, int(n_floats), n_vectors, C.byref(h)))
            self.handle = h
            buf = (C.c_char * 64)()
            check(lib.glb_xchg_export(h, buf))
            mine = bytes(buf)
        except GlbError as e:
            err = e
        handles = all_gather_bytes(mine)
        if err is not None or any(len(b) != 64 for b in handles):
            self.close()
            raise err or GlbError("peer exchange: another rank could not export its block")
        check(lib.glb_xchg_connect(self.handle, rank, nranks, C.c_char_p(b"".join(handles))))

    @classmethod
    def adopt(cls, ctx, n_floats, rank, nranks, block_ptrs, multicast_ptr=None, n_vectors=2, keep=None):
        """Exchange over blocks the host mapped itself (torch symmetric memory: ``hdl.buffer_ptrs``,
        ``hdl.multicast_ptr``); ``keep`` = objects that own the mapping."""
        self = cls.__new__(cls)
        self.ctx, self.handle, self.n, self.rank, self.nranks = ctx, None, int(n_floats), rank, nranks
        self.n_vectors, self._keep = n_vectors, keep
        arr = (C.c_void_p * nranks)(*[int(p) for p in block_ptrs])
        h = _vp()
        check(lib.glb_xchg_adopt(ctx.handle, int(n_floats), n_vectors, rank, nranks, arr,
                                 _vp(int(multicast_ptr)) if multicast_ptr else None, C.byref(h)))
        self.handle = h
        return self

    @classmethod
    def open_multicast(cls, ctx, n_floats, rank, nranks, share_fd, agree, n_vectors=2):
        """Exchange over an NVSwitch multicast object the LIBRARY creates (glb_xchg_mc_open / _bind; no framework maps
        anything).  Collective.  ``share_fd(fd)``: rank 0 passes the multicast object's file descriptor, every rank gets
        back a descriptor valid in its own process (rank 0: the same one); ``agree(ok) -> bool``: logical AND over the
        ranks (doubles as the barrier between the steps).  Returns None (on every rank) when any step failed anywhere."""
        self = cls.__new__(cls)
        self.ctx, self.handle, self.n, self.rank, self.nranks = ctx, None, int(n_floats), rank, nranks
        self.n_vectors, self._keep, self.error = n_vectors, None, None
        sup = C.c_int(0)
        if not agree(lib.glb_xchg_mc_supported(ctx.handle, C.byref(sup)) == 0 and sup.value != 0):
            return None
        h, fd = _vp(), C.c_int(-1)

        def attempt(call):
            try:
                check(call())
                return True
            except GlbError as e:
                self.error = e
                return False
        ok = True
        if rank == 0:
            ok = attempt(lambda: lib.glb_xchg_mc_open(ctx.handle, int(n_floats), n_vectors, 0, nranks, -1, C.byref(h), C.byref(fd)))
        if not agree(ok):
            return None
        got = share_fd(fd.value if rank == 0 else None)
        if rank != 0:
            ok = got is not None and got >= 0 and attempt(
                lambda: lib.glb_xchg_mc_open(ctx.handle, int(n_floats), n_vectors, rank, nranks, got, C.byref(h), None))
        if got is not None and got >= 0:
            os.close(got)   # the driver holds its own reference to the object
        if ok:
            self.handle = h
        if not agree(ok):   # (also the barrier: every rank has added its device)
            self.close()
            return None
        ok = attempt(lambda: lib.glb_xchg_mc_bind(self.handle))
        if not agree(ok):
            self.close()
            return None
        return self

    @staticmethod
    def block_bytes(n_floats, n_vectors):
        return int(lib.glb_xchg_block_bytes(int(n_floats), n_vectors))

    def has_multicast(self):
        return bool(lib.glb_xchg_has_multicast(self.handle))

    def vector(self, which):
        p = _vp()
        check(lib.glb_xchg_vector(self.handle, which, C.byref(p)))
        return p.value

    def spmv(self, matrix, op, zero, mask_type, src_vec, dst_vec, mask=None, epilogue=None):
        check(lib.glb_spmv_exchange(self.ctx.handle, matrix.handle, op, zero, mask_type, self.handle, src_vec, dst_vec,
                                    _ptr(mask), C.byref(epilogue) if epilogue is not None else None))

    def spmv_iterate(self, matrix, op, zero, mask_type, src_vec, dst_vec, n_steps, mask=None, epilogues=None, plan=None):
        """glb_spmv_exchange_iterate: ``n_steps`` ping-pong iterations src -> dst -> src ...;
        ``epilogues``: None or one Epilogue per step; ``plan``: None or [(read, written)] per step."""
        eps = None
        if epilogues is not None:
            assert len(epilogues) == n_steps
            eps = (Epilogue * n_steps)(*epilogues)
        vp = None
        if plan is not None:
            assert len(plan) == n_steps
            vp = (C.c_int * (2 * n_steps))(*[v for pair in plan for v in pair])
        check(lib.glb_spmv_exchange_iterate(self.ctx.handle, matrix.handle, op, zero, mask_type, self.handle, src_vec, dst_vec,
                                            _ptr(mask), eps, n_steps, vp))

    def spmv_host_batch(self, matrix, op, zero, mask_type, x_hosts, mask_hosts, y_hosts):
        """glb_spmv_host_batch_exchange: every rank uploads its slice of each x, NVLink completes it."""
        n = len(x_hosts)
        arr = C.c_void_p * n
        xs, ys = arr(*[_ptr(v) for v in x_hosts]), arr(*[_ptr(v) for v in y_hosts])
        ms = arr(*[_ptr(v) for v in mask_hosts]) if mask_hosts is not None else None
        check(lib.glb_spmv_host_batch_exchange(self.ctx.handle, matrix.handle, op, zero, mask_type, self.handle, n, xs, ms, ys))

    def buffer(self, which):
        """Vector ``which`` as a (non-owning) DeviceBuffer."""
        return DeviceBuffer.view(self.ctx, self.vector(which), 4 * self.n, tag=("xchg", which))

    def allgather(self, which, offset, count):
        check(lib.glb_xchg_allgather(self.ctx.handle, self.handle, which, offset, count))

    def barrier(self):
        check(lib.glb_xchg_barrier(self.ctx.handle, self.handle))

    def timed_out(self):
        t = C.c_int(0)
        check(lib.glb_xchg_status(self.handle, C.byref(t)))
        return bool(t.value)

    def close(self):
        if self.handle:
            lib.glb_xchg_destroy(self.handle)
            self.handle = None


class CscMatrix:
    """Device-resident CSC (glb_csc_t); ``m.indptr`` runs over columns, ``m.indices`` are row ids."""

    def __init__(self, ctx, m, row_begin=0, row_end=None):
        ip, ix, d = (np.ascontiguousarray(m.indptr, np.uint32), np.ascontiguousarray(m.indices, np.uint32),
                     np.ascontiguousarray(m.data, np.float32))
        row_end = int(m.num_rows) if row_end is None else row_end
        h = _vp()
        check(lib.glb_csc_create_rows(ctx.handle, int(m.num_rows), int(m.num_cols), ip.ctypes.data, ix.ctypes.data,
                                      d.ctypes.data, row_begin, row_end, C.byref(h)))
        self.ctx, self.handle = ctx, h
        self.num_rows, self.num_cols = int(m.num_rows), int(m.num_cols)

    def spmspv(self, op, zero, mask_type, x, mask, y, epilogue=None, nxt=None):
        """glb_spmspv_fused: ``epilogue`` = SpmspvEpilogue or None, ``nxt`` = SpmspvNext or None."""
        check(lib.glb_spmspv_fused(self.ctx.handle, self.handle, op, zero, mask_type, _ptr(x), _ptr(mask), _ptr(y),
                                   C.byref(epilogue) if epilogue is not None else None,
                                   C.byref(nxt) if nxt is not None else None))

    def push_state(self):
        """(keep_pushing, push_levels) of the device-side direction decision; blocking."""
        k, n = C.c_uint32(0), C.c_uint32(0)
        check(lib.glb_spmspv_push_state(self.ctx.handle, self.handle, C.byref(k), C.byref(n)))
        return bool(k.value), int(n.value)

    def reset_levels(self):
        check(lib.glb_spmspv_reset_levels(self.ctx.handle, self.handle))

    def close(self):
        if self.handle:
            lib.glb_csc_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- apply operators ----------------------------------------------------------------
def ewise_add(ctx, src, dst, length, val):
    check(lib.glb_ewise_add(ctx.handle, _ptr(src), _ptr(dst), length, val))


def assign_dense(ctx, mask, inout, length, val, mask_type):
    check(lib.glb_assign_dense(ctx.handle, _ptr(mask), _ptr(inout), length, val, mask_type))


def assign_sparse(ctx, sparse_list, inout, val):
    check(lib.glb_assign_sparse(ctx.handle, _ptr(sparse_list), _ptr(inout), val))


def assign_sparse_relax(ctx, sparse_list, inout, new_frontier):
    check(lib.glb_assign_sparse_relax(ctx.handle, _ptr(sparse_list), _ptr(inout), _ptr(new_frontier)))


def sparse_count(ctx, sparse_list):
    n = C.c_uint32(0)
    check(lib.glb_sparse_count(ctx.handle, _ptr(sparse_list), C.byref(n)))
    return n.value


def sparse_to_dense(ctx, sparse_list, dense, length, zero):
    check(lib.glb_sparse_to_dense(ctx.handle, _ptr(sparse_list), _ptr(dense), length, zero))


def sparse_to_dense_rows(ctx, sparse_list, dense, row_begin, row_end, zero):
    check(lib.glb_sparse_to_dense_rows(ctx.handle, _ptr(sparse_list), _ptr(dense), row_begin, row_end, zero))


def dense_to_sparse(ctx, dense, length, zero, sparse_list):
    check(lib.glb_dense_to_sparse(ctx.handle, _ptr(dense), length, zero, _ptr(sparse_list)))


def d2d(ctx, dst, src, nbytes):
    check(lib.glb_buffer_d2d(ctx.handle, _ptr(dst), _ptr(src), nbytes))


def format_host(m, row_begin=0, row_end=None, tile_k=0, with_data=True):
    """Host-only view of the lane-segment layout (no GPU needed); returns a dict of numpy arrays."""
    row_end = int(m.num_rows) if row_end is None else row_end
    ip, ix = np.ascontiguousarray(m.indptr, np.uint32), np.ascontiguousarray(m.indices, np.uint32)
    d = np.ascontiguousarray(m.data, np.float32) if with_data else None
    L = HostLayout()
    check(lib.glb_csr_format_host(int(m.num_rows), int(m.num_cols), ip.ctypes.data, ix.ctypes.data,
                                  d.ctypes.data if d is not None else None, row_begin, row_end, int(tile_k), C.byref(L)))

    def arr(p, n):
        return np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, np.uint32)

    out = dict(group=L.group, max_groups=L.max_groups, row_cap=L.row_cap, nnz=int(L.nnz), n_chunks=L.n_chunks,
               n_groups=L.n_groups, tile_k=L.tile_k, stream=arr(L.stream, 256 * L.n_groups),
               flags=arr(L.flags, 32 * L.n_chunks).reshape(-1, 32), chunk_goff=arr(L.chunk_goff, L.n_chunks + 1),
               chunk_first=arr(L.chunk_first, L.n_chunks), nz_rows=arr(L.nz_rows, L.n_nz_rows),
               empty_rows=arr(L.empty_rows, L.n_empty), fixups=arr(L.fixups, 3 * L.n_fixups).reshape(-1, 3),
               hot_cols=arr(L.hot_cols, L.n_hot))
    lib.glb_host_layout_free(C.byref(L))
    return out
