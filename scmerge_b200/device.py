"""Device-side plumbing above the C ABI: context, device buffers, cell x gene matrices, pinned host arrays.

No arithmetic happens here; every numeric result comes out of libscmerge_b200's kernels.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib as L

_DT = {np.dtype(np.float64): 8, np.dtype(np.int32): 4, np.dtype(np.int64): 8, np.dtype(np.uint8): 1}


class Context:
    """Owns a `scm_ctx` (one GPU, one stream).  `stream` may be a raw cudaStream_t (int), e.g.
    `torch.cuda.current_stream().cuda_stream`, so that torch events and NCCL order with our kernels."""

    def __init__(self, device: int = 0, stream: int | None = None, _borrowed=None, background: bool = False):
        self._h = C.c_void_p()
        self.device = int(device)
        self.stream = stream
        self._cb = None  # keep the ctypes callback alive
        self.world, self.rank, self._ar_fn = 1, None, None
        if _borrowed is not None:  # a context owned by a MultiContext (scm_multi_ctx)
            self._h = C.c_void_p(_borrowed)
            self._fin = None
            r, w = C.c_int(0), C.c_int(1)
            L.check(L.lib().scm_ctx_comm_info(self._h, C.byref(r), C.byref(w), None))
            self.rank, self.world = r.value, w.value
            return
        # stream=None: the library creates its own non-blocking stream.  stream=0 is torch's default stream, i.e. the
        # legacy NULL stream, which must be passed as the cudaStreamLegacy handle (1) because NULL means "own stream".
        # background=True: an own stream of the least priority (handle 2): work that should yield the SMs to another context
        handle = (2 if background else 0) if stream is None else (1 if int(stream) == 0 else int(stream))
        L.check(L.lib().scm_ctx_create(int(device), C.c_void_p(handle), C.byref(self._h)))
        self._fin = weakref.finalize(self, L.lib().scm_ctx_destroy, self._h)

    @property
    def handle(self):
        return self._h

    def sync(self):
        L.check(L.lib().scm_ctx_sync(self._h))

    def release_scratch(self):
        """Hand the library's cached scratch (e.g. the m x n pre-shifted copy a fit leaves in the pool) back to the driver."""
        L.check(L.lib().scm_ctx_release_scratch(self._h))

    def launch_count(self) -> int:
        n = C.c_int64()
        L.check(L.lib().scm_ctx_launch_count(self._h, C.byref(n)))
        return n.value

    def timers(self) -> dict:
        ms = (C.c_double * 10)()
        L.check(L.lib().scm_timers(self._h, ms, 10))
        names = ["seg_colsum", "gram", "allreduce", "eig", "coef", "adjust", "standardise", "gram_syrk_kernel", "ingest", "preshift_copy"]
        return dict(zip(names, list(ms)))

    def set_allreduce(self, fn, world_size: int, rank: int | None = None):
        """fn(dev_ptr:int, count:int, dtype:int(0 f64, 1 i64), stream:int) -> None; sums in place across ranks.
        `rank` (this process's index among the `world_size` ranks sharing the cells) is only needed by the host-side
        gathers of the ruvK selection (scRUVIII with several k)."""
        self.world, self.rank, self._ar_fn = int(world_size), rank, fn
        def _cb(ptr, count, dtype, stream, _user):
            try:
                fn(ptr, count, dtype, stream)
                return 0
            except Exception as exc:  # the error text travels through scm_last_error's generic message
                import traceback
                traceback.print_exc()
                del exc
                return 1
        self._cb = L.ALLREDUCE_FN(_cb)
        L.check(L.lib().scm_ctx_set_allreduce(self._h, self._cb, None, int(world_size)))

    def init_nccl(self, rank: int, world_size: int, bcast):
        """Built-in NCCL communicator, one rank per context (scm_ctx_comm_init_nccl; collective).  `bcast(b: bytes | None) ->
        bytes` must return rank 0's argument on every rank (any transport: torch.distributed, MPI, a shared file):
        it carries the 128-byte NCCL unique id.  After this the library's collectives never leave C."""
        uid = C.create_string_buffer(128)
        if rank == 0 and world_size > 1:
            L.check(L.lib().scm_nccl_unique_id(uid))
        if world_size > 1:
            got = bcast(bytes(uid.raw) if rank == 0 else None)
            uid = C.create_string_buffer(bytes(got), 128)
        L.check(L.lib().scm_ctx_comm_init_nccl(self._h, uid, int(rank), int(world_size)))
        self.world, self.rank, self._ar_fn, self._cb = int(world_size), int(rank), None, None

    def transport(self) -> str:
        t = C.c_int(0)
        L.check(L.lib().scm_ctx_comm_info(self._h, None, None, C.byref(t)))
        return ("none", "nccl", "hook")[t.value]

    def fit_info(self) -> dict:
        """Rayleigh-Ritz steps, final subspace-angle estimate, converged rank and extra products of the last eigen-solve."""
        it, cr, pr, rs = C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_double(0.0)
        L.check(L.lib().scm_ctx_fit_info(self._h, C.byref(it), C.byref(rs), C.byref(cr), C.byref(pr)))
        return {"iters": it.value, "resid": rs.value, "conv_rank": cr.value, "products": pr.value}

    def allreduce(self, arr: "DeviceArray"):
        """Sum a float64 / int64 device array in place over the ranks (no-op for a single rank)."""
        if self.world <= 1:
            return
        if arr.dtype not in (np.dtype(np.float64), np.dtype(np.int64)):
            raise TypeError("all-reduce supports float64 and int64")
        L.check(L.lib().scm_allreduce_sum(self._h, arr.ptr, int(np.prod(arr.shape, dtype=np.int64)),
                                          0 if arr.dtype == np.dtype(np.float64) else 1))  # enqueued on the context's stream: reads after it are ordered

    def allgather_bytes(self, payload: bytes) -> list:
        """Every rank's byte string, in rank order, through the sum all-reduce (one-hot slices of a zero buffer).  Small
        payloads only (label tables): each byte travels as an int64."""
        if self.world <= 1:
            return [payload]
        if self.rank is None:
            raise ValueError("multi-rank host logic needs the rank: Context.set_allreduce(fn, world_size, rank) or init_nccl")
        sizes = np.zeros(self.world, dtype=np.int64)
        sizes[self.rank] = len(payload)
        d = self.to_device(sizes)
        self.allreduce(d)
        sizes = d.to_host()
        d.free()
        off = np.concatenate([[0], np.cumsum(sizes)])
        buf = np.zeros(int(off[-1]) + 1, dtype=np.int64)
        buf[off[self.rank]:off[self.rank + 1]] = np.frombuffer(payload, dtype=np.uint8)
        d = self.to_device(buf)
        self.allreduce(d)
        buf = d.to_host()
        d.free()
        return [buf[off[r]:off[r + 1]].astype(np.uint8).tobytes() for r in range(self.world)]

    # ---- memory -------------------------------------------------------------------------------
    def empty(self, shape, dtype=np.float64) -> "DeviceArray":
        return DeviceArray(self, shape, dtype)

    def to_device(self, host: np.ndarray) -> "DeviceArray":
        host = np.ascontiguousarray(host)
        d = DeviceArray(self, host.shape, host.dtype)
        if host.nbytes:
            L.check(L.lib().scm_h2d(self._h, d.ptr, host.ctypes.data_as(C.c_void_p), host.nbytes))
            self.sync()  # `host` may be a temporary
        return d


class MultiContext:
    """One process, several GPUs (scm_multi_create): a context + NCCL rank per device inside the library, one host thread per
    device per call.  This is the handle a single R process would hold; `ruv3_host` is scm_multi_ruv3_host."""

    def __init__(self, devices=None, ndev: int | None = None):
        self._h = C.c_void_p()
        if devices is None:
            n = int(ndev if ndev is not None else 1)
            arr = None
        else:
            devices = [int(d) for d in devices]
            n = len(devices)
            arr = (C.c_int * n)(*devices)
        L.check(L.lib().scm_multi_create(arr, n, C.byref(self._h)))
        self.size = n
        self._fin = weakref.finalize(self, L.lib().scm_multi_destroy, self._h)

    @property
    def handle(self):
        return self._h

    def ctx(self, i: int) -> Context:
        p = L.lib().scm_multi_ctx(self._h, int(i))
        if not p:
            raise IndexError(i)
        c = Context(_borrowed=p)
        c._owner = self  # keep the handle alive while the borrowed context is in use
        return c


def bind_host_to_gpu(device: int = 0):
    """Pin the calling process to the CPUs of the NUMA node the GPU is attached to (sysfs `local_cpulist` of its PCI
    device), so that page-locked buffers allocated afterwards (first touch) and the staging of pageable copies are local to
    the GPU's PCIe root.  Returns the cpu list applied, or None when the topology cannot be read (no-op)."""
    import os
    buf = C.create_string_buffer(32)
    try:
        L.check(L.lib().scm_device_pci_bus_id(int(device), buf, 32))
        bus = buf.value.decode().lower()
        txt = open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return txt
    except Exception:
        return None


_default_ctx: dict = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class DeviceArray:
    """A dense C-contiguous device buffer owned by the library (scm_dev_alloc / scm_dev_free)."""

    def __init__(self, ctx: Context, shape, dtype=np.float64):
        self.ctx = ctx
        self.shape = tuple(int(x) for x in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        self.dtype = np.dtype(dtype)
        if self.dtype not in _DT:
            raise TypeError(f"unsupported dtype {self.dtype}")
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        p = C.c_void_p()
        L.check(L.lib().scm_dev_alloc(ctx.handle, self.nbytes, C.byref(p)))
        self._p = p
        self._fin = weakref.finalize(self, _free, ctx, p)

    @property
    def ptr(self) -> C.c_void_p:
        return self._p

    @property
    def address(self) -> int:
        return int(self._p.value or 0)

    def to_host(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.shape, dtype=self.dtype)
        assert out.flags.c_contiguous and out.nbytes == self.nbytes
        if self.nbytes:
            L.check(L.lib().scm_d2h(self.ctx.handle, out.ctypes.data_as(C.c_void_p), self._p, self.nbytes))
        return out

    def free(self):
        self._fin()

    @property
    def __cuda_array_interface__(self):  # lets torch / cupy wrap the buffer without a copy
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self.address, False), "version": 3,
                "strides": None}


def _free(ctx: Context, p):
    try:
        L.lib().scm_dev_free(ctx.handle, p)
    except Exception:
        pass


class DeviceMatrix:
    """cells x genes expression matrix in the canonical device layout (row-major, leading dimension `ld` even
    so that every row is 16-byte aligned for the vector loads)."""

    def __init__(self, ctx: Context, m: int, n: int, ld: int | None = None, buf: DeviceArray | None = None):
        self.ctx, self.m, self.n = ctx, int(m), int(n)
        self.ld = int(ld) if ld is not None else (self.n + 1) // 2 * 2
        self.buf = buf if buf is not None else DeviceArray(ctx, (self.m * self.ld,), np.float64)

    @property
    def ptr(self):
        return self.buf.ptr

    @classmethod
    def from_host(cls, ctx: Context, Y: np.ndarray) -> "DeviceMatrix":
        """Y: 2-D float64, cells x genes.  C-order is copied as is; Fortran order (what R hands to a direct
        fastRUVIII(Y) call: cells x genes column-major) is uploaded and transposed on the device."""
        Y = np.asarray(Y)
        if Y.ndim != 2:
            raise ValueError("Y must be a 2-D matrix")
        if Y.dtype != np.float64:
            Y = Y.astype(np.float64)
        m, n = Y.shape
        out = cls(ctx, m, n)
        lib = L.lib()
        if m == 0 or n == 0:
            return out
        if Y.flags.c_contiguous:
            L.check(lib.scm_copy2d(ctx.handle, out.ptr, out.ld * 8, Y.ctypes.data_as(C.c_void_p), n * 8, n * 8, m, 0))
            ctx.sync()
        elif Y.flags.f_contiguous:
            tmp = ctx.to_device(Y.T)  # genes x cells row-major == the column-major bytes
            L.check(lib.scm_transpose(ctx.handle, tmp.ptr, n, m, m, out.ptr, out.ld))
            ctx.sync()
            tmp.free()
        else:
            return cls.from_host(ctx, np.ascontiguousarray(Y))
        return out

    def to_host(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.m, self.n), dtype=np.float64)
        assert out.shape == (self.m, self.n) and out.flags.c_contiguous and out.dtype == np.float64
        if self.m and self.n:
            L.check(L.lib().scm_copy2d(self.ctx.handle, out.ctypes.data_as(C.c_void_p), self.n * 8, self.ptr, self.ld * 8,
                                       self.n * 8, self.m, 1))
        return out

    def free(self):
        self.buf.free()


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """numpy array backed by page-locked host memory (cudaMallocHost): full-speed async H2D / D2H."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    p = C.c_void_p()
    L.check(L.lib().scm_host_alloc_pinned(nbytes, C.byref(p)))
    buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)
    weakref.finalize(buf, L.lib().scm_host_free_pinned, p)
    return arr
