"""Host-side mirror of mfmg's device operator interface on top of the C ABI.

Same names, argument meaning and error behaviour as the reference classes, so the parity tests
read like the reference's own tests (tests/test_sparse_matrix_device.cu, test_smoother_device.cu,
test_direct_solver_device.cu, test_hierarchy_device.cu):

  CudaHandle            source/cuda/cuda_handle.cu:17-56       (context; no library handles inside)
  DeviceVector          dealii::LinearAlgebra::distributed::Vector<double, MemorySpace::CUDA>
  SparseMatrixDevice    include/mfmg/cuda/sparse_matrix_device.cuh:28-104
  CudaMatrixOperator    source/cuda/cuda_matrix_operator.cu     (alias SparseMatrixDeviceOperator)
  CudaSmoother          source/cuda/cuda_smoother.cu            (alias SmootherDevice)
  CudaSolver            source/cuda/cuda_solver.cu              (alias DirectSolverDevice)
  Hierarchy             include/mfmg/common/hierarchy.hpp:159-309 (alias HierarchyDevice)

Everything numeric happens in libmfmg_b200.so; this module only holds handles.
"""
from __future__ import annotations

import ctypes
import os
from enum import Enum

import numpy as np

from . import _lib
from ._lib import MfmgError, NoConvergence, NotImplementedExc, check  # noqa: F401


class OperatorMode(Enum):
    """include/mfmg/common/operator.hpp:19-23"""
    NO_TRANS = 0
    TRANS = 1


def _get(params, key: str, default=None):
    """boost::property_tree-style lookup: nested dicts addressed by a dotted path."""
    if params is None:
        return default
    node = params
    if key in node:
        return node[key]
    for part in key.split("."):
        if not isinstance(node, dict) or part not in node:
            return default
        node = node[part]
    return node


class CudaHandle:
    """Owns the mfmgb context (device + stream)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = _lib.load()
        p = ctypes.c_void_p()
        rc = self.lib.mfmgb_ctx_create(device, ctypes.c_void_p(stream) if stream else None, ctypes.byref(p))
        check(None, rc)
        self.ctx = p
        self.device = device
        self.nranks, self.rank = None, None   # set by init_comm*

    def synchronize(self) -> None:
        check(self.ctx, self.lib.mfmgb_ctx_synchronize(self.ctx))

    @property
    def launch_count(self) -> int:
        return int(self.lib.mfmgb_ctx_launch_count(self.ctx))

    @property
    def stream(self) -> int:
        return int(self.lib.mfmgb_ctx_stream(self.ctx) or 0)

    # ---- multi-GPU: one process per GPU, NCCL communicator owned by the context ----
    def unique_id(self) -> bytes:
        """128-byte NCCL id (create on rank 0, broadcast with the launcher's own transport)."""
        buf = ctypes.create_string_buffer(128)
        check(self.ctx, self.lib.mfmgb_comm_unique_id(buf))
        return buf.raw

    def init_comm(self, unique_id: bytes, nranks: int, rank: int) -> None:
        check(self.ctx, self.lib.mfmgb_comm_init(self.ctx, unique_id, nranks, rank))
        self.nranks, self.rank = nranks, rank

    def init_comm_from_torch(self) -> None:
        """Rendezvous through an initialised torch.distributed process group (any backend)."""
        import torch.distributed as dist

        rank, world = dist.get_rank(), dist.get_world_size()
        obj = [self.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        self.init_comm(obj[0], world, rank)

    def close(self) -> None:
        if getattr(self, "ctx", None):
            self.lib.mfmgb_comm_finalize(self.ctx)
            self.lib.mfmgb_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def cuda_malloc(handle: "CudaHandle", nbytes: int) -> int:
    """cuda_malloc (include/mfmg/cuda/utils.cuh:66-73): a plain cudaMalloc'ed block, returned as an address."""
    p = ctypes.c_void_p()
    check(handle.ctx, handle.lib.mfmgb_dev_malloc(handle.ctx, int(nbytes), ctypes.byref(p)))
    return int(p.value or 0)


def cuda_free(handle: "CudaHandle", ptr: int) -> None:
    check(handle.ctx, handle.lib.mfmgb_dev_free(handle.ctx, ctypes.c_void_p(ptr)))


def cuda_mem_copy_to_dev(handle: "CudaHandle", host: np.ndarray, ptr: int) -> None:
    """cuda_mem_copy_to_dev (include/mfmg/cuda/utils.cuh:83-90)"""
    host = np.ascontiguousarray(host)
    check(handle.ctx, handle.lib.mfmgb_dev_upload(handle.ctx, ctypes.c_void_p(ptr), host.ctypes.data, host.nbytes))


def cuda_mem_copy_to_host(handle: "CudaHandle", ptr: int, host: np.ndarray) -> None:
    """cuda_mem_copy_to_host (include/mfmg/cuda/utils.cuh:92-99)"""
    assert host.flags["C_CONTIGUOUS"]
    check(handle.ctx, handle.lib.mfmgb_dev_download(handle.ctx, ctypes.c_void_p(ptr), host.ctypes.data, host.nbytes))


class DeviceVector:
    """A device-resident FP64 vector (raw cudaMalloc storage, like Vector<double, CUDA>::get_values())."""

    def __init__(self, handle: CudaHandle, n: int):
        self.handle = handle
        self.size = int(n)
        p = ctypes.c_void_p()
        check(handle.ctx, handle.lib.mfmgb_vec_alloc(handle.ctx, self.size, ctypes.byref(p)))
        self.ptr = p

    @staticmethod
    def from_host(handle: CudaHandle, a) -> "DeviceVector":
        a = np.ascontiguousarray(a, dtype=np.float64)
        v = DeviceVector(handle, a.shape[0])
        v.upload(a)
        return v

    def upload(self, a) -> None:
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.shape[0] != self.size:
            raise MfmgError(_lib.ERR_INVALID, "DeviceVector.upload: size mismatch")
        check(self.handle.ctx, self.handle.lib.mfmgb_vec_upload(self.handle.ctx, self.ptr, a.ctypes.data, self.size))

    def to_host(self) -> np.ndarray:
        out = np.empty(self.size, dtype=np.float64)
        check(self.handle.ctx,
              self.handle.lib.mfmgb_vec_download(self.handle.ctx, self.ptr, out.ctypes.data, self.size))
        return out

    def fill(self, value: float) -> None:
        check(self.handle.ctx, self.handle.lib.mfmgb_vec_fill(self.handle.ctx, self.ptr, float(value), self.size))

    def add(self, a: float, other: "DeviceVector") -> None:
        """Vector::add(a, v): this += a * v"""
        check(self.handle.ctx,
              self.handle.lib.mfmgb_vec_axpy(self.handle.ctx, self.ptr, float(a), other.ptr, self.size))

    def dot(self, other: "DeviceVector") -> float:
        r = ctypes.c_double()
        check(self.handle.ctx,
              self.handle.lib.mfmgb_vec_dot(self.handle.ctx, self.ptr, other.ptr, self.size, ctypes.byref(r)))
        return r.value

    def l2_norm(self) -> float:
        return float(np.sqrt(self.dot(self)))

    def free(self) -> None:
        if getattr(self, "ptr", None) and getattr(self.handle, "ctx", None):
            self.handle.lib.mfmgb_vec_free(self.handle.ctx, self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class SparseMatrixDevice:
    """Device CSR matrix.  Construct from host CSR arrays (convert_matrix semantics: the data is
    copied, source/cuda/utils.cu:39-168)."""

    def __init__(self, handle: CudaHandle, n_rows: int, n_cols: int, rowptr, col, val, _adopt=None):
        self.handle = handle
        lib = handle.lib
        p = ctypes.c_void_p()
        if _adopt is not None:
            self.ptr = _adopt
        else:
            rowptr = np.ascontiguousarray(rowptr)
            col = np.ascontiguousarray(col, dtype=np.int32)
            val = np.ascontiguousarray(val, dtype=np.float64)
            if rowptr.dtype == np.int32:
                rc = lib.mfmgb_csr_upload_i32(handle.ctx, n_rows, n_cols, rowptr.ctypes.data, col.ctypes.data,
                                              val.ctypes.data, ctypes.byref(p))
            else:
                rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
                rc = lib.mfmgb_csr_upload(handle.ctx, n_rows, n_cols, rowptr.ctypes.data, col.ctypes.data,
                                          val.ctypes.data, ctypes.byref(p))
            check(handle.ctx, rc)
            self.ptr = p
        a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        check(handle.ctx, lib.mfmgb_csr_info(self.ptr, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        self._m, self._n, self._nnz = a.value, b.value, c.value

    @staticmethod
    def from_host(handle: CudaHandle, A) -> "SparseMatrixDevice":
        """A: hostsetup.HostCSR or scipy.sparse matrix."""
        if hasattr(A, "tocsr"):
            A = A.tocsr()
            return SparseMatrixDevice(handle, A.shape[0], A.shape[1], A.indptr.astype(np.int64), A.indices, A.data)
        return SparseMatrixDevice(handle, A.n_rows, A.n_cols, A.rowptr, A.col, A.val)

    @staticmethod
    def from_device_arrays(handle: CudaHandle, val_dev: int, column_index_dev: int, row_ptr_dev: int, local_nnz: int,
                           n_rows: int, n_cols: int) -> "SparseMatrixDevice":
        """The reference's own constructor (sparse_matrix_device.templates.cuh:244-272): TAKES OWNERSHIP of three
        cudaMalloc'ed device arrays (double values, int columns, int row offsets) and frees them on destruction."""
        p = ctypes.c_void_p()
        check(handle.ctx, handle.lib.mfmgb_csr_adopt_device(handle.ctx, n_rows, n_cols, local_nnz,
                                                            ctypes.c_void_p(val_dev), ctypes.c_void_p(column_index_dev),
                                                            ctypes.c_void_p(row_ptr_dev), ctypes.byref(p)))
        return SparseMatrixDevice(handle, 0, 0, None, None, None, _adopt=p)

    # sparse_matrix_device.cuh:61-71
    def m(self) -> int:
        return self._m

    def n(self) -> int:
        return self._n

    def n_local_rows(self) -> int:
        return self._m

    def local_nnz(self) -> int:
        return self._nnz

    def n_nonzero_elements(self) -> int:
        return self._nnz

    def vmult(self, dst: DeviceVector, src: DeviceVector) -> None:
        """dst = A src (sparse_matrix_device.templates.cuh:351-371)"""
        if src.size != self._n or dst.size != self._m:
            raise MfmgError(_lib.ERR_INVALID, "SparseMatrixDevice.vmult: size mismatch")
        check(self.handle.ctx, self.handle.lib.mfmgb_spmv(self.handle.ctx, self.ptr, src.ptr, dst.ptr))

    def transpose(self) -> "SparseMatrixDevice":
        p = ctypes.c_void_p()
        check(self.handle.ctx, self.handle.lib.mfmgb_csr_transpose(self.handle.ctx, self.ptr, ctypes.byref(p)))
        return SparseMatrixDevice(self.handle, 0, 0, None, None, None, _adopt=p)

    def to_host(self):
        """(rowptr int64, col int32, val f64): round trip of tests/test_utils_device.cu:221-263."""
        rowptr = np.empty(self._m + 1, dtype=np.int64)
        col = np.empty(max(self._nnz, 1), dtype=np.int32)
        val = np.empty(max(self._nnz, 1), dtype=np.float64)
        check(self.handle.ctx, self.handle.lib.mfmgb_csr_download(self.handle.ctx, self.ptr, rowptr.ctypes.data,
                                                                  col.ctypes.data, val.ctypes.data))
        return rowptr, col[:self._nnz], val[:self._nnz]

    def set_lanes_per_row(self, lanes: int) -> None:
        check(self.handle.ctx, self.handle.lib.mfmgb_csr_set_lanes_per_row(self.ptr, lanes))

    @property
    def lanes_per_row(self) -> int:
        return int(self.handle.lib.mfmgb_csr_get_lanes_per_row(self.ptr))

    KERNEL_AUTO, KERNEL_VECTOR, KERNEL_TILE = -1, 0, 1

    def set_kernel(self, kernel: int) -> None:
        """Kernel family behind vmult & the fused epilogues: KERNEL_AUTO / KERNEL_VECTOR / KERNEL_TILE."""
        check(self.handle.ctx, self.handle.lib.mfmgb_csr_set_kernel(self.ptr, kernel))

    @property
    def kernel(self) -> str:
        return "tile" if int(self.handle.lib.mfmgb_csr_get_kernel(self.ptr)) == 1 else "vector"

    def free(self) -> None:
        if getattr(self, "ptr", None) and getattr(self.handle, "ctx", None):
            self.handle.lib.mfmgb_csr_destroy(self.handle.ctx, self.ptr)
        self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class CudaMatrixOperator:
    """mfmg::CudaMatrixOperator (Operator interface, include/mfmg/common/operator.hpp:25-52)."""

    def __init__(self, matrix: SparseMatrixDevice):
        self._matrix = matrix
        self._transposed = None
        self.handle = matrix.handle

    def get_matrix(self) -> SparseMatrixDevice:
        return self._matrix

    def apply(self, x: DeviceVector, y: DeviceVector, mode: OperatorMode = OperatorMode.NO_TRANS) -> None:
        if mode == OperatorMode.NO_TRANS:
            self._matrix.vmult(y, x)
        else:
            # explicit transpose built lazily on first use (cuda_matrix_operator.cu:84-88)
            if self._transposed is None:
                self._transposed = self._matrix.transpose()
            self._transposed.vmult(y, x)

    def transpose(self) -> "CudaMatrixOperator":
        if self._transposed is None:
            self._transposed = self._matrix.transpose()
        return CudaMatrixOperator(self._transposed)

    def multiply(self, b: "CudaMatrixOperator") -> "CudaMatrixOperator":
        """C = this * b.  SETUP operation; like the reference's parallel path it runs on the host
        (sparse_matrix_device.templates.cuh:417-433) and uploads the product."""
        import scipy.sparse as sp

        ra, ca, va = self._matrix.to_host()
        rb, cb, vb = b._matrix.to_host()
        A = sp.csr_matrix((va, ca, ra), shape=(self._matrix.m(), self._matrix.n()))
        B = sp.csr_matrix((vb, cb, rb), shape=(b._matrix.m(), b._matrix.n()))
        C = (A @ B).tocsr()
        C.sort_indices()
        return CudaMatrixOperator(SparseMatrixDevice.from_host(self.handle, C))

    def multiply_transpose(self, b: "CudaMatrixOperator") -> "CudaMatrixOperator":
        """this * b^T (cuda_matrix_operator.cu:151-225); setup, host."""
        return self.multiply(b.transpose())

    def build_domain_vector(self) -> DeviceVector:
        return DeviceVector(self.handle, self._matrix.n())

    def build_range_vector(self) -> DeviceVector:
        return DeviceVector(self.handle, self._matrix.m())

    def grid_complexity(self) -> int:
        return self._matrix.m()

    def operator_complexity(self) -> int:
        return self._matrix.n_nonzero_elements()


class CudaSmoother:
    """mfmg::CudaSmoother: Jacobi only (source/cuda/cuda_smoother.cu:99-172)."""

    def __init__(self, op: CudaMatrixOperator, params=None, omega: float = 1.0):
        prec_type = str(_get(params, "smoother.type", "Jacobi")).lower()
        if prec_type != "jacobi":
            raise MfmgError(_lib.ERR_INVALID, "Only Jacobi smoother is implemented.")  # cuda_smoother.cu:110
        self._operator = op
        self.handle = op.handle
        m = op.get_matrix()
        if m.m() != m.n():
            raise MfmgError(_lib.ERR_INVALID, f"The matrix is not square. The matrix is a {m.m()} by {m.n()} .")
        p = ctypes.c_void_p()
        check(self.handle.ctx, self.handle.lib.mfmgb_jacobi_setup(self.handle.ctx, m.ptr, float(omega), ctypes.byref(p)))
        self.ptr = p

    def apply(self, b: DeviceVector, x: DeviceVector) -> None:
        """x <- x - D^-1 (A x - b)"""
        check(self.handle.ctx, self.handle.lib.mfmgb_jacobi_apply(self.handle.ctx, self.ptr,
                                                                 self._operator.get_matrix().ptr, b.ptr, x.ptr))

    def __del__(self):
        try:
            if getattr(self, "ptr", None) and getattr(self.handle, "ctx", None):
                self.handle.lib.mfmgb_jacobi_destroy(self.handle.ctx, self.ptr)
        except Exception:
            pass


class CudaSolver:
    """mfmg::CudaSolver (source/cuda/cuda_solver.cu:196-515).  `solver.type`: "lu_dense" (default)
    is the hand-written dense factor-and-solve; "cholesky" and "lu_sparse_host" -- cuSOLVER wrappers
    in the reference -- are served by the same dense kernel (identical mathematical result);
    "amgx" is not available (the north star forbids library solvers)."""

    def __init__(self, handle: CudaHandle, op: CudaMatrixOperator, params=None):
        solver = str(_get(params, "solver.type", "lu_dense"))
        if solver == "amgx":
            raise NotImplementedExc(_lib.ERR_NOT_IMPLEMENTED, "solver.type amgx is not available in mfmg_b200")
        if solver not in ("lu_dense", "cholesky", "lu_sparse_host"):
            raise MfmgError(_lib.ERR_INVALID, f"The provided solver name {solver} is invalid.")  # cuda_solver.cu:70
        self.handle = handle
        self._operator = op
        p = ctypes.c_void_p()
        check(handle.ctx, handle.lib.mfmgb_dense_factor(handle.ctx, op.get_matrix().ptr, ctypes.byref(p)))
        self.ptr = p

    def apply(self, b: DeviceVector, x: DeviceVector) -> None:
        check(self.handle.ctx, self.handle.lib.mfmgb_dense_solve(self.handle.ctx, self.ptr, b.ptr, x.ptr))

    @property
    def num_swaps(self) -> int:
        return int(self.handle.lib.mfmgb_dense_num_swaps(self.ptr))

    @property
    def solve_mode(self):
        """("inverse" | "substitution", min |u_kk| / max |u_kk|)"""
        r = ctypes.c_double()
        m = self.handle.lib.mfmgb_dense_solve_mode(self.ptr, ctypes.byref(r))
        return ("substitution" if m else "inverse"), r.value

    def __del__(self):
        try:
            if getattr(self, "ptr", None) and getattr(self.handle, "ctx", None):
                self.handle.lib.mfmgb_dense_destroy(self.handle.ctx, self.ptr)
        except Exception:
            pass


class MatrixFreeLaplaceDevice:
    """The matrix-free fine-level operator (CudaMatrixFreeOperator slot)."""

    KERNELS = {0: "generic colour-phase cell kernel", 1: "Q1 node-owner z-sweep, per-cell coefficient",
               2: "Q1 node-owner z-sweep, per-quadrature-point coefficient",
               3: "Q1 constant-coefficient factorised 27-point stencil, persistent z-sweep"}

    def __init__(self, handle: CudaHandle, dim, degree, cells, h, coef_per_q, constrained, own_planes=None):
        """own_planes = (begin, end): the z-slab form of a row-partitioned grid -- `cells` is the local box, node
        planes [begin, end) are owned, vectors / `constrained` are laid out [owned | ghost below | ghost above]."""
        self.handle = handle
        cells_a = np.ascontiguousarray(list(cells) + [1] * (3 - dim), dtype=np.int64)
        h_a = np.ascontiguousarray(list(h) + [1.0] * (3 - dim), dtype=np.float64)
        coef = np.ascontiguousarray(coef_per_q, dtype=np.float64)
        constr = np.ascontiguousarray(constrained, dtype=np.uint8)
        p = ctypes.c_void_p()
        if own_planes is None:
            check(handle.ctx, handle.lib.mfmgb_mf_laplace_create(handle.ctx, dim, degree, cells_a.ctypes.data,
                                                                 h_a.ctypes.data, coef.ctypes.data, constr.ctypes.data,
                                                                 ctypes.byref(p)))
        else:
            check(handle.ctx, handle.lib.mfmgb_mf_laplace_create_slab(
                handle.ctx, dim, degree, cells_a.ctypes.data, h_a.ctypes.data, coef.ctypes.data, constr.ctypes.data,
                int(own_planes[0]), int(own_planes[1]), ctypes.byref(p)))
        self.ptr = p
        self.size = int(handle.lib.mfmgb_mf_size(p))
        self.vector_size = int(handle.lib.mfmgb_mf_vector_size(p))
        self.kernel = self.KERNELS[int(handle.lib.mfmgb_mf_kernel(p))]

    def apply(self, x: DeviceVector, y: DeviceVector, mode: OperatorMode = OperatorMode.NO_TRANS) -> None:
        # the operator is symmetric: TRANS == NO_TRANS (cuda_matrix_free_operator.cu:60-70)
        check(self.handle.ctx, self.handle.lib.mfmgb_mf_apply(self.handle.ctx, self.ptr, x.ptr, y.ptr))

    def diagonal(self) -> DeviceVector:
        d = DeviceVector(self.handle, self.size)
        check(self.handle.ctx, self.handle.lib.mfmgb_mf_diagonal(self.handle.ctx, self.ptr, d.ptr))
        return d

    def build_domain_vector(self) -> DeviceVector:
        return DeviceVector(self.handle, self.size)

    build_range_vector = build_domain_vector

    def __del__(self):
        try:
            if getattr(self, "ptr", None) and getattr(self.handle, "ctx", None):
                self.handle.lib.mfmgb_mf_destroy(self.handle.ctx, self.ptr)
        except Exception:
            pass


class HaloPlan:
    """Halo plan of a row-partitioned level (hostsetup.partition.LocalPart -> mfmgb_halo)."""

    def __init__(self, handle: CudaHandle, part):
        self.handle = handle
        nb = np.ascontiguousarray(part.neighbors, dtype=np.int32)
        sc = np.ascontiguousarray([len(s) for s in part.send_indices], dtype=np.int64)
        si = np.ascontiguousarray(np.concatenate(part.send_indices) if len(part.send_indices) else np.zeros(0),
                                  dtype=np.int32)
        rc = np.ascontiguousarray(part.recv_counts, dtype=np.int64)
        p = ctypes.c_void_p()
        check(handle.ctx, handle.lib.mfmgb_halo_create(handle.ctx, part.n_owned, part.n_ghost, len(nb),
                                                       nb.ctypes.data, sc.ctypes.data, si.ctypes.data, rc.ctypes.data,
                                                       ctypes.byref(p)))
        self.ptr = p
        self.n_owned, self.n_ghost = part.n_owned, part.n_ghost

    def exchange(self, v: DeviceVector) -> None:
        assert v.size >= self.n_owned + self.n_ghost
        check(self.handle.ctx, self.handle.lib.mfmgb_halo_exchange(self.handle.ctx, self.ptr, v.ptr))

    def __del__(self):
        try:
            if getattr(self, "ptr", None) and getattr(self.handle, "ctx", None):
                self.handle.lib.mfmgb_halo_destroy(self.handle.ctx, self.ptr)
        except Exception:
            pass


class CoarseDD:
    """Domain-decomposed dense coarse solve of a row-partitioned hierarchy (csrc/coarse_dd.cu) from a
    hostsetup.coarse_dd_plan; a collective constructor (every rank of the communicator must call it)."""

    def __init__(self, handle: CudaHandle, plan):
        self.handle = handle
        self.blocks = [SparseMatrixDevice.from_host(handle, plan[k]) for k in ("A_II", "A_IS", "A_SI", "A_SS")]
        sep = np.ascontiguousarray(plan["sep_index"], dtype=np.int32)
        p = ctypes.c_void_p()
        check(handle.ctx, handle.lib.mfmgb_coarse_dd_create(
            handle.ctx, int(plan["n_c"]), int(plan["own_begin"]), int(plan["n_S"]), int(plan["adj_begin"]),
            int(plan["own_sep_begin"]), int(plan["own_sep_n"]), self.blocks[0].ptr, self.blocks[1].ptr,
            self.blocks[2].ptr, self.blocks[3].ptr, sep.ctypes.data, ctypes.byref(p)))
        self.ptr = p
        self.n_interior, self.n_separator = int(plan["n_I"]), int(plan["n_S"])
        self.n_adjacent = int(plan["A_IS"].n_cols)
        # bytes one solve reads on this rank: A_II^-1, the Schur inverse, E = A_II^-1 A_IS
        self.bytes_per_solve = 8 * (self.n_interior ** 2 + self.n_separator ** 2 + self.n_interior * self.n_adjacent)

    def solve(self, b_c: DeviceVector, x_c: DeviceVector) -> None:
        check(self.handle.ctx, self.handle.lib.mfmgb_coarse_dd_solve(self.handle.ctx, self.ptr, b_c.ptr, x_c.ptr))

    def __del__(self):
        try:
            if getattr(self, "ptr", None) and getattr(self.handle, "ctx", None):
                self.handle.lib.mfmgb_coarse_dd_destroy(self.handle.ctx, self.ptr)
        except Exception:
            pass


class Hierarchy:
    """mfmg::Hierarchy on device vectors.  Built from already-assembled level operators
    (`from_operators`) -- the setup that produces them stays on the host path
    (mfmg_b200.hostsetup), exactly as the north star prescribes.

    params keys honoured (include/mfmg/common/hierarchy.hpp:168-172): "is preconditioner",
    "smoother.n_smoothing_steps", "smoother.type", "solver.type"."""

    def __init__(self, handle: CudaHandle, operators, restrictors, params=None, omega: float = 1.0,
                 prolongators=None, halo: "HaloPlan | None" = None, boundary=(0, 0), coarse_offsets=None,
                 coarse_dd: "CoarseDD | None" = None, restrict_split: int = 0, restrict_no_halo: bool = False,
                 restrict_below: "SparseMatrixDevice | None" = None):
        self.handle = handle
        self.restrict_below = restrict_below
        self.halo = halo
        self.coarse_dd = coarse_dd
        lib = handle.lib
        self.params = params or {}
        smoother = str(_get(params, "smoother.type", "Jacobi")).lower()
        if smoother not in ("jacobi", "chebyshev"):
            # CudaSmoother: "Only Jacobi smoother is implemented." (cuda_smoother.cu:110); Chebyshev is the smoother of
            # the reference's matrix-free path (dealii_matrix_free_smoother.cc:34-60), available here on the device
            raise MfmgError(_lib.ERR_INVALID, f'Unknown smoother name: "{smoother}"')
        solver = str(_get(params, "solver.type", "lu_dense"))
        if solver == "amgx":
            raise NotImplementedExc(_lib.ERR_NOT_IMPLEMENTED, "solver.type amgx is not available in mfmg_b200")
        self.is_preconditioner = bool(_get(params, "is preconditioner", True))
        self.n_smoothing_steps = int(_get(params, "smoother.n_smoothing_steps", 1))
        self.operators = list(operators)
        self.restrictors = list(restrictors)
        self.prolongators = list(prolongators) if prolongators else [None] * len(self.restrictors)
        n_levels = len(self.operators)
        if len(self.restrictors) != n_levels - 1:
            raise MfmgError(_lib.ERR_INVALID, "Hierarchy: need one restrictor per level transition")
        p = ctypes.c_void_p()
        check(handle.ctx, lib.mfmgb_hierarchy_create(handle.ctx, n_levels, self.n_smoothing_steps,
                                                     int(self.is_preconditioner), float(omega), ctypes.byref(p)))
        self.ptr = p
        for li, op in enumerate(self.operators):
            if isinstance(op, MatrixFreeLaplaceDevice):
                check(handle.ctx, lib.mfmgb_hierarchy_set_mf_operator(self.ptr, op.ptr))
            else:
                check(handle.ctx, lib.mfmgb_hierarchy_set_operator(self.ptr, li, op.ptr))
        if smoother == "chebyshev":
            check(handle.ctx, lib.mfmgb_hierarchy_set_smoother_chebyshev(
                self.ptr, int(_get(params, "smoother.degree", 0)), float(_get(params, "smoother.smoothing_range", 0.0)),
                float(_get(params, "smoother.max_eigenvalue", 1.0)), int(_get(params, "smoother.eig_cg_n_iterations", 8))))
        if coarse_offsets is not None:  # (marks the hierarchy as row-partitioned before R / P are checked)
            co = np.ascontiguousarray(coarse_offsets, dtype=np.int64)
            check(handle.ctx, lib.mfmgb_hierarchy_set_coarse_offsets(self.ptr, co.ctypes.data, len(co) - 1))
        for li, r in enumerate(self.restrictors):
            pr = self.prolongators[li]
            check(handle.ctx, lib.mfmgb_hierarchy_set_restrictor(self.ptr, li + 1, r.ptr, pr.ptr if pr else None))
        if restrict_split:
            check(handle.ctx, lib.mfmgb_hierarchy_set_restrict_split(self.ptr, 1, int(restrict_split)))
        if halo is not None:
            check(handle.ctx, lib.mfmgb_hierarchy_set_halo(self.ptr, 0, halo.ptr, int(boundary[0]), int(boundary[1])))
        if coarse_dd is not None:
            check(handle.ctx, lib.mfmgb_hierarchy_set_coarse_dd(self.ptr, coarse_dd.ptr))
        if restrict_no_halo:
            check(handle.ctx, lib.mfmgb_hierarchy_set_restrict_no_halo(
                handle.ctx, self.ptr, 1, restrict_below.ptr if restrict_below is not None else None))
        check(handle.ctx, lib.mfmgb_hierarchy_finalize(handle.ctx, self.ptr))
        self.n = self.operators[0].size if isinstance(self.operators[0], MatrixFreeLaplaceDevice) \
            else self.operators[0].m()
        # length of gathered vectors (owned + ghost tail); == n on a single GPU
        self.vector_size = int(lib.mfmgb_hierarchy_vector_size(self.ptr, 0))

    @staticmethod
    def from_host(handle: CudaHandle, A, R, A_c, params=None, omega: float = 1.0) -> "Hierarchy":
        """Two-level hierarchy from host CSR operators (A fine, R restrictor, A_c = R A R^T)."""
        ops = [SparseMatrixDevice.from_host(handle, A), SparseMatrixDevice.from_host(handle, A_c)]
        res = [SparseMatrixDevice.from_host(handle, R)]
        return Hierarchy(handle, ops, res, params, omega)

    @staticmethod
    def from_partition(handle: CudaHandle, part, params=None, omega: float = 1.0,
                       matrix_free: bool = False, coarse_dd: "bool | str" = "auto") -> "Hierarchy":
        """Row-partitioned two-level hierarchy of one rank (hostsetup.partition.LocalPart); the context must have
        an initialised communicator (CudaHandle.init_comm*).  matrix_free: level 0 is the matrix-free slab operator
        (part.mf, hostsetup.slab) instead of the assembled rows; R, P and A_c stay assembled."""
        if matrix_free:
            mf = part.mf
            fine = MatrixFreeLaplaceDevice(handle, 3, mf["degree"], mf["cells"], mf["h"], mf["coef"],
                                           mf["constrained"], own_planes=mf["own_planes"])
        else:
            fine = SparseMatrixDevice.from_host(handle, part.A)
        ops = [fine, SparseMatrixDevice.from_host(handle, part.Ac)]
        pro = [SparseMatrixDevice.from_host(handle, part.P)]
        plan = HaloPlan(handle, part)
        # coarsest level: the domain-decomposed direct solve when the coarse operator is block tridiagonal in the
        # ranks (z-slab partitions are) -- per-GPU work independent of the number of ranks; else the dense inverse,
        # replicated or split by rows.  Every rank takes the same decision (it only depends on replicated data).
        dd = None
        if coarse_dd and part.world > 1 and os.environ.get("MFMGB_COARSE_DD", "1") != "0":
            from .hostsetup import coarse_dd_plan

            ddp = coarse_dd_plan(part.Ac, part.coarse_offsets, part.rank)
            if ddp is not None:
                if not np.all(np.isin(part.P.col, ddp["valid_cols"])):
                    raise MfmgError(_lib.ERR_INVALID, "from_partition: prolongation rows read coarse entries outside "
                                                      "this rank's rows and the separators")
                dd = CoarseDD(handle, ddp)
            elif coarse_dd is True:
                raise MfmgError(_lib.ERR_INVALID, "from_partition: the coarse operator is not block tridiagonal in the "
                                                  "ranks' row blocks; the domain-decomposed coarse solve cannot be used")
        # With the domain-decomposed coarse solve the residual's halo is not exchanged at all: the entries of R on the
        # ghost plane are dropped and the neighbour above supplies them through its R_below rows, which join the
        # solver's all-reduce (MFMGB_RESTRICT_NO_HALO=0 keeps the exchange).
        R_host, r_below, no_halo = part.R, None, False
        if dd is not None and os.environ.get("MFMGB_RESTRICT_NO_HALO", "1") != "0" and \
                (part.rank == 0 or part.R_below is not None):
            from .hostsetup.partition import rows_restricted

            no_halo = True
            R_host = rows_restricted(part.R, 0, part.R.n_rows, 0, part.n_owned, part.R.n_cols)
            if part.rank > 0:
                f0, nb = ddp["sep_below_first_row"], ddp["n_sep_below"]
                Rb = part.R_below
                if int(Rb.rowptr[f0]) != 0 or Rb.n_rows - f0 != nb:
                    raise MfmgError(_lib.ERR_INVALID, "from_partition: this rank's entries reach restrictor rows of the "
                                                      "lower neighbour outside its separator")
                r_below = SparseMatrixDevice.from_host(
                    handle, rows_restricted(Rb, f0, Rb.n_rows, 0, part.n_owned, part.R.n_cols))
        res = [SparseMatrixDevice.from_host(handle, R_host)]
        # (mfmgb_hierarchy_set_restrict_split could overlap the rows of R that only read owned entries with the
        # exchange of the residual's halo; measured slower on 2 GPUs -- 0.047 vs 0.033 ms for the stage -- so it is
        # left off: restrict_split=0)
        r_split = 0
        return Hierarchy(handle, ops, res, params, omega, prolongators=pro, halo=plan,
                         boundary=(part.boundary_lo, part.boundary_hi), coarse_offsets=part.coarse_offsets,
                         coarse_dd=dd, restrict_split=r_split, restrict_no_halo=no_halo, restrict_below=r_below)

    @property
    def transport(self) -> str:
        """How the ghost entries and the coarse reduction travel between the GPUs of this hierarchy."""
        fn = getattr(self.handle.lib, "mfmgb_comm_transport", None)
        if self.halo is None:
            return "none (single GPU)"
        return fn(self.handle.ctx).decode() if fn else "NCCL send/recv + all-reduce"

    def build_vector(self) -> DeviceVector:
        """A level-0 vector with room for the ghost tail (Level::build_vector, level.hpp:63-70)."""
        return DeviceVector(self.handle, self.vector_size)

    def use_graph(self, on: bool = True) -> None:
        check(self.handle.ctx, self.handle.lib.mfmgb_hierarchy_use_graph(self.ptr, int(on)))

    @property
    def launches_per_cycle(self) -> int:
        return int(self.handle.lib.mfmgb_hierarchy_launches_per_cycle(self.ptr))

    def vmult(self, x: DeviceVector, b: DeviceVector) -> None:
        """x = V-cycle(b)  (Hierarchy::vmult, hierarchy.hpp:238-244)"""
        check(self.handle.ctx, self.handle.lib.mfmgb_vcycle(self.handle.ctx, self.ptr, b.ptr, x.ptr))

    def apply(self, b: DeviceVector, x: DeviceVector, level_index: int = 0) -> None:
        check(self.handle.ctx, self.handle.lib.mfmgb_hierarchy_apply(self.handle.ctx, self.ptr, b.ptr, x.ptr,
                                                                    level_index))

    def vmult_host(self, x_host: np.ndarray, b_host: np.ndarray) -> None:
        """Hierarchy<Vector<double, Host>>::vmult: host arrays in/out (H2D + cycle + D2H)."""
        assert x_host.dtype == np.float64 and b_host.dtype == np.float64
        check(self.handle.ctx, self.handle.lib.mfmgb_vcycle_host(self.handle.ctx, self.ptr, b_host.ctypes.data,
                                                                 x_host.ctypes.data))

    def vmult_host_ptr(self, x_ptr: int, b_ptr: int) -> None:
        check(self.handle.ctx, self.handle.lib.mfmgb_vcycle_host(self.handle.ctx, self.ptr, ctypes.c_void_p(b_ptr),
                                                                 ctypes.c_void_p(x_ptr)))

    def vmult_host_batch_ptr(self, x_ptrs, b_ptrs) -> None:
        """x_j = V-cycle(b_j) for a batch of independent HOST vectors (addresses of pinned buffers), pipelined:
        H2D of j+1, the cycle of j and D2H of j-1 overlap (mfmgb_vcycle_host_batch)."""
        n = len(b_ptrs)
        assert len(x_ptrs) == n
        bp = (ctypes.c_void_p * n)(*[ctypes.c_void_p(p) for p in b_ptrs])
        xp = (ctypes.c_void_p * n)(*[ctypes.c_void_p(p) for p in x_ptrs])
        check(self.handle.ctx, self.handle.lib.mfmgb_vcycle_host_batch(self.handle.ctx, self.ptr, n, bp, xp))

    def vmult_host_batch(self, xs, bs) -> None:
        """The same for lists of float64 numpy arrays."""
        self.vmult_host_batch_ptr([x.ctypes.data for x in xs], [b.ctypes.data for b in bs])

    STAGES = ("pre_smooth", "residual", "restrict", "coarse", "prolong_correct", "post_smooth")

    def profile(self, x: DeviceVector, b: DeviceVector) -> dict:
        """One un-captured V-cycle with CUDA events between the level-0 stages -> {stage: ms}."""
        ms = np.zeros(6)
        check(self.handle.ctx, self.handle.lib.mfmgb_vcycle_profile(self.handle.ctx, self.ptr, b.ptr, x.ptr,
                                                                    ms.ctypes.data))
        return dict(zip(self.STAGES, ms.tolist()))

    def chebyshev_info(self, level: int = 0):
        """(lambda_min, lambda_max incl. the safety factor 1.2, theta, delta) of a level's Chebyshev smoother."""
        out = np.zeros(4)
        check(self.handle.ctx, self.handle.lib.mfmgb_hierarchy_chebyshev_info(self.ptr, level, out.ctypes.data))
        return tuple(out)

    def timeline(self, x: DeviceVector, b: DeviceVector, graph: bool = True):
        """[(piece, ms), ...] of one V-cycle, resolved per kernel group; graph=True: times of a CUDA-graph replay."""
        names = ctypes.create_string_buffer(4096)
        ms = np.zeros(128)
        n = ctypes.c_int(0)
        check(self.handle.ctx, self.handle.lib.mfmgb_vcycle_timeline(self.handle.ctx, self.ptr, b.ptr, x.ptr, int(graph),
                                                                     names, 4096, ms.ctypes.data, 128, ctypes.byref(n)))
        labels = names.value.decode().split(";") if n.value else []
        return list(zip(labels, ms[:n.value].tolist()))

    def grid_complexity(self) -> float:
        sizes = [op.size if isinstance(op, MatrixFreeLaplaceDevice) else op.m() for op in self.operators]
        return sum(sizes) / sizes[0]

    def __del__(self):
        try:
            if getattr(self, "ptr", None) and getattr(self.handle, "ctx", None):
                self.handle.lib.mfmgb_hierarchy_destroy(self.handle.ctx, self.ptr)
        except Exception:
            pass


def solver_cg(handle: CudaHandle, A: SparseMatrixDevice | None, x: DeviceVector, b: DeviceVector,
              hierarchy: Hierarchy | None, tol: float = 1e-6, max_it: int | None = None):
    """dealii::SolverCG(SolverControl(max_it, tol)).solve(A, x, b, hierarchy)
    (tests/hierarchy_driver.cc:200-213).  Returns (last_step, residual_history).  Raises
    NoConvergence like deal.II when max_it is exhausted."""
    n = x.size
    if max_it is None:
        max_it = n
    hist = np.zeros(max_it + 1)
    it = ctypes.c_int(0)
    rc = handle.lib.mfmgb_pcg(handle.ctx, hierarchy.ptr if hierarchy else None, A.ptr if A else None, b.ptr, x.ptr,
                              float(tol), int(max_it), ctypes.byref(it), hist.ctypes.data)
    if rc == _lib.ERR_NOT_CONVERGED:
        msg = handle.lib.mfmgb_last_error(handle.ctx).decode()
        raise NoConvergence(rc, msg, it.value, hist[:it.value + 1].copy())
    check(handle.ctx, rc)
    return it.value, hist[:it.value + 1].copy()


# names used by BASELINE.json / older mfmg snapshots (SURVEY.md section 8b "Name mapping")
SparseMatrixDeviceOperator = CudaMatrixOperator
SmootherDevice = CudaSmoother
DirectSolverDevice = CudaSolver
HierarchyDevice = Hierarchy
