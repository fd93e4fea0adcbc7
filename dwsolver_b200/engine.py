"""ctypes mirror of the engine's C ABI (include/dwgpu.h -> dwsolver_b200/lib/libdwgpu.so).

The names, argument meaning and error behaviour follow the C header one to one; the Python
layer only marshals numpy arrays.  There is NO fallback: if the shared library is missing or
no CUDA device is usable this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# DWGPU_LIB: another build of the same library (tuning experiments, e.g. a different -DDWK_TG_OCC)
LIB_PATH = os.environ.get("DWGPU_LIB") or os.path.join(_HERE, "lib", "libdwgpu.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)

DWGPU_UNDEF, DWGPU_NOFEAS, DWGPU_OPT, DWGPU_UNBND = 1, 4, 5, 6

#: every symbol include/dwgpu.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = [
    "dwgpu_options_init", "dwgpu_create", "dwgpu_destroy", "dwgpu_initial_solve", "dwgpu_iterate",
    "dwgpu_iterate_device", "dwgpu_device_results", "dwgpu_fetch_results", "dwgpu_fetch_x",
    "dwgpu_get_counters", "dwgpu_get_block_iterations", "dwgpu_last_launch_count", "dwgpu_get_block_paths",
    "dwgpu_sync", "dwgpu_last_error", "dwgpu_version", "dwgpu_last_kernel_ms", "dwgpu_measure_smem_bandwidth",
    "dwgpu_reset", "dwgpu_initial_solve_device", "dwgpu_initial_solve_packed", "dwgpu_iterate_packed", "dwgpu_pack_gathered",
    "dwgpu_device_record", "dwgpu_pack_gathered_records", "dwgpu_history_push", "dwgpu_history_combine",
    "dwgpu_initial_solve_mailboxes", "dwgpu_iterate_mailboxes", "dwgpu_device_count",
    # the restricted master on the device (dwsolver_b200/master.py)
    "dwgpu_master_create", "dwgpu_master_destroy", "dwgpu_master_set_rhs", "dwgpu_master_add_cols",
    "dwgpu_master_set_costs", "dwgpu_master_set_coef", "dwgpu_master_del_cols", "dwgpu_master_cols",
    "dwgpu_master_solve", "dwgpu_master_get", "dwgpu_master_quality", "dwgpu_master_last_error",
    # multi-GPU rounds from the C layer: NCCL + the NVLink peer-memory up-link
    "dwgpu_comm_unique_id", "dwgpu_comm_init", "dwgpu_comm_release", "dwgpu_uplink_create", "dwgpu_uplink_attach",
    "dwgpu_round_comm",
    # several dual vectors per round
    "dwgpu_iterate_multi", "dwgpu_dense_pricing_blocks",
]


class BlockDesc(C.Structure):
    _fields_ = [("m", C.c_int), ("n", C.c_int),
                ("a_rowptr", c_ip), ("a_col", c_ip), ("a_val", c_dp),
                ("row_lb", c_dp), ("row_ub", c_dp), ("col_lb", c_dp), ("col_ub", c_dp),
                ("c_file", c_dp), ("c_master", c_dp),
                ("d_nnz", C.c_int), ("d_row", c_ip), ("d_col", c_ip), ("d_val", c_dp)]


class Options(C.Structure):
    _fields_ = [("device", C.c_int), ("clamp_lo_cols", C.c_int), ("force_path", C.c_int),
                ("max_iter_mult", C.c_int), ("tol_dj", C.c_double), ("tol_bnd", C.c_double),
                ("tol_piv", C.c_double), ("cluster_size", C.c_int), ("no_dual_start", C.c_int), ("reserved", C.c_int * 6)]


class Proposals(C.Structure):
    _fields_ = [("obj", c_dp), ("cost", c_dp), ("status", c_ip), ("len", c_ip), ("ind", c_ip),
                ("val", c_dp), ("x", c_dp)]


class Packed(C.Structure):
    _fields_ = [("obj", c_dp), ("cost", c_dp), ("status", c_ip), ("colptr", c_ip), ("ind", c_ip), ("val", c_dp),
                ("nnz", C.c_longlong)]


class Counters(C.Structure):
    _fields_ = [("pivots", C.c_longlong), ("flips", C.c_longlong), ("rebuilds", C.c_longlong),
                ("phase1_iter", C.c_longlong), ("solves", C.c_longlong), ("rows_swept", C.c_longlong),
                ("cells_swept", C.c_longlong), ("price_cells", C.c_longlong), ("mailbox_bytes", C.c_longlong)]


class Mailboxes(C.Structure):
    _fields_ = [("obj", c_dp), ("cost", c_dp), ("val", c_dp), ("status", c_ip), ("len", c_ip), ("ind", c_ip),
                ("K", C.c_int), ("L", C.c_int)]


_LIB = None


def load_library():
    """Load libdwgpu.so (raises if it has not been built: there is no CPU fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing — build it with `make -C dwsolver_b200/csrc` "
                               "(or __graft_entry__.build()); the engine has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.dwgpu_options_init.argtypes = [C.POINTER(Options)]
        L.dwgpu_options_init.restype = None
        L.dwgpu_create.argtypes = [C.c_int, C.c_int, C.POINTER(BlockDesc), C.POINTER(Options), C.POINTER(C.c_void_p)]
        L.dwgpu_destroy.argtypes = [C.c_void_p]
        L.dwgpu_destroy.restype = None
        L.dwgpu_initial_solve.argtypes = [C.c_void_p, C.POINTER(Proposals)]
        L.dwgpu_iterate.argtypes = [C.c_void_p, c_dp, C.c_int, C.POINTER(Proposals)]
        L.dwgpu_iterate_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.dwgpu_device_results.argtypes = [C.c_void_p, C.POINTER(Proposals)]
        L.dwgpu_fetch_results.argtypes = [C.c_void_p, C.POINTER(Proposals)]
        L.dwgpu_fetch_x.argtypes = [C.c_void_p, C.c_int, c_dp]
        L.dwgpu_get_counters.argtypes = [C.c_void_p, C.POINTER(Counters)]
        L.dwgpu_get_block_iterations.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
        L.dwgpu_last_launch_count.argtypes = [C.c_void_p]
        L.dwgpu_get_block_paths.argtypes = [C.c_void_p, c_ip]
        L.dwgpu_sync.argtypes = [C.c_void_p]
        L.dwgpu_reset.argtypes = [C.c_void_p, C.c_void_p]
        L.dwgpu_initial_solve_device.argtypes = [C.c_void_p, C.c_void_p]
        L.dwgpu_initial_solve_packed.argtypes = [C.c_void_p, C.POINTER(Packed)]
        L.dwgpu_iterate_packed.argtypes = [C.c_void_p, c_dp, C.c_int, c_dp, C.POINTER(Packed)]
        L.dwgpu_pack_gathered.argtypes = [C.c_void_p, C.c_int, C.POINTER(Proposals), C.c_void_p, C.c_void_p,
                                          C.POINTER(Packed)]
        L.dwgpu_device_record.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.dwgpu_comm_unique_id.argtypes = [C.c_void_p]
        L.dwgpu_comm_init.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        L.dwgpu_comm_release.argtypes = [C.c_void_p]
        L.dwgpu_comm_release.restype = None
        L.dwgpu_uplink_create.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(Proposals)]
        L.dwgpu_uplink_attach.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        L.dwgpu_round_comm.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.dwgpu_iterate_multi.argtypes = [C.c_void_p, c_dp, C.c_int, C.c_int, C.POINTER(Packed), C.POINTER(C.c_longlong)]
        L.dwgpu_dense_pricing_blocks.argtypes = [C.c_void_p]
        L.dwgpu_pack_gathered_records.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Packed)]
        L.dwgpu_history_push.argtypes = [C.c_void_p, C.c_int, c_ip, C.POINTER(C.c_longlong)]
        L.dwgpu_history_combine.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_longlong), c_dp, c_dp]
        L.dwgpu_initial_solve_mailboxes.argtypes = [C.c_void_p, C.POINTER(Mailboxes)]
        L.dwgpu_iterate_mailboxes.argtypes = [C.c_void_p, c_dp, C.c_int, C.POINTER(Mailboxes)]
        L.dwgpu_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.dwgpu_measure_smem_bandwidth.argtypes = [C.c_int, c_dp]
        L.dwgpu_last_error.restype = C.c_char_p
        L.dwgpu_version.restype = C.c_char_p
        _LIB = L
    return _LIB


class DwGpuError(RuntimeError):
    pass


def measure_smem_bandwidth(device: int = 0) -> float:
    """GB/s, read+write, see dwgpu_measure_smem_bandwidth in include/dwgpu.h."""
    g = C.c_double(0.0)
    rc = load_library().dwgpu_measure_smem_bandwidth(device, C.byref(g))
    if rc != 0:
        raise DwGpuError(f"dwgpu_measure_smem_bandwidth failed: {load_library().dwgpu_last_error().decode()}")
    return g.value


def _check(rc: int, what: str):
    if rc != 0:
        msg = load_library().dwgpu_last_error().decode()
        raise DwGpuError(f"{what} failed (code {rc}): {msg}")


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _bnd(a):
    a = _f64(a).copy()
    a[np.isposinf(a)] = 1e30
    a[np.isneginf(a)] = -1e30
    return a


class Engine:
    """Batched subproblem engine for one BlockAngularLP shard on one GPU."""

    def __init__(self, prob, device: int = 0, block_range: Optional[range] = None, fetch_x: bool = True, **opts):
        lib = load_library()
        blocks = prob.blocks if block_range is None else [prob.blocks[k] for k in block_range]
        self.K, self.L = len(blocks), prob.L
        self.n = np.asarray([b.n for b in blocks], np.int64)
        self.m = np.asarray([b.m for b in blocks], np.int64)
        self.xoff = np.concatenate([[0], np.cumsum(self.n)])
        keep = []
        descs = (BlockDesc * self.K)()
        for k, b in enumerate(blocks):
            arrs = dict(a_rowptr=_i32(b.a_rowptr), a_col=_i32(b.a_col), a_val=_f64(b.a_val),
                        row_lb=_bnd(b.row_lb), row_ub=_bnd(b.row_ub), col_lb=_bnd(b.col_lb), col_ub=_bnd(b.col_ub),
                        c_file=_f64(b.c_file), c_master=_f64(b.c_master),
                        d_row=_i32(b.d_row), d_col=_i32(b.d_col), d_val=_f64(b.d_val))
            keep.append(arrs)
            d = descs[k]
            d.m, d.n, d.d_nnz = int(b.m), int(b.n), int(len(arrs["d_val"]))
            for name, arr in arrs.items():
                ptr_t = c_ip if arr.dtype == np.int32 else c_dp
                setattr(d, name, arr.ctypes.data_as(ptr_t))
        o = Options()
        lib.dwgpu_options_init(C.byref(o))
        o.device = device
        for key, v in opts.items():
            if not hasattr(o, key):
                raise TypeError(f"unknown engine option {key!r}")
            setattr(o, key, v)
        h = C.c_void_p()
        _check(lib.dwgpu_create(self.K, self.L, descs, C.byref(o), C.byref(h)), "dwgpu_create")
        self._h = h
        self._lib = lib
        KL = max(self.K * self.L, 1)
        self.obj = np.zeros(self.K)
        self.cost = np.zeros(self.K)
        self.status = np.zeros(self.K, np.int32)
        self.len = np.zeros(self.K, np.int32)
        self.ind = np.zeros(KL, np.int32)
        self.val = np.zeros(KL)
        self.x = np.zeros(max(int(self.xoff[-1]), 1))
        self._out = Proposals(self.obj.ctypes.data_as(c_dp), self.cost.ctypes.data_as(c_dp),
                              self.status.ctypes.data_as(c_ip), self.len.ctypes.data_as(c_ip),
                              self.ind.ctypes.data_as(c_ip), self.val.ctypes.data_as(c_dp),
                              self.x.ctypes.data_as(c_dp) if fetch_x else None)

    # ---- rounds (host buffers; arrays self.obj/... are overwritten in place) -------------
    def initial_solve(self):
        _check(self._lib.dwgpu_initial_solve(self._h, C.byref(self._out)), "dwgpu_initial_solve")

    def iterate_arrays(self, pi, phase_one: bool):
        pi = _f64(pi)
        _check(self._lib.dwgpu_iterate(self._h, pi.ctypes.data_as(c_dp), int(bool(phase_one)), C.byref(self._out)),
               "dwgpu_iterate")

    # ---- rounds, packed record (engine-owned pinned memory; views valid until the next call) ----
    def _packed_views(self, pk: "Packed") -> dict:
        K, nnz = self.K, int(pk.nnz)
        as_np = np.ctypeslib.as_array
        return dict(obj=as_np(pk.obj, (K,)), cost=as_np(pk.cost, (K,)), status=as_np(pk.status, (K,)),
                    colptr=as_np(pk.colptr, (K + 1,)),
                    ind=as_np(pk.ind, (nnz,)) if nnz else np.zeros(0, np.int32),
                    val=as_np(pk.val, (nnz,)) if nnz else np.zeros(0), nnz=nnz)

    def _mailbox_views(self, mb: "Mailboxes") -> dict:
        K, L = int(mb.K), int(mb.L)
        as_np = np.ctypeslib.as_array
        return dict(obj=as_np(mb.obj, (K,)), cost=as_np(mb.cost, (K,)), status=as_np(mb.status, (K,)), len=as_np(mb.len, (K,)),
                    ind=as_np(mb.ind, (K, L)) if L else np.zeros((K, 0), np.int32),
                    val=as_np(mb.val, (K, L)) if L else np.zeros((K, 0)))

    def initial_solve_mailboxes(self) -> dict:
        """dwgpu_initial_solve_mailboxes: views of the engine's pinned host mailboxes (the kernels store into them)."""
        mb = Mailboxes()
        _check(self._lib.dwgpu_initial_solve_mailboxes(self._h, C.byref(mb)), "dwgpu_initial_solve_mailboxes")
        return self._mailbox_views(mb)

    def iterate_mailboxes(self, pi, phase_one: bool) -> dict:
        pi = _f64(pi)
        mb = Mailboxes()
        _check(self._lib.dwgpu_iterate_mailboxes(self._h, pi.ctypes.data_as(c_dp), int(bool(phase_one)), C.byref(mb)),
               "dwgpu_iterate_mailboxes")
        return self._mailbox_views(mb)

    def initial_solve_packed(self) -> dict:
        pk = Packed()
        _check(self._lib.dwgpu_initial_solve_packed(self._h, C.byref(pk)), "dwgpu_initial_solve_packed")
        return self._packed_views(pk)

    def iterate_packed(self, pi, phase_one: bool, r=None) -> dict:
        pi = _f64(pi)
        rr = _f64(r) if r is not None else None
        pk = Packed()
        _check(self._lib.dwgpu_iterate_packed(self._h, pi.ctypes.data_as(c_dp), int(bool(phase_one)),
                                              rr.ctypes.data_as(c_dp) if rr is not None else None, C.byref(pk)),
               "dwgpu_iterate_packed")
        return self._packed_views(pk)

    def pack_gathered(self, K_total: int, ptrs: dict, stream: int = 0, r_dev_ptr: int = 0) -> dict:
        """dwgpu_pack_gathered: ``ptrs`` maps obj/cost/status/len/ind/val to device addresses of arrays
        holding K_total gathered proposals; returns the same views as iterate_packed()."""
        p = Proposals()
        for name in ("obj", "cost", "val"):
            setattr(p, name, C.cast(C.c_void_p(ptrs[name]), c_dp))
        for name in ("status", "len", "ind"):
            setattr(p, name, C.cast(C.c_void_p(ptrs[name]), c_ip))
        pk = Packed()
        _check(self._lib.dwgpu_pack_gathered(self._h, K_total, C.byref(p), C.c_void_p(r_dev_ptr) if r_dev_ptr else None,
                                             C.c_void_p(stream) if stream else None, C.byref(pk)), "dwgpu_pack_gathered")
        K, nnz = K_total, int(pk.nnz)
        as_np = np.ctypeslib.as_array
        return dict(obj=as_np(pk.obj, (K,)), cost=as_np(pk.cost, (K,)), status=as_np(pk.status, (K,)),
                    colptr=as_np(pk.colptr, (K + 1,)), ind=as_np(pk.ind, (nnz,)) if nnz else np.zeros(0, np.int32),
                    val=as_np(pk.val, (nnz,)) if nnz else np.zeros(0), nnz=nnz)

    def device_record(self):
        """(device address, bytes) of the contiguous result record of this engine (one all-gather per round)."""
        p, n = C.c_void_p(), C.c_size_t()
        _check(self._lib.dwgpu_device_record(self._h, C.byref(p), C.byref(n)), "dwgpu_device_record")
        return p.value, n.value

    def pack_gathered_records(self, world: int, records_ptr: int, stream: int = 0, r_dev_ptr: int = 0) -> dict:
        pk = Packed()
        _check(self._lib.dwgpu_pack_gathered_records(self._h, world, C.c_void_p(records_ptr),
                                                     C.c_void_p(r_dev_ptr) if r_dev_ptr else None,
                                                     C.c_void_p(stream) if stream else None, C.byref(pk)),
               "dwgpu_pack_gathered_records")
        K, nnz = world * self.K, int(pk.nnz)
        as_np = np.ctypeslib.as_array
        return dict(obj=as_np(pk.obj, (K,)), cost=as_np(pk.cost, (K,)), status=as_np(pk.status, (K,)),
                    colptr=as_np(pk.colptr, (K + 1,)), ind=as_np(pk.ind, (nnz,)) if nnz else np.zeros(0, np.int32),
                    val=as_np(pk.val, (nnz,)) if nnz else np.zeros(0), nnz=nnz)

    def history_push(self, blocks) -> np.ndarray:
        """keep the current x_k of the listed blocks on the device; returns their history ids"""
        b = _i32(blocks)
        ids = np.zeros(len(b), np.int64)
        _check(self._lib.dwgpu_history_push(self._h, len(b), b.ctypes.data_as(c_ip),
                                            ids.ctypes.data_as(C.POINTER(C.c_longlong))), "dwgpu_history_push")
        return ids

    def history_combine(self, ids, weights) -> np.ndarray:
        """x (block-major) = sum of weights[t] * kept copy ids[t]"""
        ids = np.ascontiguousarray(ids, np.int64)
        w = _f64(weights)
        x = np.zeros(max(int(self.xoff[-1]), 1))
        _check(self._lib.dwgpu_history_combine(self._h, len(ids), ids.ctypes.data_as(C.POINTER(C.c_longlong)),
                                               w.ctypes.data_as(c_dp), x.ctypes.data_as(c_dp)), "dwgpu_history_combine")
        return x

    # ---- multi-GPU rounds from the C layer (include/dwgpu.h: dwgpu_comm_*, dwgpu_uplink_*, dwgpu_round_comm) ----
    @staticmethod
    def comm_unique_id() -> bytes:
        lib = load_library()
        buf = C.create_string_buffer(128)
        _check(lib.dwgpu_comm_unique_id(buf), "dwgpu_comm_unique_id")
        return buf.raw

    def comm_init(self, unique_id: bytes, world: int, rank: int):
        _check(self._lib.dwgpu_comm_init(self._h, C.c_char_p(unique_id), world, rank), "dwgpu_comm_init")

    def uplink_create(self, K_total: int):
        """rank 0: returns (64-byte CUDA IPC handle, Proposals struct of device arrays over all K_total blocks)"""
        h = C.create_string_buffer(64)
        p = Proposals()
        _check(self._lib.dwgpu_uplink_create(self._h, K_total, h, C.byref(p)), "dwgpu_uplink_create")
        return h.raw, p

    def uplink_attach(self, handle: Optional[bytes], K_total: int, k_first: int):
        _check(self._lib.dwgpu_uplink_attach(self._h, C.c_char_p(handle) if handle is not None else None, K_total, k_first),
               "dwgpu_uplink_attach")

    def round_comm(self, d_pi_ptr: int, phase_one: bool, initial: bool, stream: int = 0):
        _check(self._lib.dwgpu_round_comm(self._h, C.c_void_p(d_pi_ptr) if d_pi_ptr else None, 1 if phase_one else 0,
                                          1 if initial else 0, C.c_void_p(stream) if stream else None), "dwgpu_round_comm")

    def pack_gathered_struct(self, K_total: int, props: "Proposals", stream: int = 0, r_dev_ptr: int = 0) -> dict:
        """dwgpu_pack_gathered on a Proposals struct of device arrays (what uplink_create returned)"""
        pk = Packed()
        _check(self._lib.dwgpu_pack_gathered(self._h, K_total, C.byref(props), C.c_void_p(r_dev_ptr) if r_dev_ptr else None,
                                             C.c_void_p(stream) if stream else None, C.byref(pk)), "dwgpu_pack_gathered")
        n = int(pk.nnz)
        return dict(nnz=n, status=np.ctypeslib.as_array(pk.status, shape=(K_total,)),
                    obj=np.ctypeslib.as_array(pk.obj, shape=(K_total,)),
                    colptr=np.ctypeslib.as_array(pk.colptr, shape=(K_total + 1,)))

    def iterate_multi(self, Pi, phase_one: bool, keep_history: bool = True):
        """dwgpu_iterate_multi: Pi is P x L; returns (list of P packed-record dicts, P x K history ids or None)"""
        Pi = np.ascontiguousarray(Pi, np.float64).reshape(-1, max(self.L, 1))
        P = Pi.shape[0]
        recs = (Packed * P)()
        ids = np.zeros((P, self.K), np.int64) if keep_history else None
        _check(self._lib.dwgpu_iterate_multi(self._h, Pi.ctypes.data_as(c_dp), P, 1 if phase_one else 0, recs,
                                             ids.ctypes.data_as(C.POINTER(C.c_longlong)) if keep_history else None),
               "dwgpu_iterate_multi")
        return [self._packed_views(recs[p]) for p in range(P)], ids

    def dense_pricing_blocks(self) -> int:
        return int(self._lib.dwgpu_dense_pricing_blocks(self._h))

    def iterate_device(self, d_pi_ptr: int, phase_one: bool, stream: int = 0):
        _check(self._lib.dwgpu_iterate_device(self._h, C.c_void_p(d_pi_ptr), int(bool(phase_one)),
                                              C.c_void_p(stream) if stream else None), "dwgpu_iterate_device")

    def reset(self, stream: int = 0):
        _check(self._lib.dwgpu_reset(self._h, C.c_void_p(stream) if stream else None), "dwgpu_reset")

    def initial_solve_device(self, stream: int = 0):
        _check(self._lib.dwgpu_initial_solve_device(self._h, C.c_void_p(stream) if stream else None),
               "dwgpu_initial_solve_device")

    def fetch_results(self):
        _check(self._lib.dwgpu_fetch_results(self._h, C.byref(self._out)), "dwgpu_fetch_results")

    def device_results(self) -> Proposals:
        p = Proposals()
        _check(self._lib.dwgpu_device_results(self._h, C.byref(p)), "dwgpu_device_results")
        return p

    def sync(self):
        _check(self._lib.dwgpu_sync(self._h), "dwgpu_sync")

    # ---- proposal dicts for the test-harness DW loop ----------------------------------------
    def _collect(self) -> List[dict]:
        out = []
        for k in range(self.K):
            ln = int(self.len[k])
            s = k * self.L
            out.append(dict(status=int(self.status[k]), obj=float(self.obj[k]), cost=float(self.cost[k]), len=ln,
                            ind=self.ind[s:s + ln].copy(), val=self.val[s:s + ln].copy(),
                            x=self.x[self.xoff[k]:self.xoff[k + 1]].copy()))
        return out

    def initial(self):
        self.initial_solve()
        return self._collect()

    def iterate(self, pi, phase_one: bool):
        self.iterate_arrays(pi, phase_one)
        return self._collect()

    # ---- diagnostics --------------------------------------------------------------------------
    def counters(self) -> dict:
        c = Counters()
        _check(self._lib.dwgpu_get_counters(self._h, C.byref(c)), "dwgpu_get_counters")
        return dict(pivots=c.pivots, flips=c.flips, rebuilds=c.rebuilds, phase1_iter=c.phase1_iter, solves=c.solves,
                    rows_swept=c.rows_swept, cells_swept=c.cells_swept, price_cells=c.price_cells,
                    mailbox_bytes=c.mailbox_bytes)

    def block_iterations(self) -> np.ndarray:
        it = np.zeros(self.K, np.int64)
        _check(self._lib.dwgpu_get_block_iterations(self._h, it.ctypes.data_as(C.POINTER(C.c_longlong))),
               "dwgpu_get_block_iterations")
        return it

    def block_paths(self) -> np.ndarray:
        p = np.zeros(self.K, np.int32)
        _check(self._lib.dwgpu_get_block_paths(self._h, p.ctypes.data_as(c_ip)), "dwgpu_get_block_paths")
        return p

    def last_kernel_ms(self):
        """(shared-memory path ms, global-memory path ms) of the last round's kernels."""
        ms = (C.c_float * 2)()
        _check(self._lib.dwgpu_last_kernel_ms(self._h, ms), "dwgpu_last_kernel_ms")
        return float(ms[0]), float(ms[1])

    def last_launch_count(self) -> int:
        return int(self._lib.dwgpu_last_launch_count(self._h))

    def sparse_stats(self, reset: bool = False) -> dict:
        """Path-5 diagnostics (packed shared-memory tableau): hand-overs, heap re-packs, probe outcomes, heap sizes."""
        out = (C.c_ulonglong * 10)()
        fn = self._lib.dwgpu_debug_sparse_stats
        fn.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.c_int]
        if fn(self._h, out, 1 if reset else 0) != 0:
            raise RuntimeError("dwgpu_debug_sparse_stats failed")
        names = ["to_large_heap", "to_dense_heap", "to_dense_residual", "repacks", "batched_sweeps", "union_checks",
                 "probe_finished", "probe_passed_on", "heap_small", "heap_large"]
        return {k: int(v) for k, v in zip(names, out)}

    def close(self):
        if getattr(self, "_h", None):
            self._lib.dwgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
