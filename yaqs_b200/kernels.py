"""ctypes binding of libyaqs_b200.so (include/yaqs_b200.h) -- torch tensors only lend ``data_ptr()``.

There is NO CPU fallback: constructing :class:`CudaKernels` without the built extension or without a CUDA device
raises.  (A numpy twin of these entry points exists under ``oracle/np_kernels.py``; it is test infrastructure and is
never imported from this package.)
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import torch

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libyaqs_b200.so")

# op codes of yb_gemm (csrc/yb_common.cuh)
OP_RCONTRACT, OP_LCONTRACT, OP_ENVR3, OP_ENVL1, OP_ENVL3, OP_MERGE, OP_ABSORB_R, OP_ABSORB_L, OP_GRAM, OP_EVEC, OP_ROTATE, OP_ROTATE_TOP = range(12)
# yb_svd_prepare modes
SVD_THETA, SVD_ROWS, SVD_COLS = 0, 1, 2
SVD_ORIENT_ROWS, SVD_ORIENT_COLS = 1 << 4, 2 << 4      # force W = M / W = M^H (Gram path)
# truncation modes (core/linalg/svd_utils.py) and singular-value distribution (decompositions.py)
TRUNC_MODES = {"discarded_weight": 0, "relative": 1, "hard_cutoff": 2, "relative_discarded_weight": 3}
DIST = {"left": 0, "right": 1, "sqrt": 2}
LANCZOS_PITCH = 32          # alpha/beta/coef rows are 32 wide (LZ_KMAX in csrc/yb_ops.cu)
SVD_META = 8


class YbExtent(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("stride", ctypes.c_int32)]


@dataclass
class Ext:
    """Per-trajectory int32 extent: value of trajectory b is ``tensor.view(-1)[offset + b * stride]``."""
    tensor: torch.Tensor
    offset: int
    stride: int

    def values(self):
        flat = self.tensor.view(-1)
        n = (flat.numel() - self.offset + self.stride - 1) // self.stride
        return flat[self.offset::self.stride][:n]


@dataclass
class Buf:
    """Per-trajectory complex128 operand: trajectory b starts at ``tensor.view(-1)[offset + b * stride]``."""
    tensor: torch.Tensor
    offset: int
    stride: int

    def at(self, extra: int) -> "Buf":
        return Buf(self.tensor, self.offset + extra, self.stride)


class KernelError(RuntimeError):
    pass


class YbWorkspace(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in ("state", "environments", "theta", "matvec", "krylov", "svd", "heev", "small", "total")]


def workspace_bytes(length: int, nbatch: int, cap, mpo_bond, order: int, two_envs: bool, gram: bool = True, dphys: int = 2) -> dict:
    """yb_workspace_bytes (include/yaqs_b200.h): device bytes per component for a batch -- host arithmetic of the library, callable
    without a device.  ``Simulator`` sizes its waves from it."""
    lib = load_library()
    capa = (ctypes.c_int32 * (length + 1))(*[int(c) for c in cap])
    da = (ctypes.c_int32 * (length + 1))(*[int(c) for c in mpo_bond])
    out = YbWorkspace()
    if lib.yb_workspace_bytes(int(length), int(nbatch), int(dphys), capa, da, int(order), int(bool(two_envs)), int(bool(gram)), ctypes.byref(out)) != 0:
        raise KernelError(lib.yb_last_error().decode())
    return {n: int(getattr(out, n)) for n, _ in YbWorkspace._fields_}


def load_library() -> ctypes.CDLL:
    if not os.path.exists(_LIB_PATH):
        raise KernelError(f"{_LIB_PATH} is missing: build it with `python -m yaqs_b200.build` (there is no CPU fallback)")
    lib = ctypes.CDLL(_LIB_PATH)
    lib.yb_last_error.restype = ctypes.c_char_p
    lib.yb_launch_count.restype = ctypes.c_int64
    return lib


EXPORTED_SYMBOLS = [
    "yb_last_error", "yb_version", "yb_launch_count", "yb_set_flop_counter", "yb_set_option", "yb_workspace_bytes", "yb_gemm", "yb_rcontract_mpo", "yb_wapply", "yb_env_identity", "yb_lanczos_init",
    "yb_lanczos_step", "yb_lanczos_combine", "yb_svd_prepare", "yb_svd_round", "yb_svd_check", "yb_svd_finalize", "yb_svd_ambiguous", "yb_heev", "yb_heev_rows", "yb_heev_rows_top", "yb_svd_finalize_top", "yb_cholqr_factor", "yb_cholqr_refine", "yb_cholqr_finalize", "yb_svd_select", "yb_svd_gather",
    "yb_jump_test", "yb_jump_select", "yb_observables", "yb_diagnostics", "yb_mask_andnot", "yb_graph_begin", "yb_graph_end", "yb_graph_launch", "yb_graph_destroy", "yb_unpack_vh", "yb_copy_packed", "yb_apply_local_op", "yb_mpo_site_apply", "yb_norm2", "yb_scale", "yb_local_rdm", "yb_philox_uniform",
]


class CudaKernels:
    """The product backend: every method enqueues one kernel of libyaqs_b200.so on torch's current stream."""

    name = "cuda"

    def __init__(self, device: torch.device | str | None = None):
        if not torch.cuda.is_available():
            raise KernelError("yaqs_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = load_library()
        self.device = torch.device(device if device is not None else "cuda:0")
        if self.device.type != "cuda":
            raise KernelError("yaqs_b200 kernels run on CUDA devices only")
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
        self._timed: dict[str, list] | None = None
        self._detail: dict[tuple, list] | None = None
        self._flops = None
        self.phase = ""              # label set by the engine (tdvp / dissipation / jumps / observables): second key of the detailed timing

    # -- instrumentation (bench.py): CUDA events around every launch of a kernel family + on-device flop counter --------
    def start_timing(self) -> None:
        self._timed = {}
        self._detail = {}
        self._flops = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.lib.yb_set_flop_counter(ctypes.c_void_p(self._flops.data_ptr()))

    def stop_timing(self) -> dict:
        """Returns {family: {"launches": n, "ms": total}} and "gemm_flops" (algorithmic, counted on the device)."""
        torch.cuda.synchronize(self.device)
        self.lib.yb_set_flop_counter(ctypes.c_void_p(0))
        out = {k: {"launches": len(v), "ms": float(sum(a.elapsed_time(b) for a, b in v))} for k, v in (self._timed or {}).items()}
        out["gemm_flops"] = int(self._flops.item()) if self._flops is not None else 0
        # (family, sub-key, phase) -> launches / ms: where inside the step a family's time goes (bench.py `kernel_breakdown`)
        out["detail"] = {k: {"launches": len(v), "ms": float(sum(a.elapsed_time(b) for a, b in v))} for k, v in (self._detail or {}).items()}
        self._timed, self._detail, self._flops = None, None, None
        return out

    def set_option(self, name: str, value: int) -> None:
        """Diagnostic switch of the library (include/yaqs_b200.h: yb_set_option)."""
        self._check(self.lib.yb_set_option(name.encode(), int(value)))

    def _timed_call(self, family: str, fn, sub: str = "") -> None:
        if self._timed is None:
            fn()
            return
        st = torch.cuda.current_stream(self.device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        fn()
        e1.record(st)
        self._timed.setdefault(family, []).append((e0, e1))
        self._detail.setdefault((family, sub, self.phase), []).append((e0, e1))

    # -- helpers ------------------------------------------------------------------------------------------------------
    def _stream(self):
        # torch's CURRENT stream of the calling thread, as a raw cudaStream_t (the private accessor skips the Stream object: this runs
        # once per launch, and small-chi workloads are bounded by the host cost of a launch)
        if self._raw_stream is not None:
            return ctypes.c_void_p(self._raw_stream(self._dev_index))
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _cptr(buf: Buf | None):
        if buf is None:
            return ctypes.c_void_p(0)
        if buf.tensor.dtype != torch.complex128:
            raise KernelError(f"complex128 operand expected, got {buf.tensor.dtype}")
        return ctypes.c_void_p(buf.tensor.data_ptr() + 16 * buf.offset)

    @staticmethod
    def _ptr(t: torch.Tensor | None, offset_bytes: int = 0):
        if t is None:
            return ctypes.c_void_p(0)
        return ctypes.c_void_p(t.data_ptr() + offset_bytes)

    @staticmethod
    def _ext(e: Ext | None):
        if e is None:
            return YbExtent(ctypes.c_void_p(0), 0)
        return YbExtent(ctypes.c_void_p(e.tensor.data_ptr() + 4 * e.offset), e.stride)

    def _check(self, rc: int):
        if rc != 0:
            raise KernelError(self.lib.yb_last_error().decode())

    def launch_count(self, reset: bool = False) -> int:
        return int(self.lib.yb_launch_count(1 if reset else 0))

    def sync_int(self, t: torch.Tensor) -> int:
        """Read a device int (blocks the host until the stream reaches it)."""
        return int(t.item())

    # -- contractions -------------------------------------------------------------------------------------------------
    def gemm(self, op, nbatch, ngroups, mask, e0, e1, e2, dphys, d_left, d_right, a: Buf, b: Buf, c: Buf, conj_b, max_m, max_n):
        self._timed_call("gemm", lambda: self._check(self.lib.yb_gemm(
            op, nbatch, ngroups, self._ptr(mask), self._ext(e0), self._ext(e1), self._ext(e2), dphys, d_left, d_right,
            self._cptr(a), ctypes.c_int64(a.stride), self._cptr(b), ctypes.c_int64(b.stride), self._cptr(c), ctypes.c_int64(c.stride),
            int(conj_b), int(max_m), int(max_n), self._stream())), f"op{op}:{max_m}x{max_n}" if self._timed is not None else "")

    def rcontract_mpo(self, nbatch, mask, el, er, dphys, d_left, d_right, x: Buf, renv: Buf, wt: torch.Tensor, out: Buf, max_l, max_r):
        if wt.dtype != torch.complex128:
            raise KernelError(f"rcontract_mpo: complex128 MPO tensor expected, got {wt.dtype}")
        self._timed_call("gemm", lambda: self._check(self.lib.yb_rcontract_mpo(
            nbatch, self._ptr(mask), self._ext(el), self._ext(er), int(dphys), int(d_left), int(d_right), self._cptr(x), ctypes.c_int64(x.stride),
            self._cptr(renv), ctypes.c_int64(renv.stride), self._ptr(wt), self._cptr(out), ctypes.c_int64(out.stride), int(max_l), int(max_r),
            self._stream())), f"rw:{dphys * max_l}x{d_right * max_r}" if self._timed is not None else "")

    def wapply(self, nbatch, mask, el, er, nx, nu, ny, nv, wt: torch.Tensor, src: Buf, dst: Buf, max_lr):
        if wt.dtype != torch.complex128:
            raise KernelError(f"wapply: complex128 MPO tensor expected, got {wt.dtype}")
        self._timed_call("wapply", lambda: self._check(self.lib.yb_wapply(
            nbatch, self._ptr(mask), self._ext(el), self._ext(er), nx, nu, ny, nv, self._ptr(wt), self._cptr(src),
            ctypes.c_int64(src.stride), self._cptr(dst), ctypes.c_int64(dst.stride), int(max_lr), self._stream())))

    def env_identity(self, nbatch, mask, echi, dmpo, env: Buf, max_chi):
        self._timed_call("local", lambda: self._check(self.lib.yb_env_identity(nbatch, self._ptr(mask), self._ext(echi), dmpo, self._cptr(env), ctypes.c_int64(env.stride),
                                             int(max_chi), self._stream())))

    # -- Lanczos ------------------------------------------------------------------------------------------------------
    def lanczos_init(self, nbatch, mask, el, er, dphys, x: Buf, basis: Buf, stride_vec, nrm, active, kdone, n_zero=None):
        self._timed_call("lanczos_misc", lambda: self._check(self.lib.yb_lanczos_init(
            nbatch, self._ptr(mask), self._ext(el), self._ext(er), dphys, self._cptr(x), ctypes.c_int64(x.stride), self._cptr(basis),
            ctypes.c_int64(basis.stride), ctypes.c_int64(stride_vec), self._ptr(nrm), self._ptr(active), self._ptr(kdone), self._ptr(n_zero),
            self._stream())))

    def lanczos_step(self, nbatch, j, kmax, dt, tol, el, er, dphys, basis: Buf, stride_vec, alpha, beta, nrm, active, kdone, coef, n_active, n_zero=None):
        self._timed_call("lanczos_step", lambda: self._check(self.lib.yb_lanczos_step(
            nbatch, j, kmax, ctypes.c_double(dt), ctypes.c_double(tol), self._ext(el), self._ext(er), dphys, self._cptr(basis),
            ctypes.c_int64(basis.stride), ctypes.c_int64(stride_vec), self._ptr(alpha), self._ptr(beta), self._ptr(nrm), self._ptr(active),
            self._ptr(kdone), self._ptr(coef), self._ptr(n_active), self._ptr(n_zero), self._stream())))

    def lanczos_combine(self, nbatch, mask, el, er, dphys, basis: Buf, stride_vec, coef, kmax, kdone, out: Buf):
        self._timed_call("lanczos_misc", lambda: self._check(self.lib.yb_lanczos_combine(
            nbatch, self._ptr(mask), self._ext(el), self._ext(er), dphys, self._cptr(basis), ctypes.c_int64(basis.stride),
            ctypes.c_int64(stride_vec), self._ptr(coef), kmax, self._ptr(kdone), self._cptr(out), ctypes.c_int64(out.stride), self._stream())))

    # -- SVD ----------------------------------------------------------------------------------------------------------
    def svd_prepare(self, nbatch, mask, mode, el, er, dphys, src: Buf, w: Buf, q: Buf, meta, rot_count, n_zero=None, status_zero=None):
        self._timed_call("svd_misc", lambda: self._check(self.lib.yb_svd_prepare(
            nbatch, self._ptr(mask), mode, self._ext(el), self._ext(er), dphys, self._cptr(src), ctypes.c_int64(src.stride), self._cptr(w),
            ctypes.c_int64(w.stride), self._cptr(q), ctypes.c_int64(q.stride), self._ptr(meta), self._ptr(rot_count), self._ptr(n_zero), self._ptr(status_zero),
            self._stream())))

    def svd_round(self, nbatch, mask, meta, block_rows, nblocks, rnd, full_panel, inner_sweeps, w: Buf, q: Buf, rot_count, max_len, max_k, n_zero=None):
        self._timed_call("svd_round", lambda: self._check(self.lib.yb_svd_round(
            nbatch, self._ptr(mask), self._ptr(meta), block_rows, nblocks, rnd, int(full_panel), inner_sweeps, self._cptr(w),
            ctypes.c_int64(w.stride), self._cptr(q), ctypes.c_int64(q.stride), self._ptr(rot_count), int(max_len), int(max_k), self._ptr(n_zero),
            self._stream())))

    def svd_check(self, nbatch, mask, meta, rot_count, n_pending):
        self._timed_call("svd_misc", lambda: self._check(self.lib.yb_svd_check(nbatch, self._ptr(mask), self._ptr(meta), self._ptr(rot_count), self._ptr(n_pending), self._stream())))

    def svd_finalize(self, nbatch, mask, meta, trunc_mode, threshold, max_bond, min_keep, dist, w: Buf, q: Buf, sval, stride_s, keep: Ext,
                     uout: Buf, vout: Buf, status, max_k):
        self._timed_call("svd_misc", lambda: self._check(self.lib.yb_svd_finalize(
            nbatch, self._ptr(mask), self._ptr(meta), trunc_mode, ctypes.c_double(threshold), int(max_bond), int(min_keep), dist,
            self._cptr(w), ctypes.c_int64(w.stride), self._cptr(q), ctypes.c_int64(q.stride), self._ptr(sval), ctypes.c_int64(stride_s),
            self._ptr(keep.tensor, 4 * keep.offset), keep.stride, self._cptr(uout), ctypes.c_int64(uout.stride), self._cptr(vout),
            ctypes.c_int64(vout.stride), self._ptr(status), int(max_k), self._stream())))

    def svd_ambiguous(self, nbatch, mask, meta, sval, stride_s, trunc_mode, threshold, two_level, tau_rel, noise_rel, status, amb, n_amb):
        self._timed_call("svd_misc", lambda: self._check(self.lib.yb_svd_ambiguous(nbatch, self._ptr(mask), self._ptr(meta), self._ptr(sval), ctypes.c_int64(stride_s), int(trunc_mode),
                                              ctypes.c_double(threshold), int(bool(two_level)), ctypes.c_double(tau_rel), ctypes.c_double(noise_rel),
                                              self._ptr(status), self._ptr(amb), self._ptr(n_amb), self._stream())))

    def heev(self, nbatch, mask, ek: Ext, g: Buf, qh: Buf, zt: Buf, zr: torch.Tensor, stride_zr, d, e, tau: torch.Tensor, stride_vec, meta, status,
             max_k):
        self._timed_call("heev", lambda: self._check(self.lib.yb_heev(
            nbatch, self._ptr(mask), self._ext(ek), self._cptr(g), ctypes.c_int64(g.stride), self._cptr(qh), ctypes.c_int64(qh.stride),
            self._cptr(zt), ctypes.c_int64(zt.stride), self._ptr(zr), ctypes.c_int64(stride_zr), self._ptr(d), self._ptr(e), self._ptr(tau),
            ctypes.c_int64(stride_vec), self._ptr(meta), self._ptr(status), int(max_k), self._stream())), f"k{max_k}")

    def heev_rows(self, nbatch, mask, ek: Ext, g: Buf, eh: Buf, zr: torch.Tensor, stride_zr, d, e, tau: torch.Tensor, stride_vec, meta, status, max_k):
        self._timed_call("heev", lambda: self._check(self.lib.yb_heev_rows(
            nbatch, self._ptr(mask), self._ext(ek), self._cptr(g), ctypes.c_int64(g.stride), self._cptr(eh), ctypes.c_int64(eh.stride),
            self._ptr(zr), ctypes.c_int64(stride_zr), self._ptr(d), self._ptr(e), self._ptr(tau), ctypes.c_int64(stride_vec), self._ptr(meta),
            self._ptr(status), int(max_k), self._stream())), f"k{max_k}")

    def heev_rows_top(self, nbatch, mask, ek: Ext, g: Buf, eh: Buf, zr: torch.Tensor, stride_zr, d, e, tau: torch.Tensor, stride_vec, meta, status, max_k,
                      top, order, dsorted, ntop):
        self._timed_call("heev", lambda: self._check(self.lib.yb_heev_rows_top(
            nbatch, self._ptr(mask), self._ext(ek), self._cptr(g), ctypes.c_int64(g.stride), self._cptr(eh), ctypes.c_int64(eh.stride),
            self._ptr(zr), ctypes.c_int64(stride_zr), self._ptr(d), self._ptr(e), self._ptr(tau), ctypes.c_int64(stride_vec), self._ptr(meta),
            self._ptr(status), int(max_k), int(top), self._ptr(order), self._ptr(dsorted), self._ptr(ntop), self._stream())), f"k{max_k}top{top}")

    def svd_finalize_top(self, nbatch, mask, meta, trunc_mode, threshold, max_bond, min_keep, dist, w: Buf, q: Buf, sval, stride_s, keep: Ext,
                         uout: Buf, vout: Buf, status, max_k, lam, stride_lam, top):
        self._timed_call("svd_misc", lambda: self._check(self.lib.yb_svd_finalize_top(
            nbatch, self._ptr(mask), self._ptr(meta), trunc_mode, ctypes.c_double(threshold), int(max_bond), int(min_keep), dist,
            self._cptr(w), ctypes.c_int64(w.stride), self._cptr(q), ctypes.c_int64(q.stride), self._ptr(sval), ctypes.c_int64(stride_s),
            self._ptr(keep.tensor, 4 * keep.offset), keep.stride, self._cptr(uout), ctypes.c_int64(uout.stride), self._cptr(vout),
            ctypes.c_int64(vout.stride), self._ptr(status), int(max_k), self._ptr(lam), ctypes.c_int64(stride_lam), int(top), self._stream())))

    # -- truncation-free gauge shifts by Cholesky-QR (csrc/yb_cholqr.cuh) ----------------------------------------------
    def cholqr_factor(self, nbatch, mask, meta, g: Buf, linv: Buf, threshold, flag, n_flag, max_k):
        self._timed_call("cholqr", lambda: self._check(self.lib.yb_cholqr_factor(
            nbatch, self._ptr(mask), self._ptr(meta), self._cptr(g), ctypes.c_int64(g.stride), self._cptr(linv), ctypes.c_int64(linv.stride),
            ctypes.c_double(threshold), self._ptr(flag), self._ptr(n_flag), int(max_k), self._stream())), f"k{max_k}")

    def cholqr_refine(self, nbatch, mask, meta, g2: Buf, x: Buf, tol, flag, n_flag):
        self._timed_call("svd_misc", lambda: self._check(self.lib.yb_cholqr_refine(
            nbatch, self._ptr(mask), self._ptr(meta), self._cptr(g2), ctypes.c_int64(g2.stride), self._cptr(x), ctypes.c_int64(x.stride),
            ctypes.c_double(tol), self._ptr(flag), self._ptr(n_flag), self._stream())))

    def cholqr_finalize(self, nbatch, mask, meta, y2: Buf, l2: Buf, keep: Ext, uout: Buf, vout: Buf, flag, max_elems):
        self._timed_call("svd_misc", lambda: self._check(self.lib.yb_cholqr_finalize(
            nbatch, self._ptr(mask), self._ptr(meta), self._cptr(y2), ctypes.c_int64(y2.stride), self._cptr(l2), ctypes.c_int64(l2.stride),
            self._ptr(keep.tensor, 4 * keep.offset), keep.stride, self._cptr(uout), ctypes.c_int64(uout.stride), self._cptr(vout),
            ctypes.c_int64(vout.stride), self._ptr(flag), int(max_elems), self._stream())))

    def svd_select(self, nbatch, mask, meta, w: Buf, tau_rel, sel_n, sel_idx, max_k):
        self._timed_call("svd_misc", lambda: self._check(self.lib.yb_svd_select(nbatch, self._ptr(mask), self._ptr(meta), self._cptr(w), ctypes.c_int64(w.stride),
                                           ctypes.c_double(tau_rel), self._ptr(sel_n), self._ptr(sel_idx), int(max_k), self._stream())))

    def svd_gather(self, nbatch, mask, meta, sel_n, sel_idx, max_k, direction, w: Buf, q: Buf, ws: Buf, qs: Buf):
        self._timed_call("svd_misc", lambda: self._check(self.lib.yb_svd_gather(nbatch, self._ptr(mask), self._ptr(meta), self._ptr(sel_n), self._ptr(sel_idx), int(max_k),
                                           int(direction), self._cptr(w), ctypes.c_int64(w.stride), self._cptr(q), ctypes.c_int64(q.stride),
                                           self._cptr(ws), ctypes.c_int64(ws.stride), self._cptr(qs), ctypes.c_int64(qs.stride), self._stream())))

    # -- site-tensor utilities ----------------------------------------------------------------------------------------
    def unpack_vh(self, nbatch, mask, ek, er, dphys, src: Buf, dst: Buf, max_elems, set_ext: Ext | None = None):
        self._timed_call("pack", lambda: self._check(self.lib.yb_unpack_vh(nbatch, self._ptr(mask), self._ext(ek), self._ext(er), dphys, self._cptr(src),
                                          ctypes.c_int64(src.stride), self._cptr(dst), ctypes.c_int64(dst.stride), int(max_elems), self._ext(set_ext), self._stream())))

    def copy_packed(self, nbatch, mask, e0, e1, mult, src: Buf, dst: Buf, max_elems, set_ext: Ext | None = None):
        self._timed_call("pack", lambda: self._check(self.lib.yb_copy_packed(nbatch, self._ptr(mask), self._ext(e0), self._ext(e1), mult, self._cptr(src),
                                            ctypes.c_int64(src.stride), self._cptr(dst), ctypes.c_int64(dst.stride), int(max_elems), self._ext(set_ext), self._stream())))

    def apply_local_op(self, nbatch, mask, sel, sel_value, el, er, dphys, op: torch.Tensor, op_stride, a: Buf, max_lr):
        if op.dtype != torch.complex128:
            raise KernelError(f"apply_local_op: complex128 operator expected, got {op.dtype}")
        self._timed_call("local", lambda: self._check(self.lib.yb_apply_local_op(nbatch, self._ptr(mask), self._ptr(sel), int(sel_value), self._ext(el), self._ext(er), dphys,
                                               self._ptr(op), ctypes.c_int64(op_stride), self._cptr(a), ctypes.c_int64(a.stride), int(max_lr),
                                               self._stream())))

    def mpo_site_apply(self, nbatch, mask, el, er, dphys, d_left, d_right, w: torch.Tensor, src: Buf, dst: Buf, max_out_lr):
        if w.dtype != torch.complex128:
            raise KernelError(f"mpo_site_apply: complex128 MPO tensor expected, got {w.dtype}")
        self._timed_call("local", lambda: self._check(self.lib.yb_mpo_site_apply(nbatch, self._ptr(mask), self._ext(el), self._ext(er), dphys, int(d_left), int(d_right), self._ptr(w),
                                               self._cptr(src), ctypes.c_int64(src.stride), self._cptr(dst), ctypes.c_int64(dst.stride),
                                               int(max_out_lr), self._stream())))

    def norm2(self, nbatch, mask, el, er, dphys, a: Buf, out):
        self._timed_call("local", lambda: self._check(self.lib.yb_norm2(nbatch, self._ptr(mask), self._ext(el), self._ext(er), dphys, self._cptr(a), ctypes.c_int64(a.stride),
                                      self._ptr(out), self._stream())))

    def scale(self, nbatch, mask, el, er, dphys, scale, a: Buf, max_elems, rsqrt_mode: bool = False):
        self._timed_call("local", lambda: self._check(self.lib.yb_scale(nbatch, self._ptr(mask), self._ext(el), self._ext(er), dphys, self._ptr(scale), self._cptr(a),
                                      ctypes.c_int64(a.stride), int(max_elems), int(bool(rsqrt_mode)), self._stream())))

    # -- stochastic_process on the device -------------------------------------------------------------------------------
    def jump_test(self, nbatch, norm2, table, col, pos, philox, dp, u1, u2, jmask, njump):
        """table: [B, n] float64 rows (table mode) or None; philox: (seed, traj int64 [B], tag, block0) or None; pos: int64 [B] or None."""
        use_philox = philox is not None
        seed, traj, tag, block0 = philox if use_philox else (0, None, 0, 0)
        if not use_philox and (table.dtype != torch.float64 or table.stride(-1) != 1):
            raise KernelError("jump_test: contiguous float64 rows expected")
        stride = 0 if table is None else (table.stride(0) if table.dim() == 2 else 0)
        self._timed_call("local", lambda: self._check(self.lib.yb_jump_test(
            int(nbatch), self._ptr(norm2), self._ptr(table), ctypes.c_int64(stride), ctypes.c_int64(int(col)), self._ptr(pos), int(use_philox),
            ctypes.c_uint64(int(seed) & (2 ** 64 - 1)), self._ptr(traj), ctypes.c_uint32(int(tag)), ctypes.c_uint32(int(block0)), self._ptr(dp),
            self._ptr(u1), self._ptr(u2), self._ptr(jmask), self._ptr(njump), self._stream())))

    def jump_select(self, nbatch, jmask, u2, rdm1, rdm2, nproc, proc_site, proc_two, proc_ldl, proc_strength, dt, ops_table, nops, w, total, choice,
                    ops_sel, status):
        self._timed_call("local", lambda: self._check(self.lib.yb_jump_select(
            int(nbatch), self._ptr(jmask), self._ptr(u2), self._ptr(rdm1), self._ptr(rdm2), int(nproc), self._ptr(proc_site), self._ptr(proc_two),
            self._ptr(proc_ldl), self._ptr(proc_strength), ctypes.c_double(dt), self._ptr(ops_table), int(nops), self._ptr(w), self._ptr(total),
            self._ptr(choice), self._ptr(ops_sel), self._ptr(status), self._stream())))

    def observables(self, nbatch, nobs, col, site, two, mats, rdm1, rdm2, out, out_stride, status):
        self._timed_call("local", lambda: self._check(self.lib.yb_observables(
            int(nbatch), int(nobs), self._ptr(col), self._ptr(site), self._ptr(two), self._ptr(mats), self._ptr(rdm1), self._ptr(rdm2), self._ptr(out),
            ctypes.c_int64(int(out_stride)), self._ptr(status), self._stream())))

    def diagnostics(self, nbatch, length, dphys, bond, out, out_stride):
        self._timed_call("local", lambda: self._check(self.lib.yb_diagnostics(int(nbatch), int(length), int(dphys), self._ptr(bond), self._ptr(out),
                                                                             ctypes.c_int64(int(out_stride)), self._stream())))

    def mask_andnot(self, nbatch, mask, flag, work):
        self._timed_call("local", lambda: self._check(self.lib.yb_mask_andnot(int(nbatch), self._ptr(mask), self._ptr(flag), self._ptr(work), self._stream())))

    # -- CUDA graphs of fixed launch sequences (include/yaqs_b200.h) ----------------------------------------------------------
    supports_graphs = True

    def graph_begin(self) -> None:
        self._check(self.lib.yb_graph_begin(self._stream()))

    def graph_end(self) -> int:
        handle = ctypes.c_void_p(0)
        self._check(self.lib.yb_graph_end(self._stream(), ctypes.byref(handle)))
        return int(handle.value)

    def graph_launch(self, handle: int) -> None:
        self._check(self.lib.yb_graph_launch(ctypes.c_void_p(handle), self._stream()))

    def graph_destroy(self, handle: int) -> None:
        self.lib.yb_graph_destroy(ctypes.c_void_p(handle))

    @property
    def timing_active(self) -> bool:
        return self._timed is not None

    def on_default_stream(self) -> bool:
        """The legacy default stream cannot be captured: sequences issued there stay eager."""
        if self._raw_stream is not None:
            return int(self._raw_stream(self._dev_index)) == 0
        return int(torch.cuda.current_stream(self.device).cuda_stream) == 0

    def philox_uniform(self, seed: int, traj_index, tag: int, nblocks: int, block0: int, out):
        """out [nbatch, 2 * nblocks] float64 (contiguous rows); traj_index int64 [nbatch] on the device."""
        nb = int(traj_index.numel())
        self._check(self.lib.yb_philox_uniform(nb, ctypes.c_uint64(int(seed) & (2 ** 64 - 1)), self._ptr(traj_index), ctypes.c_uint32(tag), int(nblocks),
                                               ctypes.c_uint32(block0), self._ptr(out), ctypes.c_int64(out.stride(0)), self._stream()))

    def local_rdm(self, nbatch, mask, el, er, dphys, a: Buf, x: Buf, g):
        self._timed_call("local", lambda: self._check(self.lib.yb_local_rdm(nbatch, self._ptr(mask), self._ext(el), self._ext(er), dphys, self._cptr(a), ctypes.c_int64(a.stride),
                                          self._cptr(x), ctypes.c_int64(x.stride), self._ptr(g), self._stream())))
