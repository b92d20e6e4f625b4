"""Host loop of the B200 NMF solver: device-resident X shard, alternating factor updates.

One iteration (sklearn `_fit_coordinate_descent` loop body, SP/sklearn/decomposition/_nmf.py:491-502,
or `_fit_multiplicative_update`, _nmf.py:826-864) is, for the gene factor G (g x k) and the cell
factor C (k x n):

    K1  sweep(X^T, Cs32)     -> reduce      -> R.xht (kp x ldg fp64)       one sweep (of the X^T copy)
    K4  gram(C64)            -> R.cct (kp x kp fp64)
        [all-reduce R over the cell shards: the ONLY per-iteration collective]
    K2  update(G | R.xht, R.cct)            (MU ratio or CD sweep; fp64; writes the fp32 shadow)
    K4  gram(G64)            -> gtg
    K3  sweep(X, Gs32)       -> reduce      -> nb (kp x ldn fp64)          second sweep (of X)
        update(C | nb, gtg)                 (cells are local to the shard: no communication)
        iter_end                            (device-side stop rule, _nmf.py:504-516)

`first='C'` runs the cell half-step first (NmfClustering_initW factorises X^T, scRNA/nmf_clustering.py
:71-72, so sklearn's W is our C^T and is updated first).  The permutation stream of the CD solver is
pre-drawn on the host from RandomState(seed) in sklearn's consumption order.

Nothing here computes on the CPU: every array op is a C-ABI kernel; the host only sequences them,
polls the 64-byte control block and (MU) applies the every-10-iterations stop rule to a scalar.
"""
import math
import os

import numpy as np

from . import _lib


H_PARTS = 2   # fp16 parts of the factor in the 16-bit sweep (2: 22 significant bits, as the 3xTF32 sweep; 3: 33)


def round_up(a, m):
    return ((int(a) + m - 1) // m) * m


def pad_k(k, tensor_cores=False):
    """Padded component count: a multiple of 4 (8 above 32) for the FP32-pipe sweeps, a multiple of 16 for
    the tcgen05 sweep (the MMA's N)."""
    if k < 1 or k > _lib.MAX_K:
        raise ValueError("number of components must be in [1, %d], got %d" % (_lib.MAX_K, k))
    if tensor_cores:
        return round_up(k, 16)
    kp = round_up(k, 4)
    if kp > 32:
        kp = round_up(k, 8)
    return kp


def shard_bounds(n_total, world, rank):
    """Contiguous, balanced split of the cell axis: first `n_total % world` shards get
    one extra cell."""
    base, extra = divmod(int(n_total), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def _from_numpy(torch, a):
    """torch.from_numpy without the "array is not writable" warning: read-only views (pp_data of a pass-through
    pre-processing, remembered stage results) are only ever read."""
    if a.flags.writeable:
        return torch.from_numpy(a)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)
        return torch.from_numpy(a)


class DeviceMatrix:
    """X shard: genes x cells, row-major fp32 on the device, leading dimension padded to 32 cells
    (128 B) and zero-filled beyond the real cells."""

    def __init__(self, data, n, tf32_exact=None, f16_exact=None):
        self.tf32_exact = tf32_exact   # every element representable in TF32 (raw counts < 2048): lazily probed
        self.f16_exact = f16_exact     # every element representable in fp16 (raw counts <= 2048): lazily probed
        self._half = None              # fp16 twin (16-bit storage path), built once
        self.data = data
        self.g = int(data.shape[0])
        self.ldx = int(data.shape[1])
        self.n = int(n)
        self.ncols = round_up(self.n, 4)
        assert self.ldx % 4 == 0 and self.ncols <= self.ldx

    is_half = False

    @classmethod
    def from_host(cls, arr, kernels, device=None, band_bytes=256 << 20):
        """Upload a host genes x cells array (any float dtype) in row bands; fp64 bands are narrowed
        on the device (scrna_convert_pad_f64_f32)."""
        torch = kernels.torch
        arr = np.asarray(arr)
        if arr.ndim != 2:
            raise ValueError("expected a genes x cells matrix")
        g, n = arr.shape
        ldx = round_up(max(n, 1), 32)
        dev = device if device is not None else getattr(kernels, "device", None)
        if dev is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        out = torch.empty((g, ldx), dtype=torch.float32, device=dev)
        if arr.dtype != np.float64:
            arr = np.ascontiguousarray(arr, dtype=np.float64) if arr.dtype != np.float32 else arr
        if arr.dtype == np.float32:
            out.zero_()
            out[:, :n].copy_(_from_numpy(torch, np.ascontiguousarray(arr)), non_blocking=False)
            return cls(out, n)
        rows_per_band = max(1, int(band_bytes // max(1, n * 8)))
        for r0 in range(0, g, rows_per_band):
            r1 = min(g, r0 + rows_per_band)
            band = _from_numpy(torch, np.ascontiguousarray(arr[r0:r1])).to(dev)
            kernels.convert_pad(band, n, out[r0:r1], ldx, r1 - r0, n)
        if out.is_cuda:
            torch.cuda.current_stream().synchronize()
        return cls(out, n)

    def transposed(self, kernels, half=False):
        """Resident transposed copy (cells x genes): fp32 row-major with the leading dimension padded to 32 genes,
        or -- 16-bit storage path -- the panel-major fp16 matrix."""
        torch = kernels.torch
        if half:
            out = torch.empty(kernels.panel_elems(max(self.n, 1), self.g), dtype=torch.float16, device=self.data.device)
            if self.n > 0:
                kernels.transpose_f16(self.data, self.ldx, self.g, self.n, out)
            else:
                out.zero_()
            return HalfMatrix(out, max(self.n, 1), self.g, kernels.panel_rows(max(self.n, 1)))
        ldg = round_up(self.g, 32)
        out = torch.zeros((max(self.n, 1), ldg), dtype=torch.float32, device=self.data.device)
        if self.n > 0:
            kernels.transpose(self.data, self.ldx, self.g, self.n, out, ldg)
        return DeviceMatrix(out, self.g, self.tf32_exact, self.f16_exact)

    def half_twin(self, kernels):
        """Panel-major fp16 storage of an fp16-exact matrix, built once and kept."""
        if self._half is None:
            torch = kernels.torch
            out = torch.empty(kernels.panel_elems(self.g, max(self.n, 1)), dtype=torch.float16, device=self.data.device)
            if self.n > 0:
                kernels.narrow_f16(self.data, self.ldx, self.g, self.n, out)
            else:
                out.zero_()
            self._half = HalfMatrix(out, self.g, self.n, kernels.panel_rows(self.g))
        return self._half

    def is_tf32_exact(self, kernels):
        if self.tf32_exact is None:
            self.tf32_exact = bool(kernels.tf32_exact(self.data))
        return self.tf32_exact

    def is_f16_exact(self, kernels):
        if self.f16_exact is None:
            self.f16_exact = bool(kernels.f16_exact(self.data))
        return self.f16_exact


class HalfMatrix:
    """fp16 storage of a (rows x cols) matrix in PANELS of 256 columns: element (i, j) at
    ((j // 256) * rows_pad + i) * 256 + j % 256, rows_pad = rows rounded up to 64, padding zero
    (include/scrna_b200.h, scrna_nmf_wtx_partial_h): the rows a sweep CTA streams are contiguous in HBM."""
    is_half = True
    PANEL = 256

    def __init__(self, data, rows, cols, rows_pad):
        self.data = data
        self.g = int(rows)
        self.n = int(cols)
        self.rows_pad = int(rows_pad)
        self.ldx = self.rows_pad          # what the C ABI takes in place of a leading dimension
        self.ncols = round_up(self.n, 4)
        self.npanels = (self.n + self.PANEL - 1) // self.PANEL

    def dense(self):
        """rows x cols view of the stored values (tests)."""
        v = self.data.view(self.npanels, self.rows_pad, self.PANEL).permute(1, 0, 2)
        return v.reshape(self.rows_pad, self.npanels * self.PANEL)[: self.g, : self.n]


class TorchDistComm:
    """All-reduce plumbing over torch.distributed (NCCL on the GPUs; gloo in the CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def all_reduce(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)

    def barrier(self):
        self.dist.barrier(group=self.group)

    def symmetric_arena(self, nbytes, device):
        """(zeroed uint8 tensor of nbytes in torch symmetric memory, [base address of every rank's copy in THIS rank's
        address space]) -- or None when the backend has no peer-mapped memory (gloo, no P2P): the caller then keeps
        the NCCL all-reduce path.  Collective: every rank of the group calls it with the same size."""
        if os.environ.get("SCRNA_B200_P2P", "1") == "0" or device.type != "cuda":
            return None
        try:
            if self.dist.get_backend(self.group) != "nccl":
                return None
            import torch
            import torch.distributed._symmetric_memory as symm
            group = self.group if self.group is not None else self.dist.group.WORLD
            arena = symm.empty(int(nbytes), dtype=torch.uint8, device=device)
            hdl = symm.rendezvous(arena, group)
            arena.zero_()
            torch.cuda.synchronize()
            self.dist.barrier(group=self.group)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            if len(ptrs) != self.world:
                return None
            self._symm_handles = getattr(self, "_symm_handles", []) + [hdl]
            return arena, ptrs
        except Exception as e:   # no symmetric memory on this build / topology: NCCL path
            import sys
            sys.stderr.write("scrna_b200: symmetric memory unavailable (%s: %s); using the NCCL all-reduce path\n"
                             % (type(e).__name__, e))
            return None

    def gather_cols(self, local, bounds, n_total):
        """Every rank contributes `local` (r x n_local, the columns bounds[rank]); returns the full r x n_total
        tensor on every rank (zero-padded sum: payloads here are the small factors, not X)."""
        t = local.new_zeros((local.shape[0], int(n_total)))
        c0, c1 = bounds[self.rank]
        t[:, c0:c1].copy_(local[:, : c1 - c0])
        self.all_reduce(t)
        return t

    def all_gather_rows(self, full, bounds):
        """`full` (rows x cols) holds this rank's row block bounds[rank]; afterwards every rank holds all blocks."""
        for src, (r0, r1) in enumerate(bounds):
            if r1 > r0:
                self.dist.broadcast(full[r0:r1], src=self.dist.get_global_rank(self.group, src)
                                    if self.group is not None else src, group=self.group)


class NmfResult:
    def __init__(self, G, C, n_iter, flags, violations=None, errors=None):
        self.G = G            # g x k  float64 (replicated)
        self.C = C            # k x n_local float64 (this shard's cells)
        self.n_iter = n_iter
        self.flags = flags
        self.violation_init, self.violation_last = violations if violations else (None, None)
        self.errors = errors


class NmfSolver:
    """Device state of one NMF problem on one cell shard.

    Layouts: X genes x cells fp32 (+ the resident transposed copy X^T unless
    transpose_copy=False); factor masters component-major fp64 (G64: kp x ldg, C64: kp x ldn);
    fp32 row-major shadows Gs32 (g x kp) / Cs32 (n x kp) that the sweeps broadcast from shared
    memory, and the component-major fp32 shadow Ct32 for the residual kernel."""

    def __init__(self, X, k, kernels=None, comm=None, transpose_copy=True, tensor_cores="auto", XT=None):
        """tensor_cores: 'auto' = the 16-bit sweep (tcgen05 kind::f16 on fp16 storage of X and X^T) when every
        element of X is exactly representable in fp16 (raw counts); else the 3xTF32 tcgen05 sweep for k > 8
        (where the FP32 pipe falls behind HBM) and, when X is TF32-exact, also for k <= 8; else the FP32-pipe
        sweep.  True = as 'auto' but never the FP32 pipe; 'f16' / 'tf32' / False force one path."""
        if kernels is None:
            from .kernels import default_kernels
            kernels = default_kernels()
        self.kn = kernels
        self.torch = kernels.torch
        self.X = X
        self.k = int(k)
        # SCRNA_B200_TC overrides the automatic choice (tests, A/B timing): "0" FP32 pipe, "1" / "tf32" the 3xTF32
        # sweep, "f16" the 16-bit sweep wherever X allows it (else 3xTF32)
        env = os.environ.get("SCRNA_B200_TC", "")
        if env and tensor_cores == "auto":
            tensor_cores = {"0": False, "1": "tf32", "tf32": "tf32", "f16": True}.get(env, "auto")
        if tensor_cores not in ("auto", True, False, "f16", "tf32"):
            raise ValueError("tensor_cores must be 'auto', True, False, 'f16' or 'tf32'")
        self.comm = comm if (comm is not None and comm.world > 1) else None

        def everywhere(flag):
            # the shards must agree on the path (it sizes the all-reduce payload)
            bad = kernels.torch.tensor([0.0 if flag else 1.0], dtype=kernels.torch.float64, device=X.data.device)
            if self.comm is not None:
                self.comm.all_reduce(bad)
            return float(bad.item()) == 0.0

        can_h = (bool(transpose_copy) and tensor_cores in ("auto", True, "f16")
                 and getattr(kernels, "h_supported", lambda kq: False)(round_up(self.k, 8)))
        self.half = can_h and everywhere(X.is_f16_exact(kernels))
        if tensor_cores == "f16" and not self.half:
            raise ValueError("tensor_cores='f16' needs the transposed copy and an fp16-exact X")
        can_tc = (not self.half and bool(transpose_copy) and tensor_cores in ("auto", True, "tf32")
                  and getattr(kernels, "tc_supported", lambda kp: False)(pad_k(self.k, True)))
        exact = can_tc and everywhere(X.is_tf32_exact(kernels))
        if tensor_cores == "auto":
            self.tc = can_tc and (self.k > 8 or exact)
        else:
            self.tc = can_tc
        self.kp = round_up(self.k, 8) if self.half else pad_k(self.k, self.tc)
        self.tc_flags = 32 if (self.tc and exact) else 0
        t = self.torch
        dev = X.data.device
        g, n, kp = X.g, max(X.n, 1), self.kp
        self.ldn = ldn = X.ldx
        self.ldg = ldg = round_up(g, 32)
        self.gcols = round_up(g, 4)
        f64, f32 = t.float64, t.float32
        self.Xs = X.half_twin(kernels) if self.half else X     # the matrix the sweeps stream
        if XT is not None and transpose_copy and XT.is_half == self.half:
            self.XT = XT          # a resident transposed copy built by another solver on the same X
        else:
            self.XT = X.transposed(kernels, half=self.half) if transpose_copy else None
        self.h_parts = hp = int(os.environ.get("SCRNA_B200_H_PARTS", H_PARTS))
        self.arena = None
        if self.half and hasattr(kernels, "gstep_mu"):
            self._build_arena(g, kp, ldg, hp)     # gene-side state in ONE allocation that peers can map
        else:
            self.G64 = t.zeros((kp, ldg), dtype=f64, device=dev)
            self.Gs32 = t.zeros((g, kp), dtype=f32, device=dev)
        self.C64 = t.zeros((kp, ldn), dtype=f64, device=dev)
        self.Cs32 = t.zeros((n, kp), dtype=f32, device=dev)
        self.Ct32 = t.zeros((kp, ldn), dtype=f32, device=dev)
        self.Gtc = self.Ctc = self.Ch = self.Cinv = None
        if not self.half:
            self.Gh = self.Ginv = None
        if self.half:
            if self.arena is None:
                self.Gh = t.zeros(kernels.h_shadow_elems(g, kp, hp), dtype=t.float16, device=dev)
                self.Ginv = t.ones(kernels.h_groups(g) * kp, dtype=f32, device=dev)
            self.Ch = t.zeros(kernels.h_shadow_elems(n, kp, hp), dtype=t.float16, device=dev)
            self.Cinv = t.ones(kernels.h_groups(n) * kp, dtype=f32, device=dev)
            self.groupsC = int(kernels.h_groups(n))
            self.gpartC_grp = t.zeros(self.groupsC * kp * kp, dtype=f64, device=dev)
            self.SA = int(os.environ.get("SCRNA_B200_TC_SA", 0)) or kernels.wtx_h_splits(n, self.gcols, kp, hp)
            self.PA = t.zeros((self.SA, kp, ldg), dtype=f32, device=dev)
            self.SB = int(os.environ.get("SCRNA_B200_TC_SB", 0)) or kernels.wtx_h_splits(g, X.ncols, kp, hp)
        elif self.tc:
            self.Gtc = t.zeros(kernels.tc_shadow_elems(g, kp), dtype=f32, device=dev)
            self.Ctc = t.zeros(kernels.tc_shadow_elems(n, kp), dtype=f32, device=dev)
            self.SA = int(os.environ.get("SCRNA_B200_TC_SA", 0)) or kernels.wtx_tc_splits(n, self.gcols, kp)
            self.PA = t.zeros((self.SA, kp, ldg), dtype=f32, device=dev)
            self.SB = int(os.environ.get("SCRNA_B200_TC_SB", 0)) or kernels.wtx_tc_splits(g, X.ncols, kp)
        elif self.XT is not None:
            self.SA = kernels.wtx_splits(n, self.gcols, kp)
            self.PA = t.zeros((self.SA, kp, ldg), dtype=f32, device=dev)
            self.SB = kernels.wtx_splits(g, X.ncols, kp)
        else:
            self.SA = kernels.xht_splits(g, X.ncols, kp)
            self.PA = t.zeros((self.SA, g, kp), dtype=f32, device=dev)
            self.SB = kernels.wtx_splits(g, X.ncols, kp)
        self.PB = t.zeros((self.SB, kp, ldn), dtype=f32, device=dev)
        # all-reduce payload: [ (X.C^T)^T (kp*ldg) | C.C^T (kp*kp) | 4 scalars ]
        self.R = t.zeros(kp * ldg + kp * kp + 4, dtype=f64, device=dev)
        self.R_xht = self.R[: kp * ldg]
        self.R_cct = self.R[kp * ldg: kp * ldg + kp * kp]
        self.R_scal = self.R[kp * ldg + kp * kp:]
        self.NB = t.zeros((kp, ldn), dtype=f64, device=dev)
        if self.arena is None:
            self.gtg = t.zeros(kp * kp, dtype=f64, device=dev)
        self.gram_part = t.zeros(kernels.gram_blocks() * kp * kp, dtype=f64, device=dev)
        self.nblkG = kernels.update_blocks(g)
        self.nblkC = kernels.update_blocks(n)
        self.violG = t.zeros(self.nblkG, dtype=f64, device=dev)
        self.violC = t.zeros(self.nblkC, dtype=f64, device=dev)
        self.ctrl = t.zeros(16, dtype=t.int32, device=dev)
        self.res_part = t.zeros(max(1, kernels.residual_blocks(X.ncols, kp)), dtype=f64, device=dev)
        self.res_out = t.zeros(1, dtype=f64, device=dev)
        self.gpartG = t.zeros(self.nblkG * kp * kp, dtype=f64, device=dev)
        self.gpartC = t.zeros(self.nblkC * kp * kp, dtype=f64, device=dev)
        self.cell_range = None
        self.fuse_reduce = True   # partial reduction + Gram fused into the update epilogues
        self._g_num_from_R = True
        self.perms = None

    # ---- gene-side arena of the fused MU iteration (nmf_step.cu) -------------------------------
    def _build_arena(self, g, kq, ldg, parts):
        """Exchange buffer, flags and every gene-side array of the fused iteration in ONE allocation with the same
        layout on every rank.  With a P2P-capable communicator it is torch symmetric memory: every rank maps every
        peer's arena, and the fused gene step reads / writes the peers' copies directly over NVLink."""
        t, kn = self.torch, self.kn
        dev = self.X.data.device
        groups = int(kn.h_groups(g))
        sizes = [("xch", (kq * ldg + kq * kq) * 8), ("sync", 64 * 4), ("G64", kq * ldg * 8), ("Gs32", g * kq * 4),
                 ("Gh", int(kn.h_shadow_elems(g, kq, parts)) * 2), ("Ginv", groups * kq * 4),
                 ("gpart", groups * kq * kq * 8), ("gtg", kq * kq * 8)]
        off, total = {}, 0
        for name, nbytes in sizes:
            off[name] = total
            total += (nbytes + 255) // 256 * 256
        self.p2p = False
        arena = None
        # peer-mapped (symmetric) memory only for the opt-in fused gene step: its rendezvous is a collective that costs
        # tens of milliseconds per solver, which the default NCCL path has no use for
        if (self.comm is not None and getattr(self.comm, "symmetric_arena", None) is not None
                and os.environ.get("SCRNA_B200_FUSED", "0") == "1"):
            got = self.comm.symmetric_arena(total, dev)
            if got is not None:
                arena, ptrs = got
                self.p2p = True
        if arena is None:
            arena = t.zeros(total, dtype=t.uint8, device=dev)
            ptrs = [arena.data_ptr()]
        self.arena, self.arena_off, self.groupsG = arena, off, groups
        self.peer_ptrs = t.tensor(ptrs, dtype=t.int64, device=dev)

        def view(name, dtype, shape):
            nbytes = int(np.prod(shape)) * t.empty(0, dtype=dtype).element_size()
            return arena[off[name]: off[name] + nbytes].view(dtype).view(*shape)
        self.G64 = view("G64", t.float64, (kq, ldg))
        self.Gs32 = view("Gs32", t.float32, (g, kq))
        self.Gh = view("Gh", t.float16, (int(kn.h_shadow_elems(g, kq, parts)),))
        self.Ginv = view("Ginv", t.float32, (groups * kq,))
        self.Ginv.fill_(1.0)
        self.gtg = view("gtg", t.float64, (kq * kq,))
        self.xch = view("xch", t.float64, (kq * ldg + kq * kq,))

    # ---- helpers ------------------------------------------------------------------------------
    def load_G(self, G0):
        """Gene factor (g x k host array) -> fp64 master + the operand shadows of the sweeps."""
        t = self.torch
        X, k, kp = self.X, self.k, self.kp
        G0 = np.ascontiguousarray(np.asarray(G0, dtype=np.float64).T)
        if G0.shape != (k, X.g):
            raise ValueError("init shapes do not match X %dx%d, k=%d" % (X.g, X.n, k))
        self.G64.zero_()
        self.G64[:k, : X.g].copy_(t.from_numpy(G0).to(X.data.device))
        self.kn.shadows(self.G64, self.ldg, X.g, k, kp, self.Gs32, None)
        if self.tc:
            self.kn.tc_shadow(self.G64, self.ldg, X.g, k, kp, self.Gtc)
        if self.half:   # the fp16 parts are scaled by the column norm: Gram first
            self._gram_G(None)
            self.kn.h_shadow(self.G64, self.ldg, X.g, k, kp, self.h_parts, self.gtg, 1, kp, None, self.Gh, self.Ginv)

    def load_C(self, C0):
        """Cell factor (k x n host array) -> fp64 master + shadows."""
        t = self.torch
        X, k, kp = self.X, self.k, self.kp
        C0 = np.ascontiguousarray(C0, dtype=np.float64)
        if C0.shape != (k, X.n):
            raise ValueError("init shapes do not match X %dx%d, k=%d" % (X.g, X.n, k))
        self.C64.zero_()
        self.C64[:k, : X.n].copy_(t.from_numpy(C0).to(X.data.device))
        self.kn.shadows(self.C64, self.ldn, max(X.n, 1), k, kp, self.Cs32, self.Ct32)
        if self.tc:
            self.kn.tc_shadow(self.C64, self.ldn, max(X.n, 1), k, kp, self.Ctc)
        if self.half:
            self._gram_C(None)
            self.kn.h_shadow(self.C64, self.ldn, max(X.n, 1), k, kp, self.h_parts, self.R_cct, 1, kp, None, self.Ch,
                             self.Cinv)

    def _load_factors(self, G0, C0):
        self.load_G(G0)
        self.load_C(C0)

    # ---- skinny products with the resident X (device NNDSVD init, scrna_b200/nmf_init.py) -------------
    def set_cell_range(self, c0, c1, n_total):
        """This shard holds cells [c0, c1) of n_total: `x_times` then takes the GLOBAL (n_total x k) operand and
        `xt_times_full` returns the global product (device NNDSVD init on cell shards)."""
        self.cell_range = (int(c0), int(c1), int(n_total))

    def xt_times_full(self, Q):
        """X^T @ Q over all shards: (n_total x k), identical on every rank."""
        loc = self.xt_times(Q)
        if self.comm is None or self.cell_range is None:
            return loc
        c0, c1, n_total = self.cell_range
        t = self.torch
        full = t.zeros((n_total, loc.shape[1]), dtype=t.float64, device=self.X.data.device)
        full[c0:c1].copy_(t.from_numpy(loc).to(full.device))
        self.comm.all_reduce(full)
        return full.cpu().numpy()

    def x_times(self, Q):
        """X @ Q for a host (n x k) matrix Q of any sign: one sweep of the X^T copy.  Returns g x k fp64 (summed
        over the cell shards).  With a cell range set, Q is the global operand and this shard uses its rows."""
        Q = np.asarray(Q)
        if self.cell_range is not None and Q.shape[0] == self.cell_range[2]:
            Q = Q[self.cell_range[0]: self.cell_range[1]]
        self.load_C(Q.T)
        keep = self._g_num_from_R
        self._g_num_from_R = True
        self._products_for_G(None)
        self._g_num_from_R = keep
        self._allreduce_R()
        return np.ascontiguousarray(self.R_xht.view(self.kp, self.ldg)[: self.k, : self.X.g].cpu().numpy().T)

    def xt_times(self, Q):
        """X^T @ Q for a host (g x k) matrix Q: one sweep of X.  Returns this shard's n x k fp64."""
        kn, X, kp = self.kn, self.X, self.kp
        self.load_G(Q)
        self._sweep_X(None)
        kn.reduce_cols(self.PB, self.SB, kp, self.ldn, X.ncols, self.NB, None)
        return np.ascontiguousarray(self.NB[: self.k, : X.n].cpu().numpy().T)

    def x_sum(self):
        """sum_ij |X_ij| over this shard (= sum X for the non-negative NMF input): the residual kernel
        against all-zero factors."""
        X = self.Xs
        self.Gs32.zero_()
        self.Ct32.zero_()
        self.kn.residual(X.data, X.ldx, X.g, X.ncols, self.Gs32, self.Ct32, self.ldn, self.kp, 1, self.res_part,
                         self.res_out)
        if self.comm is not None:
            self.comm.all_reduce(self.res_out)
        return float(self.res_out.cpu()[0])

    def read_ctrl(self):
        raw = self.ctrl.cpu().numpy()
        ints = raw[:4]
        dbl = raw[4:].view(np.float64)
        return dict(iter=int(ints[0]), done=int(ints[1]), n_iter=int(ints[2]), flags=int(ints[3]),
                    viol_init=float(dbl[0]), viol_last=float(dbl[1]), err_last=float(dbl[2]))

    def factors_to_host(self):
        X, k = self.X, self.k
        G = np.ascontiguousarray(self.G64[:k, : X.g].cpu().numpy().T)
        C = self.C64[:k, : X.n].cpu().numpy().copy()
        return G, C

    # ---- the two sweeps -----------------------------------------------------------------------
    def _sweep_X(self, ctrl):
        """K3: G^T . X over this shard -> split partials PB (SB x kp x ldn)."""
        kn, X, kp = self.kn, self.Xs, self.kp
        if self.half:
            kn.wtx_partial_h(X.data, X.rows_pad, X.g, X.ncols, self.Gh, self.Ginv, kp, self.h_parts, self.PB,
                             self.ldn, self.SB, ctrl)
        elif self.tc:
            kn.wtx_partial_tc(X.data, X.ldx, X.g, X.ncols, self.Gtc, kp, self.PB, self.ldn, self.SB, ctrl,
                              flags=self.tc_flags)
        else:
            kn.wtx_partial(X.data, X.ldx, X.g, X.ncols, self.Gs32, kp, self.PB, self.ldn, self.SB, ctrl)

    def _sweep_XT(self, ctrl):
        """K1 on the transposed copy: (X . C^T)^T -> split partials PA (SA x kp x ldg)."""
        kn, XT, kp = self.kn, self.XT, self.kp
        if self.half:
            kn.wtx_partial_h(XT.data, XT.rows_pad, XT.g, self.gcols, self.Ch, self.Cinv, kp, self.h_parts, self.PA,
                             self.ldg, self.SA, ctrl)
        elif self.tc:
            kn.wtx_partial_tc(XT.data, XT.ldx, XT.g, self.gcols, self.Ctc, kp, self.PA, self.ldg, self.SA, ctrl,
                              flags=self.tc_flags)
        else:
            kn.wtx_partial(XT.data, XT.ldx, XT.g, self.gcols, self.Cs32, kp, self.PA, self.ldg, self.SA, ctrl)

    # ---- the two half steps -------------------------------------------------------------------
    def _gram_C(self, ctrl):
        self.kn.gram(self.C64, 1, self.ldn, max(self.X.n, 1), self.k, self.kp, self.gram_part, self.R_cct, ctrl)

    def _gram_G(self, ctrl):
        self.kn.gram(self.G64, 1, self.ldg, self.X.g, self.k, self.kp, self.gram_part, self.gtg, ctrl)

    def _products_for_G(self, ctrl, with_gram=False):
        """K1: this shard's (X.C^T)^T as split partials PA (and, when they must travel through the
        all-reduce buffer, their fixed-order sum R.xht).  R.cct = C.C^T is produced by the previous
        cell-factor epilogue (or by _gram_C for the initial / a fixed C)."""
        kn, X, kp = self.kn, self.X, self.kp
        if self.XT is not None:
            self._sweep_XT(ctrl)
            if self._g_num_from_R:
                kn.reduce_cols(self.PA, self.SA, kp, self.ldg, self.gcols, self.R_xht, ctrl)
        else:
            kn.xht_partial(X.data, X.ldx, X.g, X.ncols, self.Ct32, self.ldn, kp, self.PA, self.SA, ctrl)
            kn.reduce_rows(self.PA, self.SA, X.g, kp, self.R_xht, self.ldg, ctrl)
        if with_gram:
            self._gram_C(ctrl)

    def _allreduce_R(self):
        if self.comm is not None:
            self.comm.all_reduce(self.R)

    def _update_G(self, solver, l1, l2, stride, offset, ctrl):
        fused = not self._g_num_from_R
        self.kn.update(solver, self.G64, self.ldg, self.X.g, self.k, self.kp,
                       None if fused else self.R_xht, self.R_cct, l1, l2, self.perms, stride, offset,
                       self.Gs32, None, self.violG, ctrl,
                       num_parts=self.PA if fused else None, nsplit=self.SA if fused else 0,
                       gram_partials=self.gpartG, gram_out=self.gtg, TC32=self.Gtc, Fh=self.Gh, inv_scale=self.Ginv,
                       kq=self.kp, parts=self.h_parts)

    def _half_C(self, solver, l1, l2, stride, offset, ctrl):
        kn, X, kp = self.kn, self.X, self.kp
        self._sweep_X(ctrl)
        if self.fuse_reduce:
            kn.update(solver, self.C64, self.ldn, max(X.n, 1), self.k, kp, None, self.gtg, l1, l2,
                      self.perms, stride, offset, self.Cs32, self.Ct32, self.violC, ctrl,
                      num_parts=self.PB, nsplit=self.SB, gram_partials=self.gpartC, gram_out=self.R_cct,
                      TC32=self.Ctc, Fh=self.Ch, inv_scale=self.Cinv, kq=kp, parts=self.h_parts)
        else:
            kn.reduce_cols(self.PB, self.SB, kp, self.ldn, X.ncols, self.NB, ctrl)
            kn.update(solver, self.C64, self.ldn, max(X.n, 1), self.k, kp, self.NB, self.gtg, l1, l2,
                      self.perms, stride, offset, self.Cs32, self.Ct32, self.violC, ctrl,
                      gram_partials=self.gpartC, gram_out=self.R_cct, TC32=self.Ctc, Fh=self.Ch, inv_scale=self.Cinv,
                      kq=kp, parts=self.h_parts)

    def frobenius_error(self):
        """||X - G.C||_F over all shards (sklearn _beta_divergence(..., square_root=True))."""
        X = self.Xs
        self.kn.residual(X.data, X.ldx, X.g, X.ncols, self.Gs32, self.Ct32, self.ldn, self.kp, 2, self.res_part,
                         self.res_out)
        if self.comm is not None:
            self.comm.all_reduce(self.res_out)
        return math.sqrt(max(float(self.res_out.cpu()[0]), 0.0))

    # ---- public -------------------------------------------------------------------------------
    def prepare(self, G0, C0, solver="cd", l1_G=0.0, l2_G=0.0, l1_C=0.0, l2_C=0.0, tol=1e-4, max_iter=200,
                update_G=True, update_C=True, first="G", seed=0, shuffle=True, perms=None):
        """Load the factors, pre-draw the permutation stream and build the per-iteration launch
        sequence (`self.step()`); nothing is iterated yet."""
        if solver not in ("cd", "mu"):
            raise ValueError("solver must be 'cd' or 'mu'")
        if first not in ("G", "C"):
            raise ValueError("first must be 'G' or 'C'")
        if not (update_G or update_C):
            raise ValueError("nothing to update")
        t, kn = self.torch, self.kn
        code = _lib.SOLVER_CD if solver == "cd" else _lib.SOLVER_MU
        max_iter = int(max_iter)
        self._load_factors(G0, C0)
        ctrl = self.ctrl
        kn.ctrl_reset(ctrl)

        draws = int(update_G) + int(update_C)
        if solver == "cd":
            if perms is None:
                rng = np.random.RandomState(seed)
                host = np.zeros((max(1, max_iter) * draws, self.kp), dtype=np.int32)
                for i in range(host.shape[0]):
                    host[i, : self.k] = rng.permutation(self.k) if shuffle else np.arange(self.k)
            else:
                p = np.asarray(perms, dtype=np.int32)
                host = np.zeros((p.shape[0], self.kp), dtype=np.int32)
                host[:, : self.k] = p
                if host.shape[0] < max_iter * draws:
                    raise ValueError("not enough pre-drawn permutations")
            self.perms = t.from_numpy(host).to(self.X.data.device)
        else:
            self.perms = None
        offG = 0 if (first == "G" or not update_C) else 1
        offC = 0 if (first == "C" or not update_G) else 1
        cd_tol = float(tol) if solver == "cd" else -1.0

        # Grams of the initial factors (afterwards every epilogue emits the Gram of what it just wrote)
        self._gram_C(None)
        self._gram_G(None)
        # a fixed factor makes its products constant: compute them once
        # R.xht must be materialised when it travels through the all-reduce, when the low-memory K1
        # produced row-major partials, or when C is fixed (then it is computed once and reused)
        self._g_num_from_R = (self.comm is not None) or (self.XT is None) or (not update_C) or (not self.fuse_reduce)
        if not update_C:
            self._products_for_G(None)
            self._allreduce_R()

        use_iter_end = solver == "cd"
        # The fused step kernels (nmf_step.cu) carry the cross-GPU exchange inside the gene step (peer loads / stores on
        # torch symmetric memory, no NCCL in the loop).  They are correct on 1, 2 and 8 GPUs (same Frobenius error to
        # 1e-9) but measured SLOWER than the separate epilogues + one NCCL all-reduce: 1.2405 vs 1.2196 ms per
        # iteration on one B200, 0.674 vs 0.651 ms on two, 0.2588 vs 0.2392 ms on eight (20k x 100k, k = 8;
        # profiles/r02_fused_step_vs_nccl.md) -- two in-kernel grid + cross-GPU barriers cost more than they save at
        # this payload (1.3 MB).  Hence opt-in: SCRNA_B200_FUSED=1.
        fused = (solver == "mu" and self.half and self.arena is not None and update_G and update_C and first == "G"
                 and (self.comm is None or self.p2p) and os.environ.get("SCRNA_B200_FUSED", "0") == "1")
        self.fused = fused
        if fused:
            # the fused gene step sums per-group Gram partials of the cell factor: seed them with the initial C.C^T
            self.gpartC_grp.zero_()
            self.gpartC_grp[: self.kp * self.kp].copy_(self.R_cct)
            nanflag = self.ctrl[3:4]
            nb = kn.h_nb(self.kp, self.h_parts)
            world = self.comm.world if self.comm is not None else 1
            rank = self.comm.rank if self.comm is not None else 0
            if self.comm is not None:
                self.torch.cuda.synchronize()
                self.comm.barrier()          # every rank's arena holds its initial factors before a peer touches it

            def fused_iteration():
                X = self.X
                self._sweep_XT(None)
                kn.gstep_mu(self.peer_ptrs, world, rank, self.arena_off, X.g, self.ldg, self.k, self.kp, nb, self.h_parts,
                            l1_G, l2_G, self.PA, self.SA, self.gpartC_grp, self.groupsC, nanflag)
                self._sweep_X(None)
                kn.cstep_mu(self.C64, self.ldn, max(X.n, 1), self.k, self.kp, nb, self.h_parts, self.PB, self.SB, self.gtg,
                            l1_C, l2_C, self.Cs32, self.Ct32, self.Ch, self.Cinv, self.gpartC_grp, nanflag)

            self.step = fused_iteration
            self._one_iteration = fused_iteration
            self._graph = None
            self._plan = dict(solver=solver, tol=float(tol), max_iter=max_iter)
            return

        def one_iteration():
            if first == "G":
                if update_G:
                    if update_C:
                        self._products_for_G(ctrl)
                        self._allreduce_R()
                    self._update_G(code, l1_G, l2_G, draws, offG, ctrl)
                if update_C:
                    self._half_C(code, l1_C, l2_C, draws, offC, ctrl)
                    if self.comm is not None and solver == "cd":
                        kn.sum_f64(self.violC, self.nblkC, self.R_scal[0:1])
                        self.comm.all_reduce(self.R_scal[0:1])
            else:
                if update_C:
                    self._half_C(code, l1_C, l2_C, draws, offC, ctrl)
                    if not update_G and self.comm is not None and solver == "cd":
                        # the cell factor alone is updated (fixed dictionary): its violation is the whole stop rule
                        kn.sum_f64(self.violC, self.nblkC, self.R_scal[0:1])
                        self.comm.all_reduce(self.R_scal[0:1])
                if update_G:
                    if update_C:
                        self._products_for_G(ctrl)
                        if self.comm is not None and solver == "cd":
                            kn.sum_f64(self.violC, self.nblkC, self.R_scal[0:1])
                        self._allreduce_R()
                    self._update_G(code, l1_G, l2_G, draws, offG, ctrl)
            va, na = (self.violG, self.nblkG) if update_G else (None, 0)
            if update_C:
                vb, nb = (self.R_scal, 1) if self.comm is not None else (self.violC, self.nblkC)
            else:
                vb, nb = None, 0
            if use_iter_end:
                kn.iter_end(ctrl, va, na, vb, nb, cd_tol, max_iter)

        self.step = one_iteration
        self._one_iteration = one_iteration
        self._graph = None
        self._plan = dict(solver=solver, tol=float(tol), max_iter=max_iter)

    def capture_graph(self, iterations=1):
        """Capture `iterations` iterations (kernels and, when sharded, the NCCL all-reduce) into one CUDA graph
        and make `step()` replay it: the loop has no host decision inside -- the CD stop rule lives in the device
        control block and every kernel early-exits once `done` is set.  Worth it where the iteration is short
        (cell shards on 4-8 GPUs: ~0.4 ms for 7 kernels + one collective); call after `prepare()` and a few
        eager warm-up steps.  `step()` then advances `iterations` iterations per call."""
        t = self.torch
        if self.kn.sweep_events is not None:
            raise RuntimeError("per-sweep event timing and graph capture are mutually exclusive")
        t.cuda.synchronize()
        g = t.cuda.CUDAGraph()
        side = t.cuda.Stream()
        side.wait_stream(t.cuda.current_stream())
        with t.cuda.stream(side):
            with t.cuda.graph(g, stream=side):
                for _ in range(int(iterations)):
                    self._one_iteration()
        t.cuda.current_stream().wait_stream(side)
        self._graph = g
        self.step = g.replay
        return g

    def release(self):
        """Drop the captured CUDA graph and the launch-sequence closures (`step` <-> `one_iteration` reference
        this object, so `del solver` alone never frees a graph that captured NCCL work).  Call before
        `torch.distributed.destroy_process_group()`."""
        self._graph = None
        self.step = None
        self._one_iteration = None

    def fit(self, G0, C0, solver="cd", l1_G=0.0, l2_G=0.0, l1_C=0.0, l2_C=0.0, tol=1e-4, max_iter=200,
            update_G=True, update_C=True, first="G", seed=0, shuffle=True, perms=None, poll=4,
            to_host=True):
        """Run the alternating updates from (G0, C0).  Returns NmfResult.

        solver 'cd' reproduces sklearn's coordinate descent (stop: violation/violation_init <= tol),
        solver 'mu' sklearn's multiplicative update (stop: every 10 iterations,
        (previous_error - error)/error_at_init < tol; tol=0 runs exactly max_iter iterations)."""
        self.prepare(G0, C0, solver=solver, l1_G=l1_G, l2_G=l2_G, l1_C=l1_C, l2_C=l2_C, tol=tol,
                     max_iter=max_iter, update_G=update_G, update_C=update_C, first=first, seed=seed,
                     shuffle=shuffle, perms=perms)
        max_iter = int(max_iter)
        errors = None
        error_at_init = previous_error = None
        if solver == "mu" and tol > 0:
            error_at_init = previous_error = self.frobenius_error()
            errors = [(0, error_at_init)]
        n_done = 0
        while n_done < max_iter:
            self.step()
            n_done += 1
            if solver == "mu":
                if tol > 0 and n_done % 10 == 0:
                    error = self.frobenius_error()
                    errors.append((n_done, error))
                    if (previous_error - error) / error_at_init < tol:
                        break
                    previous_error = error
            elif n_done % poll == 0 or n_done == max_iter:
                if self.read_ctrl()["done"]:
                    break
        state = self.read_ctrl()
        n_iter = state["n_iter"] if (solver == "cd" and state["done"]) else n_done
        res = NmfResult(None, None, n_iter, state["flags"], (state["viol_init"], state["viol_last"]), errors)
        if to_host:
            res.G, res.C = self.factors_to_host()
        return res
