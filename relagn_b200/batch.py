"""Batch driver: evaluates many parameter sets per call on one GPU, or sharded over the GPUs of a box.

PyTorch owns the buffers (device tensors, pinned host staging) and provides the stream and the process group;
every number comes from the CUDA library behind relagn_b200._capi (no CPU fallback).

The reference evaluates one parameter set per Python object (relagn.py:232-357 ctor + get_totSED :1552-1577);
this driver is the batched form of exactly that path: params rows follow the constructor's positional order.
"""
import numpy as np
import torch

from . import _capi
from . import tables as _tables

MODELS = {"relagn": _capi.MODEL_RELAGN, "relqso": _capi.MODEL_RELQSO}


class BatchEvaluator:
    """Spectra for batches of parameter sets on one CUDA device.

    ebins  : nE+1 bin edges in keV (the reference's `Ebins` / `new_ear`, relagn.py:335,780-805)
    model  : "relagn" (15 params) or "relqso" (7 params)
    tables : None -> default Kerr transfer tables (generated on the device on first use and cached on disk),
             or a dict(spins, thetas, NR, NPHI, r_in, data) to upload.
    """

    def __init__(self, ebins, model="relagn", device=None, tables=None, dr_dex=50.0):
        if not torch.cuda.is_available():
            raise _capi.RelagnError(_capi.S_OK - 1, "no CUDA device: relagn_b200 has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
        self.model = MODELS[model]
        self.dr_dex = float(dr_dex)
        self.ctx = _capi.Context(self.device.index)
        self.ctx.set_grid(np.asarray(ebins, dtype=np.float64))
        self.nE = self.ctx.nE
        self._tables_ready = False
        self._tables_arg = tables
        self._pin_in = None
        self._pin_out = None

    def ensure_tables(self):
        if self._tables_ready:
            return
        t = self._tables_arg
        if t is None:
            _tables.load_default(self.ctx)
        else:
            self.ctx.tables_upload(t["spins"], t["thetas"], t["NR"], t["NPHI"], t["r_in"], t["data"])
        self._tables_ready = True

    # ---- device-resident path -------------------------------------------------
    def evaluate_device(self, params, rel=True, components=False, out=None, attrs=None, fp32=False):
        """params: CUDA float64 tensor [P, 15].  Returns (out, attrs) CUDA tensors; asynchronous on the current
        torch stream."""
        assert params.is_cuda and params.dtype == torch.float64 and params.is_contiguous()
        assert params.shape[1] == _capi.NPAR
        if rel:
            self.ensure_tables()
        P = params.shape[0]
        shape = (P, 4, self.nE) if components else (P, self.nE)
        if out is None:
            out = torch.empty(shape, dtype=torch.float64, device=self.device)
        if attrs is None:
            attrs = torch.empty((P, _capi.NATTR), dtype=torch.float64, device=self.device)
        assert tuple(out.shape) == shape and out.is_contiguous()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.ctx.eval_batch_device(params.data_ptr(), P, out.data_ptr(), attrs.data_ptr(), model=self.model, rel=rel,
                                   components=components, dr_dex=self.dr_dex, stream=stream, fp32=fp32)
        return out, attrs

    # ---- host path (the call a user of the reference's class makes, batched) ----
    def evaluate_device_pieces(self, params, on_piece, piece_sets=16384, rel=True, components=False, out=None,
                               attrs=None, fp32=False):
        """evaluate_device with a hook: on_piece(first_set, n_sets) is called as soon as the kernels that complete
        those rows of `out` are enqueued on the current torch stream (rg_eval_batch_pieces) -- record an event there
        and move the piece (NCCL gather, copy) on another stream while the next one is computed."""
        assert params.is_cuda and params.dtype == torch.float64 and params.is_contiguous()
        assert params.shape[1] == _capi.NPAR
        if rel:
            self.ensure_tables()
        P = params.shape[0]
        shape = (P, 4, self.nE) if components else (P, self.nE)
        if out is None:
            out = torch.empty(shape, dtype=torch.float64, device=self.device)
        if attrs is None:
            attrs = torch.empty((P, _capi.NATTR), dtype=torch.float64, device=self.device)
        assert tuple(out.shape) == shape and out.is_contiguous()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.ctx.eval_batch_device_pieces(params.data_ptr(), P, out.data_ptr(), piece_sets, on_piece,
                                          attrs_ptr=attrs.data_ptr(), model=self.model, rel=rel, components=components,
                                          dr_dex=self.dr_dex, stream=stream, fp32=fp32)
        return out, attrs

    def evaluate(self, params, rel=True, components=False, fp32=False):
        """params: array-like [P, n] on the host.  Returns (spectra ndarray [P, nE] or [P, 4, nE] in W/Hz,
        attrs ndarray [P, 24]).  Stages through pinned memory; H2D and D2H happen inside this call.  The returned
        arrays are views of the evaluator's pinned staging buffers: copy them before the next call if both are needed."""
        p = np.asarray(params, dtype=np.float64)
        if p.ndim == 1:
            p = p[None, :]
        P = p.shape[0]
        if rel:
            self.ensure_tables()
        ncomp = 4 if components else 1
        if self._pin_in is None or self._pin_in.shape[0] < P:
            self._pin_in = torch.zeros((P, _capi.NPAR), dtype=torch.float64).pin_memory()
        if self._pin_out is None or self._pin_out.numel() < P * ncomp * self.nE:
            self._pin_out = torch.empty(P * ncomp * self.nE, dtype=torch.float64).pin_memory()
        if getattr(self, "_pin_attr", None) is None or self._pin_attr.shape[0] < P:
            self._pin_attr = torch.empty((P, _capi.NATTR), dtype=torch.float64).pin_memory()
        pin = self._pin_in[:P]
        pin_np = pin.numpy()
        if p.shape[1] == _capi.NPAR:
            np.copyto(pin_np, p)  # one straight memcpy into the pinned staging rows
        else:
            pin.zero_()
            pin_np[:, :p.shape[1]] = p
        out = self._pin_out[:P * ncomp * self.nE].view((P, 4, self.nE) if components else (P, self.nE))
        out_np, at = self.ctx.eval_batch_host(pin_np, model=self.model, rel=rel, components=components,
                                              dr_dex=self.dr_dex, out=out.numpy(), fp32=fp32,
                                              attrs_out=self._pin_attr[:P].numpy())
        return out_np, at


    # ---- observation epilogue on the device ---------------------
    def convert_device(self, spectra, params, unit="cgs", as_flux=False, out=None):
        """The reference's getter epilogue (_to_newUnit / _to_flux, relagn.py:487-564) for a whole batch on the device.
        spectra: CUDA [P, nE] or [P, 4, nE] W/Hz; params: CUDA [P, 15].  out=None converts in place."""
        assert spectra.is_cuda and spectra.dtype == torch.float64 and spectra.is_contiguous()
        P = spectra.shape[0]
        ncomp = spectra.shape[1] if spectra.dim() == 3 else 1
        if out is None:
            out = spectra
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.ctx.convert_device(spectra.data_ptr(), params.data_ptr(), P, ncomp, unit, as_flux, out.data_ptr(),
                                model=self.model, stream=stream)
        return out

    def set_instrument(self, ear, rowptr=None, col=None, val=None, exposure=1.0):
        """Observer-frame instrument grid (bin edges, keV) and optionally its response in CSR over channels."""
        self.ctx.set_instrument(ear)
        if rowptr is not None:
            self.ctx.set_response(rowptr, col, val, exposure)

    def set_data(self, counts, sigma=None, stat="chi2"):
        self.ctx.set_data(counts, sigma, stat)

    def photar_device(self, spectra, params, out=None):
        """XSPEC additive-model output (relagn.f:98-112): photons/cm^2/s per instrument bin, CUDA [P, ne]."""
        P = spectra.shape[0]
        if out is None:
            out = torch.empty((P, self.ctx.ne), dtype=torch.float64, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.ctx.photar_device(spectra.data_ptr(), params.data_ptr(), P, out.data_ptr(), model=self.model,
                               stream=stream)
        return out

    def fitstat_device(self, params, rel=True, out=None, fp32=False):
        """params CUDA [P, 15] -> fit statistic CUDA [P]; the spectra never leave the device."""
        assert params.is_cuda and params.dtype == torch.float64 and params.is_contiguous()
        if rel:
            self.ensure_tables()
        P = params.shape[0]
        if out is None:
            out = torch.empty(P, dtype=torch.float64, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.ctx.eval_fitstat_device(params.data_ptr(), P, out.data_ptr(), model=self.model, rel=rel,
                                     dr_dex=self.dr_dex, stream=stream, fp32=fp32)
        return out

    def fitstat(self, params, rel=True, fp32=False):
        """Host in, host out: 120 B per set up, 8 B per set down (the fit / MCMC inner loop)."""
        p = np.asarray(params, dtype=np.float64)
        if p.ndim == 1:
            p = p[None, :]
        P = p.shape[0]
        if rel:
            self.ensure_tables()
        if self._pin_in is None or self._pin_in.shape[0] < P:
            self._pin_in = torch.zeros((P, _capi.NPAR), dtype=torch.float64).pin_memory()
        if getattr(self, "_pin_stat", None) is None or self._pin_stat.numel() < P:
            self._pin_stat = torch.empty(P, dtype=torch.float64).pin_memory()
        pin = self._pin_in[:P]
        if p.shape[1] != _capi.NPAR:
            pin.zero_()
        pin[:, :p.shape[1]] = torch.from_numpy(p)
        return self.ctx.eval_fitstat_host(pin.numpy(), model=self.model, rel=rel, dr_dex=self.dr_dex, fp32=fp32,
                                          out=self._pin_stat[:P].numpy())


def shard_bounds(P, world, rank):
    """Contiguous split of P parameter sets over `world` ranks."""
    base, rem = divmod(P, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def piece_edges(P, pieces=None, chunk_max=65536, piece_min=8192):
    """Piece boundaries of a shard of P sets.  pieces=None: every piece is half of what is left (multiples of
    `piece_min`, at most the library's chunk of `chunk_max` sets), so the early pieces run at the full chunk size and
    the last one -- whose transfer nobody hides -- is small (the same plan as rg_eval_batch_host's slabs).
    pieces=n: n roughly equal pieces on multiples of 16384 sets."""
    if pieces is not None:
        n_chunks = max(1, (P + 16383) // 16384)
        pieces = max(1, min(int(pieces), n_chunks))
        return [min(P, ((n_chunks * i) // pieces) * 16384) for i in range(pieces)] + [P]
    edges, lo = [0], 0
    while lo < P:
        left = P - lo
        n = left if left <= piece_min + piece_min // 2 else min(chunk_max, max(piece_min, (left // 2) // piece_min * piece_min))
        lo += n
        edges.append(lo)
    return edges


class PeerGather:
    """Gather of the spectra to one rank WITHOUT a gather step: every rank owns a window [2, world, P_max, nE] in torch
    symmetric memory and maps the destination rank's window; `evaluate_sharded(transport=...)` hands the library the
    rank's own rows of the DESTINATION's window as the output buffer, so the spectrum kernel's epilogue stores (coalesced
    8 kB rows) travel straight over NVLink / NVSwitch -- no NCCL kernels competing with the persistent compute
    kernels for SMs, no copy engines, no host callbacks; one barrier closes the step.  The window is double buffered
    (step parity): while the destination's consumers read step k, step k + 1 is written into the other half; before
    step k + 2 reuses a half, the destination waits for everything it had enqueued when step k + 1 began.
    Shards may differ in size (P_max = the largest, agreed at construction).  Needs P2P access between all GPUs of
    the group and a driver that exports symmetric allocations; construct it collectively."""

    def __init__(self, P, nE, device, group=None, dst=0):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank, self.dst = dist.get_world_size(self.group), dist.get_rank(self.group), dst
        sizes = torch.zeros(self.world, dtype=torch.int64, device=device)
        sizes[self.rank] = P
        dist.all_reduce(sizes, group=self.group)
        self.sizes = [int(v) for v in sizes.cpu()]
        self.P, self.P_max, self.nE = P, max(self.sizes), nE
        self.window = symm_mem.empty((2, self.world, self.P_max, nE), dtype=torch.float64, device=device)
        self.hdl = symm_mem.rendezvous(self.window, self.group)
        self.remote = self.hdl.get_buffer(dst, (2, self.world, self.P_max, nE), torch.float64)
        self.step = 0
        self.device = device
        self._mark = [None, None]  # destination: event recorded when the step of that parity began

    def begin(self):
        """Start of a step: returns this rank's output rows [P, nE] inside the destination's window."""
        import torch.distributed as dist
        half = self.step & 1
        if self.rank == self.dst:
            # everything enqueued on the current stream up to the start of the previous step (the consumers of the step
            # that last used this half among it) must have finished before any rank overwrites the half
            prev = self._mark[half ^ 1]
            if prev is not None:
                prev.synchronize()
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._mark[half] = ev
        if self.step >= 2:
            dist.barrier(group=self.group)
        return self.remote[half, self.rank, :self.P]

    def finish(self):
        """All rows of all ranks have landed when this returns: on the destination rank the list of per-rank views
        [P_r, nE] of the window half of this step (valid until the step after the next one begins)."""
        import torch.distributed as dist
        half = self.step & 1
        torch.cuda.current_stream(self.device).synchronize()  # this rank's stores (incl. the remote ones) are done
        dist.barrier(group=self.group)
        self.step += 1
        if self.rank != self.dst:
            return None
        return [self.window[half, r, :self.sizes[r]] for r in range(self.world)]


def evaluate_sharded(evaluator, params_local, rel=True, gather_dst=0, group=None, pieces=None, out=None, attrs=None,
                     gathered=None, hook=False, piece_sets=16384, transport=None):
    """One process per GPU: every rank evaluates its own shard (device tensor [P_local, 15]); the only exchange is the
    gather of the spectra to `gather_dst` (gather_dst=None: no gather).
    transport=PeerGather(...): the library writes the spectra straight into the destination rank's window over
    NVLink (see PeerGather); `out` is then ignored and the returned local spectra ARE this rank's window rows.
    Otherwise NCCL: the shard is evaluated in contiguous pieces (`piece_edges`) and the gather of piece i is issued on a
    side stream while piece i+1 is computed, so only the last (small) piece's transfer is exposed; shards of different
    size are padded to the largest (one extra all-reduce of the sizes) and trimmed on the destination.  hook=True: the
    library evaluates the shard in its own large chunks and the gathers are issued from the completion hook of
    rg_eval_batch_pieces instead -- measured slower at 8 GPUs (the NCCL kernels of eight gathers per step compete
    with the persistent spectrum kernel for the SMs of the receiving rank).  out / attrs / gathered: optional
    preallocated result tensors ([P_local, nE], [P_local, 24], list of world x [P_max, nE] on the destination rank).
    Returns (local_out, list of per-rank spectra on the destination rank or None)."""
    import torch.distributed as dist
    P = params_local.shape[0]
    multi = gather_dst is not None and dist.is_initialized() and dist.get_world_size(group) > 1
    if not multi:
        if out is not None:
            evaluator.evaluate_device(params_local, rel=rel, out=out, attrs=attrs)
            return out, None
        return evaluator.evaluate_device(params_local, rel=rel)[0], None
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    cuda = params_local.is_cuda
    side = torch.cuda.Stream(device=params_local.device) if cuda else None
    if cuda and transport is not None:
        assert gather_dst == transport.dst, "the transport was built for destination rank %d" % transport.dst
        assert P == transport.P, "the transport was built for a shard of %d sets, got %d" % (transport.P, P)
        rows = transport.begin()
        evaluator.evaluate_device(params_local, rel=rel, out=rows, attrs=attrs)
        return rows, transport.finish()
    # shards of different size: agree on the largest, pad the pieces, trim on the destination
    sizes = torch.zeros(world, dtype=torch.int64, device=params_local.device)
    sizes[rank] = P
    dist.all_reduce(sizes, group=group)
    sizes = [int(v) for v in sizes.cpu()]
    P_max = max(sizes)

    def trimmed(g):
        return [g[r][:sizes[r]] for r in range(world)] if g is not None else None

    nE_shape = None
    if cuda and hook and hasattr(evaluator, "evaluate_device_pieces") and P == P_max and min(sizes) == P_max:
        # the library evaluates the shard in its own (large) chunks and reports every finished piece of the spectra:
        # its gather is issued on the side stream while the next piece is computed
        if out is None:
            out = torch.empty((P, evaluator.nE), dtype=torch.float64, device=params_local.device)
        if rank == gather_dst and gathered is None:
            gathered = [torch.empty_like(out) for _ in range(world)]
        works = []

        def on_piece(lo, n):
            done = torch.cuda.Event()
            done.record()
            with torch.cuda.stream(side):
                side.wait_event(done)
                dst_views = [g[lo:lo + n] for g in gathered] if rank == gather_dst else None
                works.append(dist.gather(out[lo:lo + n], dst_views, dst=gather_dst, group=group, async_op=True))

        evaluator.evaluate_device_pieces(params_local, on_piece, piece_sets=piece_sets, rel=rel, out=out, attrs=attrs)
        for w in works:
            w.wait()
        torch.cuda.current_stream(params_local.device).wait_stream(side)
        return out, gathered if rank == gather_dst else None
    edges = piece_edges(P_max, pieces)  # the same piece plan on every rank
    outs, works = [], []
    for lo, hi in zip(edges[:-1], edges[1:]):
        if hi <= lo:
            continue
        mine_hi = min(hi, P)  # rows of this piece the rank really owns
        n_mine = max(mine_hi - lo, 0)
        if out is not None and hi <= P:
            evaluator.evaluate_device(params_local[lo:hi], rel=rel, out=out[lo:hi],
                                      attrs=None if attrs is None else attrs[lo:hi])
            piece = out[lo:hi]
        else:
            if n_mine > 0:
                dst_out = out[lo:mine_hi] if out is not None else None
                got = evaluator.evaluate_device(params_local[lo:mine_hi], rel=rel, out=dst_out,
                                                attrs=None if attrs is None else attrs[lo:mine_hi])[0]
                if nE_shape is None:
                    nE_shape = tuple(got.shape[1:])
            if n_mine == hi - lo:
                piece = got
            else:  # short or empty piece of a smaller shard: pad with zero rows (trimmed on the destination)
                if nE_shape is None:
                    nE_shape = (evaluator.nE,)
                piece = torch.zeros((hi - lo,) + nE_shape, dtype=torch.float64, device=params_local.device)
                if n_mine > 0:
                    piece[:n_mine] = got
            if out is None and n_mine > 0:
                outs.append(got)
        if rank == gather_dst and gathered is None:
            gathered = [torch.empty((P_max,) + tuple(piece.shape[1:]), dtype=piece.dtype, device=piece.device)
                        for _ in range(world)]
        dst_views = [g[lo:hi] for g in gathered] if rank == gather_dst else None
        if cuda:
            done = torch.cuda.Event()
            done.record()
            with torch.cuda.stream(side):
                side.wait_event(done)
                works.append(dist.gather(piece, dst_views, dst=gather_dst, group=group, async_op=True))
        else:
            dist.gather(piece, dst_views, dst=gather_dst, group=group)
    for w in works:
        w.wait()
    if cuda:
        torch.cuda.current_stream(params_local.device).wait_stream(side)
    if out is None:
        out = outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)
    return out, trimmed(gathered) if rank == gather_dst else None


def fitstat_sharded(evaluator, params_local, rel=True, gather_dst=0, group=None):
    """Sharded fit statistic: every rank reduces its own shard to one double per set on the device; the final
    gather moves 8 B per set instead of 8 nE ("collapses the all-gather")."""
    import torch.distributed as dist
    stat = evaluator.fitstat_device(params_local, rel=rel)
    if gather_dst is None or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return stat, None
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    gathered = [torch.empty_like(stat) for _ in range(world)] if rank == gather_dst else None
    dist.gather(stat, gathered, dst=gather_dst, group=group)
    return stat, gathered
