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
             or a dict(spins, thetas, NR, NPHI, dl, data) to upload.
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
            self.ctx.tables_upload(t["spins"], t["thetas"], t["NR"], t["NPHI"], t["dl"], t["data"])
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
    def evaluate(self, params, rel=True, components=False, fp32=False):
        """params: array-like [P, n] on the host.  Returns (spectra ndarray [P, nE] or [P, 4, nE] in W/Hz,
        attrs ndarray [P, 24]).  Stages through pinned memory; H2D and D2H happen inside this call."""
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
        pin = self._pin_in[:P]
        pin.zero_()
        pin[:, :p.shape[1]] = torch.from_numpy(p)
        out = self._pin_out[:P * ncomp * self.nE].view((P, 4, self.nE) if components else (P, self.nE))
        out_np, at = self.ctx.eval_batch_host(pin.numpy(), model=self.model, rel=rel, components=components,
                                              dr_dex=self.dr_dex, out=out.numpy(), fp32=fp32)
        return out_np, at


def shard_bounds(P, world, rank):
    """Contiguous split of P parameter sets over `world` ranks."""
    base, rem = divmod(P, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def evaluate_sharded(evaluator, params_local, rel=True, gather_dst=0, group=None):
    """One process per GPU: every rank evaluates its own shard (device tensor [P_local, 15]); the only collective
    is the final gather of the spectra to `gather_dst` over NCCL (gather_dst=None: no gather).
    Returns (local_out, gathered or None)."""
    import torch.distributed as dist
    out, _ = evaluator.evaluate_device(params_local, rel=rel)
    if gather_dst is None or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return out, None
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    gathered = None
    if rank == gather_dst:
        gathered = [torch.empty_like(out) for _ in range(world)]
    dist.gather(out, gathered, dst=gather_dst, group=group)
    return out, gathered
