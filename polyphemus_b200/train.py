"""Training-step plumbing around the hot path: losses, synthetic LMD-shaped batches, data-parallel step.

* ``vae_losses``     — same arithmetic as ``PolyphemusTrainer._losses`` (training.py:298-347), including its
                       quirks (the structure term is computed from ``s_tensor`` itself, training.py:307; the KLD
                       weight is the trainer's ``beta``, 0 as shipped, training.py:116), but returns tensors
                       instead of calling ``.item()`` seven times per step.
* ``synthetic_host_batch`` / ``device_batch`` — Bernoulli(p) structures and random note tokens in the on-disk
                       layout of the reference dataset (bool ``s_tensor``, int16 token ids, preprocess.py:210),
                       expanded on the device to the ``c_tensor`` one-hot layout of data.py:234-259.
* ``GradAllReducer`` — data-parallel gradient exchange: one flat fp32 buffer holds every ``.grad``; buckets are
                       all-reduced with NCCL as soon as backward has produced them.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import os

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import ops
from .graph import Graph, graphs_from_tensor
from .vae import MAX_SIMU_TOKENS, N_DUR_TOKENS, N_PITCH_TOKENS, LogitParts

PITCH_SOS, PITCH_EOS, PITCH_PAD = 128, 129, 130
DUR_SOS, DUR_EOS, DUR_PAD = 96, 97, 98


def vae_losses(s_tensor, s_logits, c_tensor, c_logits, mu, log_var, beta: float = 0.0, c_tokens=None):
    """Total loss and its parts (tensors). ``c_tokens`` int [N,16,2], when given, replaces the argmax over
    the one-hot ``c_tensor`` (same targets, no N x 15 x 230 read)."""
    lazy = isinstance(c_logits, LogitParts)
    parts = None if lazy else getattr(c_logits, "_parts", None)
    if lazy:
        pitch_logits = dur_logits = None
    elif parts is not None:
        pitch_logits, dur_logits = (p.reshape(-1, p.size(-1)).float() for p in parts)
    else:
        logits = c_logits.reshape(-1, c_logits.size(-1)).float()
        pitch_logits, dur_logits = logits[:, :N_PITCH_TOKENS], logits[:, N_PITCH_TOKENS:]
    if c_tokens is not None and lazy:
        tgt = c_tokens[:, 1:, :].int()                       # [N, 15, 2]: the fused NLL takes int32 targets
        pitch_true, dur_true = tgt[..., 0].reshape(-1), tgt[..., 1].reshape(-1)
    elif c_tokens is not None:
        tgt = c_tokens[:, 1:, :].reshape(-1, 2).long()
        pitch_true, dur_true = tgt[:, 0], tgt[:, 1]
    else:
        tgt = c_tensor[..., 1:, :].reshape(-1, c_tensor.size(-1))
        pitch_true = tgt[:, :N_PITCH_TOKENS].argmax(dim=1)
        dur_true = tgt[:, N_PITCH_TOKENS:].argmax(dim=1)
    # training.py:307 overwrites the structure logits with the structure tensor itself
    s_as_logits = s_tensor.reshape(-1, *s_logits.shape[2:]).float()
    s_loss = F.binary_cross_entropy_with_logits(s_as_logits.reshape(-1), s_tensor.reshape(-1).float())
    if lazy and c_logits.split is not None:
        # rows sorted drum-first, each block = [own pitch head | duration head]: the targets follow the same permutation
        out_d, out_o, perm, (wp, wu) = c_logits.split
        t = out_d.size(1)
        pitch_t = pitch_true.int().view(-1, t).index_select(0, perm)
        dur_t = dur_true.int().view(-1, t).index_select(0, perm)
        n_d = out_d.size(0)
        pitch_sum, dur_sum = 0.0, 0.0
        for out, sl in ((out_d, slice(0, n_d)), (out_o, slice(n_d, None))):
            if out.size(0) == 0:
                continue
            nll_p, nll_u = ops.token_nll_segments(out.reshape(-1, out.size(-1)),
                                                  [(0, wp, pitch_t[sl].reshape(-1), PITCH_PAD), (wp, wu, dur_t[sl].reshape(-1), DUR_PAD)])
            pitch_sum, dur_sum = pitch_sum + nll_p.sum(), dur_sum + nll_u.sum()
        pitch_loss = pitch_sum / (pitch_t != PITCH_PAD).sum()
        dur_loss = dur_sum / (dur_t != DUR_PAD).sum()
    elif lazy:
        # fused row-wise cross entropy on the three head outputs (pb_ce_fwd/bwd): the node's own pitch head is picked
        # by giving each head the targets with the other head's rows set to the ignored PAD id
        t = c_logits.drums.size(1)
        rows_drum = c_logits.is_drum.repeat_interleave(t)
        pitch_t, dur_t = pitch_true.int(), dur_true.int().contiguous()
        pad = torch.full_like(pitch_t, PITCH_PAD)
        flat = lambda x: x.reshape(-1, x.size(-1))
        t_drum, t_other = torch.where(rows_drum, pitch_t, pad), torch.where(rows_drum, pad, pitch_t)
        if c_logits.combined is not None:          # the heads are column blocks of one matrix: one gradient buffer
            w0, w1, w2 = c_logits.widths
            nll_drum, nll_other, nll_dur = ops.token_nll_segments(
                flat(c_logits.combined), [(0, w0, t_drum, PITCH_PAD), (w0, w1, t_other, PITCH_PAD), (w0 + w1, w2, dur_t, DUR_PAD)])
        else:
            nll_drum = ops.token_nll(flat(c_logits.drums), t_drum, PITCH_PAD)
            nll_other = ops.token_nll(flat(c_logits.others), t_other, PITCH_PAD)
            nll_dur = ops.token_nll(flat(c_logits.dur), dur_t, DUR_PAD)
        pitch_loss = (nll_drum.sum() + nll_other.sum()) / (pitch_t != PITCH_PAD).sum()
        dur_loss = nll_dur.sum() / (dur_t != DUR_PAD).sum()
    else:
        pitch_loss = _masked_ce(pitch_logits, pitch_true, PITCH_PAD)
        dur_loss = _masked_ce(dur_logits, dur_true, DUR_PAD)
    kld = (-0.5 * torch.sum(1 + log_var - mu.pow(2) - log_var.exp(), dim=1)).mean()
    total = pitch_loss + dur_loss + s_loss + beta * kld
    return total, {"pitch": pitch_loss, "dur": dur_loss, "structure": s_loss, "kld": kld}


def _masked_ce(logits, target: torch.Tensor, ignore_index: int) -> torch.Tensor:
    """nn.CrossEntropyLoss(ignore_index=...) (training.py:100-101): mean over the non-ignored rows of
    -log_softmax(logits)[target]. Written out because the library's nll_loss reduction runs in a single block and
    costs milliseconds on 2M rows."""
    keep = target != ignore_index
    nll = -F.log_softmax(logits, dim=1).gather(1, target.unsqueeze(1)).squeeze(1)
    return (nll * keep).sum() / keep.sum()


# ------------------------------------------------------------------------------------------ synthetic data
@dataclass
class HostBatch:
    """What a data loader would hand over: pinned host memory, either compacted (structure + the active cells'
    tokens) or — ``c_disk`` / ``s_disk`` — the samples exactly as preprocess.py:210 stores them on disk."""
    s_tensor: torch.Tensor      # bool  [B, n_bars, 4, 32]
    tokens: torch.Tensor        # int16 [N, 16, 2]  (pitch id, duration id) for every active (bar, track, t)
    c_disk: Optional[torch.Tensor] = None    # int16 [B, 4, T, 16, 2]   on-disk c_tensor of every sample
    s_disk: Optional[torch.Tensor] = None    # bool  [B, 4, T]          on-disk s_tensor of every sample

    @property
    def nbytes(self) -> int:
        if self.c_disk is not None:
            return self.c_disk.numel() * self.c_disk.element_size() + self.s_disk.numel() * self.s_disk.element_size()
        return self.s_tensor.numel() * self.s_tensor.element_size() + self.tokens.numel() * self.tokens.element_size()


def disk_layout(s_tensor: torch.Tensor, tokens: torch.Tensor):
    """(c_disk int16 [B, 4, T, 16, 2], s_disk bool [B, 4, T]) holding the given batch the way preprocess.py writes
    samples: silent cells carry [SOS, EOS, PAD, ...] (preprocess.py:120-147), active cells their tokens."""
    bsz, n_bars = s_tensor.shape[:2]
    s = s_tensor.bool()
    cells = torch.empty((bsz, n_bars, 4, 32, MAX_SIMU_TOKENS, 2), dtype=torch.int16)
    cells[..., 0], cells[..., 1] = PITCH_PAD, DUR_PAD
    cells[..., 0, 0], cells[..., 0, 1] = PITCH_SOS, DUR_SOS
    cells[..., 1, 0], cells[..., 1, 1] = PITCH_EOS, DUR_EOS
    cells[s] = tokens.to(torch.int16)
    c_disk = cells.permute(0, 2, 1, 3, 4, 5).reshape(bsz, 4, n_bars * 32, MAX_SIMU_TOKENS, 2).contiguous()
    s_disk = s.permute(0, 2, 1, 3).reshape(bsz, 4, n_bars * 32).contiguous()
    return c_disk, s_disk


def synthetic_tokens(num_nodes: int, generator: torch.Generator) -> torch.Tensor:
    """[N,16,2] token ids: SOS, k~U{1..14} notes, EOS, PAD... (constants.py:22-41, preprocess.py)."""
    k = torch.randint(1, 15, (num_nodes, 1), generator=generator)
    pos = torch.arange(MAX_SIMU_TOKENS).unsqueeze(0)
    pitch = torch.randint(0, 128, (num_nodes, MAX_SIMU_TOKENS), generator=generator)
    dur = torch.randint(0, 96, (num_nodes, MAX_SIMU_TOKENS), generator=generator)
    note, eos = (pos >= 1) & (pos <= k), pos == k + 1
    pitch = torch.where(note, pitch, torch.full_like(pitch, PITCH_PAD))
    dur = torch.where(note, dur, torch.full_like(dur, DUR_PAD))
    pitch = torch.where(eos, torch.full_like(pitch, PITCH_EOS), pitch)
    dur = torch.where(eos, torch.full_like(dur, DUR_EOS), dur)
    pitch[:, 0], dur[:, 0] = PITCH_SOS, DUR_SOS
    return torch.stack((pitch, dur), dim=-1).to(torch.int16)


def synthetic_host_batch(batch: int, n_bars: int, p: float = 0.25, seed: int = 0, pin: bool = True,
                         disk: bool = False) -> HostBatch:
    """``disk=True`` also lays the batch out as on-disk samples (the empty bars are then left empty in ``s_disk``: the
    fake activation is the decoder's job, data.py:152-153)."""
    rng = np.random.default_rng(seed)
    s = rng.random((batch, n_bars, 4, 32)) < p
    s_raw = s.copy()
    empty = ~s.reshape(batch, n_bars, -1).any(axis=-1)
    s[empty, 0, 0] = True                       # the dataset applies data.py:152-153 before filtering c_tensor
    n = int(s.sum())
    tokens = synthetic_tokens(n, torch.Generator().manual_seed(seed))
    s_t = torch.from_numpy(s)
    c_disk = s_disk = None
    if disk:
        c_disk, _ = disk_layout(s_t, tokens)
        s_disk = torch.from_numpy(s_raw).permute(0, 2, 1, 3).reshape(batch, 4, n_bars * 32).contiguous()
    if pin and torch.cuda.is_available():
        s_t, tokens = s_t.pin_memory(), tokens.pin_memory()
        if disk:
            c_disk, s_disk = c_disk.pin_memory(), s_disk.pin_memory()
    return HostBatch(s_t, tokens, c_disk, s_disk)


def onehot_content(tokens: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """int [N,16,2] -> float [N,16,230] one-hot pitch | duration (data.py:234-259), on the tokens' device."""
    n, t = tokens.shape[:2]
    out = torch.zeros((n, t, N_PITCH_TOKENS + N_DUR_TOKENS), dtype=dtype, device=tokens.device)
    idx = tokens.long()
    out.scatter_(2, idx[..., :1], 1.0)
    out.scatter_(2, idx[..., 1:] + N_PITCH_TOKENS, 1.0)
    return out


def device_batch(host: HostBatch, device, onehot: bool = False, onehot_dtype=torch.float32) -> Graph:
    """Host batch -> device graph with ``s_tensor`` and ``c_tokens`` attached (H2D copies + device graph
    build). ``onehot=True`` also expands the reference's ``c_tensor`` one-hot layout on the device."""
    if host.c_disk is not None:          # samples as stored on disk: bars reshape, silence filter etc. on the device
        from .data import decode_samples
        return decode_samples(host.c_disk, host.s_disk, int(host.s_tensor.size(1)), device=device, onehot=onehot)
    s_dev = host.s_tensor.to(device, non_blocking=True)
    tok_dev = host.tokens.to(device, non_blocking=True)
    graph = graphs_from_tensor(s_dev)
    if graph.num_nodes != tok_dev.size(0):
        raise ValueError(f"{tok_dev.size(0)} token rows for {graph.num_nodes} nodes")
    graph.s_tensor = s_dev.view(-1, 4, 32).float()
    graph.c_tokens = tok_dev
    graph.c_tensor = onehot_content(tok_dev, onehot_dtype) if onehot else None
    return graph


class BatchPrefetcher:
    """Builds the next batch's device graph on a side CUDA stream while the current step runs.

    `device_batch` ends its counting pass with the one host read-back that sizes the graph. On the training stream that
    read-back waits for everything queued before it — the whole previous step — so the host can never run ahead of
    the device and every host-bound stretch of a step shows up as device idle time. On a side stream it only waits
    for the (sub-millisecond) counting kernels; the reference overlaps graph construction with training the same way,
    through DataLoader worker processes (train.py:152-156). `take()` makes the training stream wait for the build and
    hands the tensors' memory over to it."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)

    def submit(self, host, **kw):
        """`host`: a HostBatch, or a callable returning one (run under the side stream — e.g. a clone of tensors that
        are already resident). Its tensors must not depend on work still queued on the training stream."""
        with torch.cuda.stream(self.stream):
            graph = device_batch(host() if callable(host) else host, self.device, **kw)
            ready = torch.cuda.Event()
            ready.record()
        return graph, ready

    def take(self, pending) -> Graph:
        graph, ready = pending
        main = torch.cuda.current_stream(self.device)
        main.wait_event(ready)
        for v in vars(graph).values():             # allocated on the side stream, consumed (and freed) on this one
            if isinstance(v, torch.Tensor) and v.is_cuda:
                v.record_stream(main)
        return graph


# ------------------------------------------------------------------------------------------ data parallel
class GradAllReducer:
    """Flat-buffer gradient all-reduce (NCCL average) overlapped with backward.

    Parameters are bucketed in reverse registration order (the order backward produces them) over one contiguous
    fp32 buffer. Backward leaves every gradient where autograd puts it (no per-parameter accumulation kernels); a
    post-accumulate hook counts ready parameters, and the moment bucket b is complete AND buckets 0..b-1 have been
    launched, its gradients are packed into the buffer with ONE multi-tensor copy and ``all_reduce(AVG)`` is launched on
    that slice — strictly in bucket order, so every rank issues the identical collective sequence even when a rank is
    missing a gradient (e.g. a shard without drum nodes). ``finish()`` flushes the remaining buckets (parameters that
    never receive a gradient — the 12 ``decoder.s_decoder.*`` tensors, whose loss term is constant, training.py:307 —
    keep a zero slot, identically on every rank), waits, and points every ``.grad`` at its slice of the buffer for
    the optimizer. The buckets produced last are small (``tail_mb``) so that the exposed tail after backward is short.

    A parameter that never gets a gradient would hold back its bucket — and, the launches being ordered, every bucket
    behind it — until ``finish()``: the whole exchange would sit exposed after backward. The first ``finish()`` therefore
    agrees across ranks (one MAX all-reduce of a flag vector) on the parameters NO rank produced a gradient for; from
    then on they count as ready from the start of every step. Should one of them receive a gradient later it is
    packed if its bucket has not gone out yet, counted normally from the next step on, and reported once.
    """

    def __init__(self, params, bucket_mb: float = 32.0, process_group=None, tail_mb: float = 4.0):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        dev = self.params[0].device
        self.sync = True             # False: no exchange (single-rank checks); finish() still packs the buffer
        self.flat = None
        self._hook_handles = []
        if self.world == 1:          # nothing to exchange: plain per-parameter gradients, dropped between steps
            return
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.buckets = []            # (start, end, [params])
        self._bucket_of = {}
        self._view = {}
        limit = int(bucket_mb * (1 << 20) / 4)
        tail = min(int(tail_mb * (1 << 20) / 4), limit)
        off, start, members = 0, 0, []
        for p in reversed(self.params):
            n = p.numel()
            self._view[p] = self.flat[off:off + n].view_as(p)
            self._bucket_of[p] = len(self.buckets)
            members.append(p)
            off += n
            # the last `2 * tail` elements of the buffer (the gradients backward produces last) go in small buckets
            cap = tail if total - off < 2 * tail else limit
            if off - start >= cap:
                self.buckets.append((start, off, members))
                start, members = off, []
        if members:
            self.buckets.append((start, off, members))
        self._absent = set()                       # parameters without a gradient on every rank (agreed at the first step)
        self._absent_count = [0] * len(self.buckets)
        self._calibrated = False
        self._seen = set()
        self._ready = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self._next = 0
        self._handles = []
        self._avg = dist.ReduceOp.AVG if dist.get_backend(process_group) == "nccl" else None
        for p in self.params:
            self._hook_handles.append(p.register_post_accumulate_grad_hook(self._hook))

    def close(self) -> None:
        """Remove the autograd hooks (a second reducer on the same parameters must not double-count)."""
        for h in self._hook_handles:
            h.remove()
        self._hook_handles = []

    def _pack(self, b: int) -> None:
        """Gradients of bucket b -> their slices of the flat buffer (one multi-tensor copy; missing ones stay zero)."""
        have = [p for p in self.buckets[b][2] if p.grad is not None and p.grad.data_ptr() != self._view[p].data_ptr()]
        if have:
            torch._foreach_copy_([self._view[p] for p in have], [p.grad for p in have])

    def _launch(self, b: int) -> None:
        if self._launched[b]:
            return
        self._launched[b] = True
        self._pack(b)
        if self.sync:
            s, e, _ = self.buckets[b]
            if self._avg is not None:
                op = self._avg
            else:                                # gloo has no AVG: pre-scale, then sum
                self.flat[s:e].mul_(1.0 / self.world)
                op = dist.ReduceOp.SUM
            self._handles.append(dist.all_reduce(self.flat[s:e], op=op, group=self.group, async_op=True))

    def _advance(self) -> None:
        """Launch complete buckets strictly in index order."""
        while self._next < len(self.buckets) and self._ready[self._next] >= len(self.buckets[self._next][2]):
            self._launch(self._next)
            self._next += 1

    def _hook(self, p) -> None:
        if not self.sync:
            return
        if not self._calibrated:
            self._seen.add(p)
        if p in self._absent:                       # counted as ready already; count it normally from the next step on
            self._absent.discard(p)
            self._absent_count[self._bucket_of[p]] -= 1
            import warnings
            warnings.warn("GradAllReducer: a parameter without gradient in the first step received one later; if its "
                          "bucket had already been reduced this step's contribution of it is dropped on this rank")
            return
        self._ready[self._bucket_of[p]] += 1
        self._advance()

    def _calibrate(self) -> None:
        """Agree on the parameters nobody produced a gradient for (first step only: one small all-reduce, one read-back)."""
        flags = torch.tensor([1.0 if p in self._seen else 0.0 for p in self.params], device=self.flat.device)
        dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=self.group)
        for p, f in zip(self.params, flags.tolist()):
            if f == 0.0:
                self._absent.add(p)
                self._absent_count[self._bucket_of[p]] += 1
        self._calibrated, self._seen = True, set()

    def zero_grad(self) -> None:
        for p in self.params:
            p.grad = None
        if self.flat is None:
            return
        self.flat.zero_()
        self._ready = list(self._absent_count)
        self._launched = [False] * len(self.buckets)
        self._next = 0
        self.launched_in_backward = 0

    def finish(self) -> None:
        """Flush incomplete buckets in order, wait for all reductions, expose the buffer slices as ``.grad``."""
        if self.flat is None:
            return
        self.launched_in_backward = self._next       # buckets that went out while backward was still running
        for b in range(len(self.buckets)):
            self._launch(b)
        self._next = len(self.buckets)
        if self.sync and not self._calibrated:
            self._calibrate()
        for h in self._handles:
            h.wait()
        self._handles = []
        for p in self.params:
            p.grad = self._view[p]

    def sync_buffers(self, model) -> None:
        """Average the floating-point buffers (BatchNorm running statistics are per replica during training) across
        ranks, e.g. before a checkpoint is written from rank 0."""
        if self.world == 1:
            return
        bufs = [b for b in model.buffers() if b.is_floating_point()]
        if not bufs:
            return
        flat = torch.cat([b.reshape(-1).float() for b in bufs])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.div_(self.world)
        off = 0
        for b in bufs:
            b.copy_(flat[off:off + b.numel()].view_as(b))
            off += b.numel()


class TrainStep:
    """graph build -> forward -> loss -> backward -> gradient all-reduce -> Adam (train.py:176-207 config).

    While a TrainStep is alive the content decoder returns the head outputs (``LogitParts``) instead of assembling the
    dense ``[N, 15, 230]`` logits — the step only needs the loss; ``close()`` restores the reference behaviour of
    ``model(graph)`` and removes the gradient hooks."""

    def __init__(self, model, lr: float = 1e-4, betas=(0.9, 0.98), eps: float = 1e-9, autocast_bf16: bool = False,
                 beta_kld: float = 0.0, bucket_mb: float = 32.0):
        self.model = model
        bucket_mb = float(os.environ.get("PB200_BUCKET_MB", bucket_mb))
        self.reducer = GradAllReducer(model.parameters(), bucket_mb=bucket_mb)
        self.opt = torch.optim.Adam(model.parameters(), lr=lr, betas=betas, eps=eps, fused=model_is_cuda(model))
        self.autocast_bf16 = autocast_bf16
        self.beta_kld = beta_kld
        self._lazy = [(m, m.materialize_logits) for m in model.modules() if hasattr(m, "materialize_logits")]
        for m, _ in self._lazy:
            m.materialize_logits = False

    def close(self) -> None:
        for m, was in self._lazy:
            m.materialize_logits = was
        self._lazy = []
        self.reducer.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __call__(self, graph: Graph, noise: Optional[torch.Tensor] = None):
        self.reducer.zero_grad()
        dev_type = next(self.model.parameters()).device.type
        with torch.autocast(device_type=dev_type, dtype=torch.bfloat16, enabled=self.autocast_bf16):
            (s_logits, c_logits), mu, log_var = self.model(graph, noise=noise)
            loss, parts = vae_losses(graph.s_tensor, s_logits, graph.c_tensor, c_logits, mu, log_var,
                                     beta=self.beta_kld, c_tokens=getattr(graph, "c_tokens", None))
        loss.backward()
        self.reducer.finish()
        self.opt.step()
        return loss.detach(), parts


def model_is_cuda(model) -> bool:
    return next(model.parameters()).is_cuda
