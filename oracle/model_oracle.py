"""Functional PyTorch-CPU restatement of the Polyphemus VAE with the message-passing layers spelled out.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py). Every function works on a plain ``state_dict``
(reference key names, any float dtype — fp32 for parity, fp64 for error budgeting) and on explicit
tensors, so that it is usable as an autograd-differentiable checker for both the layer kernels and the
whole training step. Reference lines followed:

* ``gcl_forward``        <- ``GCL.forward`` model.py:55-121 (branch model.py:101-112), ``GCL.message``
                            model.py:123-135, PyG propagate/scatter-mean, with the
                            per-relation boolean compaction ``masked_edge_index/attrs`` model.py:30-38
* ``gcn_forward``        <- ``GCN.forward`` model.py:190-208 (dropout p = config, GCL, BatchNorm, ReLU, residual)
* ``content_encoder``    <- model.py:344-417     ``structure_encoder`` <- model.py:434-445, 251-256
* ``encoder``            <- model.py:466-483     ``structure_decoder`` <- model.py:500-505, 294-299
* ``content_decoder``    <- model.py:536-578     ``decoder`` <- model.py:634-655  ``vae`` <- model.py:665-678
* ``losses``             <- ``PolyphemusTrainer._losses`` training.py:298-347 (including the overwrite of
                            ``s_logits`` by ``s_tensor`` at training.py:307 and beta = 0, training.py:116)

Only the configuration the reference publishes is restated: ``batch_norm=True``, config ``dropout=0``
(training.json:3-5). The GCL-internal dropout (hard-wired p=0.1, model.py:44,133) is either disabled
(``gcl_dropout=0``) or driven by an explicit keep-mask so that a counter-based device RNG can be checked
exactly.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor

N_PITCH = 131          # constants.py:28
N_DUR = 99             # constants.py:40
MAX_SIMU = 16          # constants.py:48
PITCH_PAD = 130        # constants.py:22-25
DUR_PAD = 98           # constants.py:34-37
N_REL = 6              # constants.py:58
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


@dataclass
class Ctx:
    """Mode + side outputs (new BatchNorm running statistics keyed by state-dict prefix)."""
    training: bool = True
    gcl_dropout: float = 0.0
    gcl_keep_masks: Optional[dict] = None      # {(gcn_prefix, layer): bool/float [E, d] keep-mask}
    gcl_random_dropout: bool = False           # no mask given: draw the GCL dropout with F.dropout (as model.py:133)
    emulate_bf16: bool = False                 # round where the bf16 mode of the CUDA path rounds (see _rb)
    running: dict = field(default_factory=dict)


def _lin(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def _bn(sd, prefix, x, ctx: Ctx):
    """BatchNorm1d/2d, eps 1e-5, momentum 0.1; records the running-stat update when training."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if not ctx.training:
        return F.batch_norm(x, rm, rv, w, b, False, BN_MOMENTUM, BN_EPS)
    prev_rm, prev_rv = ctx.running.get(prefix, (rm, rv))       # a module applied twice chains its updates
    new_rm, new_rv = prev_rm.detach().clone(), prev_rv.detach().clone()
    y = F.batch_norm(x, new_rm, new_rv, w, b, True, BN_MOMENTUM, BN_EPS)
    ctx.running[prefix] = (new_rm, new_rv)
    return y


def _rb(t: Tensor) -> Tensor:
    """Round to bf16 and back (straight-through gradient). NOT part of the reference: it places the roundings of the
    CUDA path's bf16 mode (GEMM operands [H_r | x] and weights in bf16, fp32 accumulation, the
    pre-BatchNorm output and the layer output stored in bf16) into the reference's arithmetic, so that a bf16-mode
    result can be checked to rounding level — against the fp32 arithmetic the two differ by ReLU decisions that bf16
    round-off flips, which says nothing about the implementation."""
    return t + (t.to(torch.bfloat16).to(t.dtype) - t).detach()


# ----------------------------------------------------------------------------- message passing
def gcl_forward(x: Tensor, edge_index: Tensor, edge_type: Tensor, edge_dist: Tensor,
                weight: Tensor, root: Tensor, bias: Tensor, nn_w: Tensor, nn_b: Tensor,
                keep_mask: Optional[Tensor] = None, p_drop: float = 0.0, random_dropout: bool = False,
                emulate_bf16: bool = False) -> Tensor:
    """One relational graph-conv layer. edge_dist = argmax of the one-hot edge_attr (model.py:194).

    out[v] = sum_r mean_{e: dst=v, type=r} dropout(relu(x[src_e] * (A[:, dist_e] + a))) @ W_r + x[v] @ root + bias
    """
    n, d_in = x.shape
    out = torch.zeros(n, weight.shape[2], dtype=x.dtype)
    onehot = F.one_hot(edge_dist, nn_w.shape[1]).to(x.dtype)
    for r in range(weight.shape[0]):
        sel = edge_type == r                                   # model.py:104-105
        src, dst = edge_index[0, sel], edge_index[1, sel]
        gate = F.linear(onehot[sel], nn_w, nn_b)[:, :d_in]     # model.py:127-129
        msg = F.relu(x.index_select(0, src) * gate)            # model.py:131-132
        if keep_mask is not None and p_drop > 0:
            msg = msg * keep_mask[sel].to(x.dtype) / (1.0 - p_drop)   # model.py:133
        elif random_dropout and p_drop > 0:
            msg = F.dropout(msg, p=p_drop, training=True)      # model.py:133 as shipped
        agg = torch.zeros(n, d_in, dtype=x.dtype).index_add_(0, dst, msg)
        cnt = torch.zeros(n, dtype=x.dtype).index_add_(0, dst, torch.ones(dst.shape[0], dtype=x.dtype))
        agg = agg / cnt.clamp(min=1).unsqueeze(1)              # scatter-mean
        if emulate_bf16:
            out = out + _rb(agg) @ _rb(weight[r])
        else:
            out = out + agg @ weight[r]                        # model.py:112
    if emulate_bf16:
        return _rb(out + _rb(x) @ _rb(root) + bias)
    out = out + x @ root                                       # model.py:116
    return out + bias                                          # model.py:119


def gcn_forward(sd, prefix: str, x: Tensor, edge_index: Tensor, edge_type: Tensor, edge_dist: Tensor,
                ctx: Ctx, n_layers: Optional[int] = None) -> Tensor:
    """The GCL stack with BatchNorm/ReLU/residual (model.py:196-206). Config dropout is 0."""
    if n_layers is None:
        n_layers = 1 + max(int(k[len(prefix) + 8:].split(".")[0]) for k in sd if k.startswith(prefix + ".layers."))
    if ctx.emulate_bf16:
        x = _rb(x)
    for i in range(n_layers):
        lp = f"{prefix}.layers.{i}"
        keep = None if ctx.gcl_keep_masks is None else ctx.gcl_keep_masks.get((prefix, i))
        h = gcl_forward(x, edge_index, edge_type, edge_dist, sd[lp + ".weight"], sd[lp + ".root"],
                        sd[lp + ".bias"], sd[lp + ".nn.weight"], sd[lp + ".nn.bias"],
                        keep_mask=keep if ctx.training else None, p_drop=ctx.gcl_dropout if ctx.training else 0.0,
                        random_dropout=ctx.gcl_random_dropout and ctx.training, emulate_bf16=ctx.emulate_bf16)
        if f"{prefix}.norm_layers.{i}.module.weight" in sd:
            h = _bn(sd, f"{prefix}.norm_layers.{i}.module", h, ctx)
        x = x + F.relu(h)
        if ctx.emulate_bf16:
            x = _rb(x)
    return x


# ----------------------------------------------------------------------------- encoder side
def structure_encoder(sd, p, s_tensor: Tensor, n_bars: int, d: int, ctx: Ctx) -> Tensor:
    h = s_tensor.reshape(-1, 1, 4, s_tensor.shape[-1])
    c = p + ".cnn_encoder"
    h = F.conv2d(h, sd[c + ".conv.0.weight"], sd[c + ".conv.0.bias"], padding=1)
    h = F.relu(_bn(sd, c + ".conv.1", h, ctx))
    h = F.max_pool2d(h, (1, 4), stride=(1, 4))
    h = F.conv2d(h, sd[c + ".conv.4.weight"], sd[c + ".conv.4.bias"], padding=1)
    h = F.relu(_bn(sd, c + ".conv.5", h, ctx)).flatten(1)
    h = _lin(sd, c + ".lin.4", F.relu(_lin(sd, c + ".lin.1", h)))
    return _lin(sd, p + ".bars_encoder", h.reshape(-1, n_bars * d))


def global_attention(sd, p, x: Tensor, seg: Tensor, n_seg: int, ctx: Ctx) -> Tensor:
    """PyG GlobalAttention with gate_nn = Sequential(MLP(d->1, 1 layer, no act), BatchNorm1d(1)) (model.py:335-340)."""
    gate = _lin(sd, p + ".gate_nn.0.layers.0", x)
    gate = _bn(sd, p + ".gate_nn.1", gate, ctx).view(-1, 1)
    seg_max = torch.full((n_seg, 1), float("-inf"), dtype=x.dtype).scatter_reduce(
        0, seg.view(-1, 1), gate.detach(), reduce="amax", include_self=True)
    e = (gate - seg_max.index_select(0, seg)).exp()
    denom = torch.zeros(n_seg, 1, dtype=x.dtype).index_add_(0, seg, e).index_select(0, seg) + 1e-16
    return torch.zeros(n_seg, x.shape[1], dtype=x.dtype).index_add_(0, seg, (e / denom) * x)


def content_encoder(sd, p, g, n_bars: int, d: int, ctx: Ctx) -> Tensor:
    c = g.c_tensor[:, 1:, :]                                   # drop SOS, model.py:349
    is_drum = g.is_drum
    out = torch.zeros(c.shape[0], d, dtype=c.dtype)
    for mask, pitch_emb, pitch_bn in ((is_drum, "drums_pitch_emb", "bn_drums"),
                                      (~is_drum, "non_drums_pitch_emb", "bn_non_drums")):
        part = c[mask]
        k, t = part.shape[0], part.shape[1]
        pe = _bn(sd, f"{p}.{pitch_bn}", _lin(sd, f"{p}.{pitch_emb}", part[..., :N_PITCH]).reshape(-1, d // 2), ctx)
        de = _bn(sd, f"{p}.bn_dur", _lin(sd, f"{p}.dur_emb", part[..., N_PITCH:]).reshape(-1, d // 2), ctx)
        tok = torch.cat((pe.view(k, t, d // 2), de.view(k, t, d // 2)), dim=-1)
        out = out.index_put((mask.nonzero(as_tuple=True)[0],),
                            F.relu(_lin(sd, f"{p}.chord_encoder", tok.reshape(k, t * d))))
    seg = g.bars + n_bars * g.batch                             # model.py:403
    h = gcn_forward(sd, p + ".graph_encoder", out, g.edge_index, g.edge_type, g.edge_dist, ctx)
    n_seg = int(g.batch.max()) * n_bars + n_bars
    pooled = global_attention(sd, p + ".graph_attention", h, seg, n_seg, ctx)
    return _lin(sd, p + ".bars_encoder", pooled.reshape(-1, n_bars * d))


def encoder(sd, g, n_bars: int, d: int, ctx: Ctx):
    z_s = structure_encoder(sd, "encoder.s_encoder", g.s_tensor, n_bars, d, ctx)
    z_c = content_encoder(sd, "encoder.c_encoder", g, n_bars, d, ctx)
    z = _lin(sd, "encoder.linear_merge", torch.cat((z_c, z_s), dim=1))
    z = F.relu(_bn(sd, "encoder.bn_linear_merge", z, ctx))
    return _lin(sd, "encoder.linear_mu", z), _lin(sd, "encoder.linear_log_var", z)


# ----------------------------------------------------------------------------- decoder side
def structure_decoder(sd, p, z_s: Tensor, n_bars: int, d: int, ctx: Ctx) -> Tensor:
    c = p + ".cnn_decoder"
    h = _lin(sd, p + ".bars_decoder", z_s).reshape(-1, d)
    h = F.relu(_lin(sd, c + ".lin.4", F.relu(_lin(sd, c + ".lin.1", h)))).reshape(-1, 16, 4, 8)
    h = F.interpolate(h, scale_factor=(1, 4), mode="nearest")
    h = F.conv2d(h, sd[c + ".conv.1.weight"], sd[c + ".conv.1.bias"], padding=1)
    h = F.relu(_bn(sd, c + ".conv.2", h, ctx))
    h = F.conv2d(h, sd[c + ".conv.4.weight"], sd[c + ".conv.4.bias"], padding=1)
    return h.reshape(z_s.shape[0], n_bars, 4, -1)


def content_decoder(sd, p, z_c: Tensor, g, n_bars: int, d: int, ctx: Ctx) -> Tensor:
    seg = g.bars + n_bars * g.batch                             # model.py:542
    x0 = _lin(sd, p + ".bars_decoder", z_c).reshape(-1, d).index_select(0, seg)   # == repeat_interleave by counts
    h = gcn_forward(sd, p + ".graph_decoder", x0, g.edge_index, g.edge_type, g.edge_dist, ctx)
    h = _lin(sd, p + ".chord_decoder", h).reshape(-1, MAX_SIMU - 1, d)
    out = torch.zeros(h.shape[0], MAX_SIMU - 1, N_PITCH + N_DUR, dtype=h.dtype)
    for mask, pitch_emb in ((g.is_drum, "drums_pitch_emb"), (~g.is_drum, "non_drums_pitch_emb")):
        part = h[mask]
        logits = torch.cat((_lin(sd, f"{p}.{pitch_emb}", part[..., : d // 2]),
                            _lin(sd, f"{p}.dur_emb", part[..., d // 2:])), dim=-1)
        out = out.index_put((mask.nonzero(as_tuple=True)[0],), logits)
    return out


def decoder(sd, z: Tensor, g, n_bars: int, d: int, ctx: Ctx):
    h = F.relu(_bn(sd, "decoder.batch_norm", _lin(sd, "decoder.lin_decoder", z), ctx))
    s_logits = structure_decoder(sd, "decoder.s_decoder", h[:, :d], n_bars, d, ctx)
    c_logits = content_decoder(sd, "decoder.c_decoder", h[:, d:], g, n_bars, d, ctx)
    return s_logits, c_logits


def vae(sd, g, n_bars: int, d: int, ctx: Ctx, eps_noise: Optional[Tensor] = None):
    """``VAE.forward`` (model.py:665-678). ``eps_noise`` replaces ``torch.randn_like`` so both sides share it."""
    mu, log_var = encoder(sd, g, n_bars, d, ctx)
    noise = torch.randn_like(mu) if eps_noise is None else eps_noise
    z = torch.exp(0.5 * log_var) * noise + mu
    s_logits, c_logits = decoder(sd, z, g, n_bars, d, ctx)
    return (s_logits, c_logits), mu, log_var


# ----------------------------------------------------------------------------- loss
def losses(s_tensor: Tensor, s_logits: Tensor, c_tensor: Tensor, c_logits: Tensor, mu: Tensor,
           log_var: Tensor, beta: float = 0.0):
    """``_losses`` (training.py:298-347). NB the structure term uses s_tensor as its own logits (training.py:307)."""
    tgt = c_tensor[..., 1:, :].reshape(-1, c_tensor.shape[-1])
    logits = c_logits.reshape(-1, c_logits.shape[-1])
    s_as_logits = s_tensor.reshape(-1, *s_logits.shape[2:])
    s_loss = F.binary_cross_entropy_with_logits(s_as_logits.reshape(-1), s_tensor.reshape(-1).to(s_as_logits.dtype))
    pitch = F.cross_entropy(logits[:, :N_PITCH], tgt[:, :N_PITCH].argmax(dim=1), ignore_index=PITCH_PAD)
    dur = F.cross_entropy(logits[:, N_PITCH:], tgt[:, N_PITCH:].argmax(dim=1), ignore_index=DUR_PAD)
    kld = (-0.5 * torch.sum(1 + log_var - mu.pow(2) - log_var.exp(), dim=1)).mean()
    total = pitch + dur + s_loss + beta * kld
    return total, {"pitch": pitch, "dur": dur, "structure": s_loss, "kld": kld}


# ----------------------------------------------------------------------------- synthetic content
@dataclass
class GraphBatch:
    """Tensor view of oracle.graph_oracle.GraphArrays plus content, as the oracle functions expect."""
    edge_index: Tensor
    edge_type: Tensor
    edge_dist: Tensor
    is_drum: Tensor
    bars: Tensor
    batch: Tensor
    num_nodes: int
    s_tensor: Tensor            # float [B*n_bars, 4, 32]
    c_tensor: Optional[Tensor]  # float [N, 16, 230] one-hot


def synthetic_tokens(num_nodes: int, seed: int = 0) -> torch.Tensor:
    """int64 [N, 16, 2] (pitch id, duration id): SOS, k notes, EOS, PAD...  (constants.py:22-41)."""
    gen = torch.Generator().manual_seed(seed)
    k = torch.randint(1, 15, (num_nodes,), generator=gen)
    pos = torch.arange(MAX_SIMU).unsqueeze(0)
    pitch = torch.randint(0, 128, (num_nodes, MAX_SIMU), generator=gen)
    dur = torch.randint(0, 96, (num_nodes, MAX_SIMU), generator=gen)
    is_note = (pos >= 1) & (pos <= k.unsqueeze(1))
    is_eos = pos == (k.unsqueeze(1) + 1)
    pitch = torch.where(is_note, pitch, torch.full_like(pitch, PITCH_PAD))
    dur = torch.where(is_note, dur, torch.full_like(dur, DUR_PAD))
    pitch = torch.where(is_eos, torch.full_like(pitch, 129), pitch)
    dur = torch.where(is_eos, torch.full_like(dur, 97), dur)
    pitch[:, 0], dur[:, 0] = 128, 96
    return torch.stack((pitch, dur), dim=-1)


def onehot_content(tokens: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """[N, 16, 2] ids -> [N, 16, 230] one-hot pitch | duration (data.py:234-259)."""
    return torch.cat((F.one_hot(tokens[..., 0], N_PITCH), F.one_hot(tokens[..., 1], N_DUR)), dim=-1).to(dtype)


def make_batch(arrays, tokens: Optional[torch.Tensor] = None, dtype=torch.float32) -> GraphBatch:
    s = torch.from_numpy(arrays.s_tensor).reshape(-1, 4, arrays.s_tensor.shape[-1]).to(dtype)
    return GraphBatch(
        edge_index=torch.from_numpy(arrays.edge_index), edge_type=torch.from_numpy(arrays.edge_type),
        edge_dist=torch.from_numpy(arrays.edge_dist), is_drum=torch.from_numpy(arrays.is_drum),
        bars=torch.from_numpy(arrays.bars), batch=torch.from_numpy(arrays.batch), num_nodes=arrays.num_nodes,
        s_tensor=s, c_tensor=None if tokens is None else onehot_content(tokens, dtype))


def leaf_state(sd: dict, dtype=torch.float32) -> dict:
    """Clone a state_dict into autograd leaves, re-tying the edge network that all layers of one GCN share
    (``edge_nn`` is a single nn.Linear handed to every GCL, model.py:175,178,183): ``layers.{i}.nn.*`` alias
    ``layers.0.nn.*`` so that its gradient accumulates over the layers exactly as in the reference."""
    out = {}
    for k, v in sd.items():
        if ".layers." in k and ".nn." in k:
            head, tail = k.split(".layers.")
            k0 = f"{head}.layers.0.{tail.split('.', 1)[1]}"
            if k0 in out:
                out[k] = out[k0]
                continue
        if v.is_floating_point():
            t = v.detach().to(dtype).clone()
            out[k] = t.requires_grad_(True) if "running_" not in k else t
        else:
            out[k] = v.clone()
    return out
