"""The Polyphemus graph VAE around the CUDA message-passing path (host side, plain PyTorch).

Only the two ``GCN`` stacks and the graph construction are the hot path; everything in this
file is the *caller* of that path and stays ordinary PyTorch. It keeps the reference's drop-in surface
(model.py:138-678): constructor kwargs (``VAE(**training.json["model"], device=...)``), forward signatures,
module/attribute names and therefore state-dict keys (255 at the published config, checked against
tests/golden/state_dict_keys.json), and the order in which parameters are created and re-initialised, so a
given ``torch.manual_seed`` produces the initial weights of the reference run on the PyG stand-in (oracle/pyg_shim.py).
Whether that equals a real torch-geometric 2.0.2 run depends on one recalled detail that cannot be checked here:
``inits.reset`` is taken to recurse into ``gate_nn``'s MLP; if 2.0.2 only resets direct children, the gate MLP keeps its
constructor draw and the random stream after it shifts. Trained checkpoints are unaffected (state-dict keys and shapes
are what they load by).

Differences from the reference are confined to *how* the same values are produced:
  * boolean-mask indexing (a host sync per mask, model.py:352-353,396-397,552-553,575-576) is replaced by
    index tensors that the device graph builder already knows the sizes of;
  * ``torch.unique`` + ``repeat_interleave`` (model.py:543-545) becomes one ``index_select``;
  * graph building inside ``Decoder`` (model.py:596-607) is one batched kernel call instead of a Python
    loop over sequences and bars.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import _ffi, ops
from .conv import GCN, _reset
from .graph import Graph, graphs_from_tensor

N_TRACKS = 4
N_PITCH_TOKENS = 131
N_DUR_TOKENS = 99
D_TOKEN_PAIR = N_PITCH_TOKENS + N_DUR_TOKENS
MAX_SIMU_TOKENS = 16
N_EDGE_TYPES = N_TRACKS + 2


class MLP(nn.Module):
    def __init__(self, input_dim=256, hidden_dim=256, output_dim=256, num_layers=2, activation=True, dropout=0.1):
        super().__init__()
        widths = [input_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(widths[:-1], widths[1:]))
        self.activation = activation
        self.p = dropout

    def forward(self, x):
        for lin in self.layers:
            x = lin(F.dropout(x, p=self.p, training=self.training))
            if self.activation:
                x = F.relu(x)
        return x


class _BatchNorm2d(nn.BatchNorm2d):
    """nn.BatchNorm2d (same parameters, buffers and state-dict keys). In training on CUDA the statistics are taken
    with plain reductions: the library's NCHW kernels run one block per channel, and with the 8 / 16 channels of the
    structure CNNs that is 8 / 16 blocks on 148 SMs (~0.5 ms each way on a [4096, 8, 4, 32] batch)."""

    def forward(self, x):
        if not (self.training and x.is_cuda and self.track_running_stats and self.momentum is not None):
            return super().forward(x)
        xf = x.float()
        var, mean = torch.var_mean(xf, dim=(0, 2, 3), unbiased=False)
        n = x.numel() // x.size(1)
        with torch.no_grad():
            self.running_mean.lerp_(mean, self.momentum)
            self.running_var.lerp_(var * (n / max(n - 1, 1)), self.momentum)
            self.num_batches_tracked.add_(1)
        scale = self.weight.float() * torch.rsqrt(var + self.eps)
        shift = self.bias.float() - mean * scale
        return (xf * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)).to(x.dtype)


class CNNEncoder(nn.Module):
    """[*, 4, 32] structure bar -> vector (model.py:211-256); Sequential indices match the reference keys."""

    def __init__(self, output_dim=256, dense_dim=256, batch_norm=False, dropout=0.1):
        super().__init__()
        conv = [nn.Conv2d(1, 8, 3, padding=1)]
        if batch_norm:
            conv.append(_BatchNorm2d(8))
        conv += [nn.ReLU(True), nn.MaxPool2d((1, 4), stride=(1, 4)), nn.Conv2d(8, 16, 3, padding=1)]
        if batch_norm:
            conv.append(_BatchNorm2d(16))
        conv.append(nn.ReLU(True))
        self.conv = nn.Sequential(*conv)
        self.flatten = nn.Flatten(start_dim=1)
        self.lin = nn.Sequential(nn.Dropout(dropout), nn.Linear(16 * 4 * 8, dense_dim), nn.ReLU(True),
                                 nn.Dropout(dropout), nn.Linear(dense_dim, output_dim))

    def forward(self, x):
        return self.lin(self.flatten(self.conv(x.unsqueeze(1))))


class CNNDecoder(nn.Module):
    """vector -> [*, 1, 1, 4, 32] structure logits (model.py:259-299)."""

    def __init__(self, input_dim=256, dense_dim=256, batch_norm=False, dropout=0.1):
        super().__init__()
        self.lin = nn.Sequential(nn.Dropout(dropout), nn.Linear(input_dim, dense_dim), nn.ReLU(True),
                                 nn.Dropout(dropout), nn.Linear(dense_dim, 16 * 4 * 8), nn.ReLU(True))
        self.unflatten = nn.Unflatten(dim=1, unflattened_size=(16, 4, 8))
        conv = [nn.Upsample(scale_factor=(1, 4), mode="nearest"), nn.Conv2d(16, 8, 3, padding=1)]
        if batch_norm:
            conv.append(_BatchNorm2d(8))
        conv += [nn.ReLU(True), nn.Conv2d(8, 1, 3, padding=1)]
        self.conv = nn.Sequential(*conv)

    def forward(self, x):
        return self.conv(self.unflatten(self.lin(x))).unsqueeze(1)


class GlobalAttention(nn.Module):
    """Soft-attention pooling over the nodes of each bar (PyG GlobalAttention, model.py:335-340,409)."""

    def __init__(self, gate_nn: nn.Module):
        super().__init__()
        self.gate_nn = gate_nn
        _reset(self.gate_nn)

    def _gate(self, x):
        """gate_nn(x); its trailing BatchNorm1d(1) over ~1e5 rows is done with plain reductions (the library's
        single-channel BatchNorm kernels reduce in one block and cost ~0.5 ms each way)."""
        nn_ = self.gate_nn
        if not (isinstance(nn_, nn.Sequential) and len(nn_) == 2 and isinstance(nn_[1], nn.BatchNorm1d)
                and nn_[1].num_features == 1 and nn_[1].training and nn_[1].track_running_stats):
            return nn_(x)
        bn = nn_[1]
        g = nn_[0](x).float()
        n = g.numel()
        mean = g.mean()
        var = (g - mean).square().mean()
        with torch.no_grad():
            m = bn.momentum if bn.momentum is not None else 0.1
            bn.running_mean.mul_(1 - m).add_(mean.detach().view(1), alpha=m)
            bn.running_var.mul_(1 - m).add_((var.detach() * (n / max(n - 1, 1))).view(1), alpha=m)
            bn.num_batches_tracked.add_(1)
        return (g - mean) * torch.rsqrt(var + bn.eps) * bn.weight + bn.bias

    def forward(self, x, batch, size: int, bar_ptr: Optional[torch.Tensor] = None):
        gate = self._gate(x).view(-1, 1)
        if bar_ptr is not None and x.is_cuda and bar_ptr.numel() == size + 1:
            return ops.bar_pool(x, gate, bar_ptr)      # segments are contiguous: fused, deterministic
        top = torch.full((size, 1), float("-inf"), dtype=gate.dtype, device=gate.device)
        top = top.scatter_reduce(0, batch.view(-1, 1), gate.detach(), reduce="amax", include_self=True)
        e = torch.exp(gate - top.index_select(0, batch))
        denom = torch.zeros((size, 1), dtype=gate.dtype, device=gate.device).index_add_(0, batch, e)
        alpha = e / (denom.index_select(0, batch) + 1e-16)
        return torch.zeros((size, x.size(-1)), dtype=x.dtype, device=x.device).index_add_(0, batch, alpha * x)


def _drum_split(graph):
    """(perm, n_drum): node ids with drum nodes first (original order kept), sized without a host sync when
    the graph came from the device builder."""
    cached = getattr(graph, "_drum_split", None)
    if cached is not None:
        return cached
    is_drum = graph.is_drum
    n_drum = getattr(graph, "n_drum", None)
    if n_drum is None:
        n_drum = int(is_drum.sum())
    perm = torch.argsort(is_drum.to(torch.uint8), descending=True, stable=True)
    graph._drum_split = (perm, int(n_drum))
    return graph._drum_split


class ContentEncoder(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        self.__dict__.update(kwargs)
        d = self.d
        self.dropout_layer = nn.Dropout(p=self.dropout)
        self.non_drums_pitch_emb = nn.Linear(N_PITCH_TOKENS, d // 2)
        self.drums_pitch_emb = nn.Linear(N_PITCH_TOKENS, d // 2)
        self.dur_emb = nn.Linear(N_DUR_TOKENS, d // 2)
        self.bn_non_drums = nn.BatchNorm1d(num_features=d // 2)
        self.bn_drums = nn.BatchNorm1d(num_features=d // 2)
        self.bn_dur = nn.BatchNorm1d(num_features=d // 2)
        self.chord_encoder = nn.Linear(d * (MAX_SIMU_TOKENS - 1), d)
        self.graph_encoder = GCN(dropout=self.dropout, input_dim=d, hidden_dim=d, n_layers=self.gnn_n_layers,
                                 num_relations=N_EDGE_TYPES, batch_norm=self.batch_norm)
        gate_nn = nn.Sequential(MLP(input_dim=d, output_dim=1, num_layers=1, activation=False, dropout=self.dropout),
                                nn.BatchNorm1d(1))
        self.graph_attention = GlobalAttention(gate_nn)
        self.bars_encoder = nn.Linear(self.n_bars * d, d)

    def _embed(self, tokens, pitch_emb, pitch_bn):
        """tokens [k, 15, 230] one-hot -> chord embedding [k, d] (model.py:355-388)."""
        k, t, half = tokens.size(0), tokens.size(1), self.d // 2
        pitch = pitch_bn(pitch_emb(tokens[..., :N_PITCH_TOKENS]).view(-1, half)).view(k, t, half)
        dur = self.bn_dur(self.dur_emb(tokens[..., N_PITCH_TOKENS:]).view(-1, half)).view(k, t, half)
        chord = self.chord_encoder(torch.cat((pitch, dur), dim=-1).view(k, t * self.d))
        return self.dropout_layer(F.relu(chord))

    @staticmethod
    def _bn_table(emb: nn.Linear, bn: nn.BatchNorm1d, ids: Optional[torch.Tensor], training: bool,
                  counts: Optional[torch.Tensor] = None) -> torch.Tensor:
        """BatchNorm(Linear(one_hot(ids))) as a [vocab, c] table.

        A Linear applied to a one-hot row is a column of its weight plus the bias, so the rows BatchNorm sees
        take only `vocab` distinct values: its batch statistics are the histogram-weighted statistics of those
        values. Same function as model.py:357-362 / 369-376 (fp32 round-off aside), without materialising the
        [N*15, 230] one-hot input or the [N*15, d/2] pre-norm activations. ``counts`` (the token histogram, [vocab])
        may be given instead of ``ids``; an empty histogram leaves the running statistics untouched.
        """
        table = emb.weight.float().t() + emb.bias.float()                       # [vocab, c]
        use_batch_stats = training or bn.running_mean is None
        if use_batch_stats:
            if counts is None:
                counts = torch.bincount(ids.reshape(-1), minlength=table.size(0))
            cnt = counts.to(table.dtype)
            n_true = cnt.sum()
            n = n_true.clamp(min=1.0)
            mean = (cnt @ table) / n
            var = (cnt @ (table - mean).square()) / n                           # biased, as BatchNorm normalises
            if training and bn.running_mean is not None:
                with torch.no_grad():
                    m = bn.momentum if bn.momentum is not None else 0.1
                    m_eff = m * (n_true > 0).to(table.dtype)
                    unbiased = var.detach() * (n / (n - 1).clamp(min=1.0))
                    bn.running_mean.add_(m_eff * (mean.detach() - bn.running_mean))
                    bn.running_var.add_(m_eff * (unbiased - bn.running_var))
                    bn.num_batches_tracked.add_((n_true > 0).to(bn.num_batches_tracked.dtype))
        else:
            mean, var = bn.running_mean.float(), bn.running_var.float()
        return (table - mean) * torch.rsqrt(var + bn.eps) * bn.weight.float() + bn.bias.float()

    def _bn_tables_batched(self, cnt_p: torch.Tensor, cnt_d: torch.Tensor):
        """The four training-mode `_bn_table`s of a step in one batched computation (same arithmetic, ~1/3 of the
        kernel launches): rows 0 / 1 = non-drum / drum pitch tables, rows 2 / 3 = non-drum / drum duration tables
        (duration rows padded to the pitch vocabulary with zero counts). cnt_p int [2, 131], cnt_d int [2, 99]
        (index 1 = drum nodes). Running statistics are updated in the reference's order: drum rows before the others,
        `bn_dur` twice."""
        bns = (self.bn_non_drums, self.bn_drums, self.bn_dur, self.bn_dur)
        pad_v = N_PITCH_TOKENS - N_DUR_TOKENS
        dur_w = F.pad(self.dur_emb.weight.float(), (0, pad_v))                           # [c, 131], zero columns
        w_all = torch.stack((self.non_drums_pitch_emb.weight.float(), self.drums_pitch_emb.weight.float(), dur_w, dur_w))
        b_all = torch.stack((self.non_drums_pitch_emb.bias.float(), self.drums_pitch_emb.bias.float(),
                             self.dur_emb.bias.float(), self.dur_emb.bias.float()))
        table = w_all.transpose(1, 2) + b_all.unsqueeze(1)                               # [4, 131, c]
        cnt = torch.cat((cnt_p, F.pad(cnt_d, (0, pad_v))), dim=0).to(table.dtype)        # [4, 131]
        n_true = cnt.sum(1, keepdim=True)                                                # [4, 1]
        n = n_true.clamp(min=1.0)
        mean = torch.bmm(cnt.unsqueeze(1), table).squeeze(1) / n                         # [4, c]
        ctr = table - mean.unsqueeze(1)
        var = torch.bmm(cnt.unsqueeze(1), ctr.square()).squeeze(1) / n                   # biased, as BatchNorm normalises
        with torch.no_grad():
            unbiased = var * (n / (n - 1).clamp(min=1.0))
            alive = (n_true > 0).to(table.dtype)                                         # empty set: statistics untouched
            for i in (1, 0, 3, 2):                                                       # drums first; bn_dur: drum, other
                bn = bns[i]
                m_eff = alive[i] * bn.momentum
                bn.running_mean.add_(m_eff * (mean[i] - bn.running_mean))
                bn.running_var.add_(m_eff * (unbiased[i] - bn.running_var))
                bn.num_batches_tracked.add_(alive[i, 0].to(bn.num_batches_tracked.dtype))
        gamma = torch.stack([bn.weight.float() for bn in bns])
        beta = torch.stack([bn.bias.float() for bn in bns])
        eps = bns[0].eps if len({bn.eps for bn in bns}) == 1 else torch.tensor(
            [bn.eps for bn in bns], dtype=table.dtype, device=table.device).view(4, 1)
        out = ctr * (torch.rsqrt(var + eps) * gamma).unsqueeze(1) + beta.unsqueeze(1)
        return out[:2], out[2:, :N_DUR_TOKENS]

    def _embed_folded(self, tokens: torch.Tensor, is_drum: torch.Tensor) -> torch.Tensor:
        """tokens int16 [N, 16, 2], is_drum bool [N] -> relu'd chord embeddings f32 [N, d] in node order.

        Everything between the token ids and chord_encoder's output is linear in the one-hot tokens (embedding,
        BatchNorm with the statistics of `_bn_table`, concatenation, Linear), so it folds into a table
        T[set, slot, token, :] = BN(emb)[token] @ W_chord[:, slot, half]^T — 2 x 15 x 230 rows built here from the live
        parameters (autograd reaches them through the einsum) — and the chord is a 30-row gather-sum per node
        (ops.chord_embed). Same function as model.py:355-388; drum / non-drum nodes pick their table set in place, so
        there is no split, permutation or concatenation of node rows either."""
        t, half, d = MAX_SIMU_TOKENS - 1, self.d // 2, self.d
        with torch.autocast(device_type=tokens.device.type, enabled=False):
            counts = (None, None)
            if self.training:
                # token histograms per table set in one kernel (torch.bincount reads the maximum back to the host: a
                # full device sync in the middle of the forward pass)
                both = torch.empty((2, N_PITCH_TOKENS + N_DUR_TOKENS), dtype=torch.int64, device=tokens.device)
                flags = is_drum.view(torch.uint8) if is_drum.dtype == torch.bool else is_drum
                with _ffi.on_device(tokens.device):
                    _ffi.call("pb_token_hist", tokens.data_ptr(), tokens.stride(0), 2, t, flags.data_ptr(), tokens.size(0),
                              N_PITCH_TOKENS, N_DUR_TOKENS, both.data_ptr(), _ffi.stream())
                cnt_p, cnt_d = both[:, :N_PITCH_TOKENS], both[:, N_PITCH_TOKENS:]
                counts = (cnt_p, cnt_d)
            if self.training and all(bn.track_running_stats and bn.momentum is not None
                                     for bn in (self.bn_drums, self.bn_non_drums, self.bn_dur)):
                p_tabs, d_tabs = self._bn_tables_batched(*counts)                  # [2, 131, c], [2, 99, c]
            else:
                pick = lambda c, i: None if c is None else c[i]
                # the reference embeds the drum rows first, then the others (bn_dur's running statistics see both)
                p_drum = self._bn_table(self.drums_pitch_emb, self.bn_drums, None, self.training, pick(counts[0], 1))
                d_drum = self._bn_table(self.dur_emb, self.bn_dur, None, self.training, pick(counts[1], 1))
                p_other = self._bn_table(self.non_drums_pitch_emb, self.bn_non_drums, None, self.training, pick(counts[0], 0))
                d_other = self._bn_table(self.dur_emb, self.bn_dur, None, self.training, pick(counts[1], 0))
                p_tabs, d_tabs = torch.stack((p_other, p_drum)), torch.stack((d_other, d_drum))
            w = self.chord_encoder.weight.float().view(d, t, 2, half)             # [out, slot, pitch|dur, half]
            t_pitch = torch.einsum("svh,oth->stvo", p_tabs, w[:, :, 0])
            t_dur = torch.einsum("svh,oth->stvo", d_tabs, w[:, :, 1])
            tables = torch.cat((t_pitch, t_dur), dim=2)                           # [2, 15, 230, d]
            return ops.chord_embed(tables, self.chord_encoder.bias, tokens, is_drum, tok_offset=2,
                                   dur_off=N_PITCH_TOKENS)

    def _embed_ids(self, ids, pitch_emb, pitch_bn):
        """ids int [k, 15, 2] (pitch id, duration id) -> chord embedding [k, d]; token-table form of `_embed`."""
        k, t = ids.size(0), ids.size(1)
        if k == 0:
            return torch.zeros((0, self.d), dtype=torch.float32, device=ids.device)
        p_ids, d_ids = ids[..., 0].long(), ids[..., 1].long()
        p_tab = self._bn_table(pitch_emb, pitch_bn, p_ids, self.training)
        d_tab = self._bn_table(self.dur_emb, self.bn_dur, d_ids, self.training)
        if ops.get_precision() == "bf16":          # gather straight into the GEMM's operand dtype
            p_tab, d_tab = p_tab.to(torch.bfloat16), d_tab.to(torch.bfloat16)
        tokens = torch.cat((ops.table_gather(p_tab, p_ids), ops.table_gather(d_tab, d_ids)), dim=-1).view(k, t * self.d)
        chord = ops.tc_linear(tokens, self.chord_encoder.weight, self.chord_encoder.bias)
        return self.dropout_layer(F.relu(chord))

    def forward(self, graph):
        ids = getattr(graph, "c_tokens", None)
        if ids is not None and ids.is_cuda and ids.dtype == torch.int16 and ids.is_contiguous() and self.d % 4 == 0:
            graph.x = self.dropout_layer(self._embed_folded(ids, graph.is_drum.contiguous()))
        else:
            perm, n_drum = _drum_split(graph)
            if ids is not None:                                    # token ids off the accelerated path (CPU, int64)
                ids = ids[:, 1:, :]                                # drop SOS
                drums = self._embed_ids(ids.index_select(0, perm[:n_drum]), self.drums_pitch_emb, self.bn_drums)
                others = self._embed_ids(ids.index_select(0, perm[n_drum:]), self.non_drums_pitch_emb, self.bn_non_drums)
            else:                                                  # reference layout: one-hot float c_tensor
                c = graph.c_tensor[:, 1:, :]
                drums = self._embed(c.index_select(0, perm[:n_drum]), self.drums_pitch_emb, self.bn_drums)
                others = self._embed(c.index_select(0, perm[n_drum:]), self.non_drums_pitch_emb, self.bn_non_drums)
            x = torch.empty((perm.numel(), self.d), dtype=drums.dtype, device=drums.device)
            x = x.index_copy(0, perm, torch.cat((drums, others.to(drums.dtype)), dim=0))
            graph.x = x.float()
        graph.distinct_bars = graph.bars + self.n_bars * graph.batch
        h = self.graph_encoder(graph)
        n_seg = _n_segments(graph, self.n_bars)
        with torch.autocast(device_type=h.device.type, enabled=False):
            pooled = self.graph_attention(h.float(), batch=graph.distinct_bars, size=n_seg,
                                          bar_ptr=getattr(graph, "bar_ptr", None))
        return self.bars_encoder(pooled.view(-1, self.n_bars * self.d))


def _n_segments(graph, n_bars: int) -> int:
    n_graphs = getattr(graph, "num_graphs", None)
    if n_graphs is None:
        n_graphs = int(graph.batch[-1]) + 1
    return int(n_graphs) * n_bars


class StructureEncoder(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        self.__dict__.update(kwargs)
        self.cnn_encoder = CNNEncoder(dense_dim=self.d, output_dim=self.d, dropout=self.dropout,
                                      batch_norm=self.batch_norm)
        self.bars_encoder = nn.Linear(self.n_bars * self.d, self.d)

    def forward(self, graph):
        bars = graph.s_tensor.view(-1, N_TRACKS, self.resolution * 4)
        return self.bars_encoder(self.cnn_encoder(bars).view(-1, self.n_bars * self.d))


class Encoder(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        self.__dict__.update(kwargs)
        self.s_encoder = StructureEncoder(**kwargs)
        self.c_encoder = ContentEncoder(**kwargs)
        self.dropout_layer = nn.Dropout(p=self.dropout)
        self.linear_merge = nn.Linear(2 * self.d, self.d)
        self.bn_linear_merge = nn.BatchNorm1d(num_features=self.d)
        self.linear_mu = nn.Linear(self.d, self.d)
        self.linear_log_var = nn.Linear(self.d, self.d)

    def forward(self, graph):
        z_s = self.s_encoder(graph)
        z_c = self.c_encoder(graph)
        z = self.dropout_layer(torch.cat((z_c, z_s), dim=1))
        z = self.dropout_layer(F.relu(self.bn_linear_merge(self.linear_merge(z))))
        return self.linear_mu(z), self.linear_log_var(z)


class StructureDecoder(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        self.__dict__.update(kwargs)
        self.bars_decoder = nn.Linear(self.d, self.d * self.n_bars)
        self.cnn_decoder = CNNDecoder(input_dim=self.d, dense_dim=self.d, dropout=self.dropout,
                                      batch_norm=self.batch_norm)

    def forward(self, z_s):
        out = self.cnn_decoder(self.bars_decoder(z_s).reshape(-1, self.d))
        return out.view(z_s.size(0), self.n_bars, N_TRACKS, -1)


class ContentDecoder(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        self.__dict__.update(kwargs)
        d = self.d
        self.bars_decoder = nn.Linear(d, d * self.n_bars)
        self.graph_decoder = GCN(dropout=self.dropout, input_dim=d, hidden_dim=d, n_layers=self.gnn_n_layers,
                                 num_relations=N_EDGE_TYPES, batch_norm=self.batch_norm)
        self.chord_decoder = nn.Linear(d, d * (MAX_SIMU_TOKENS - 1))
        self.drums_pitch_emb = nn.Linear(d // 2, N_PITCH_TOKENS)
        self.non_drums_pitch_emb = nn.Linear(d // 2, N_PITCH_TOKENS)
        self.dur_emb = nn.Linear(d // 2, N_DUR_TOKENS)
        self.dropout_layer = nn.Dropout(p=self.dropout)
        self.materialize_logits = True      # False: return LogitParts (loss-only consumers, see train.TrainStep)

    def forward(self, z_c, s):
        d, half = self.d, self.d // 2
        z_bar = self.bars_decoder(z_c).view(-1, d)                  # one row per (sequence, bar)
        s.distinct_bars = s.bars + self.n_bars * s.batch
        bar_ptr = getattr(s, "bar_ptr", None)
        if bar_ptr is not None and z_bar.is_cuda and bar_ptr.numel() == z_bar.size(0) + 1:
            s.x = ops.bar_expand(z_bar, bar_ptr, s.num_nodes)       # every node starts from its bar's code
        else:
            s.x = z_bar.index_select(0, s.distinct_bars).float()
        bf16 = ops.get_precision() == "bf16"
        dropout_active = self.training and self.dropout_layer.p > 0
        fold = z_bar.is_cuda and not self.materialize_logits and not dropout_active and d % 64 == 0
        if fold and getattr(s, "n_drum", None) is not None and ops.split_heads_enabled():
            # loss-only consumers: rows leave the GCN drum nodes first, so that each block only gets the pitch head of
            # its own instrument class (5/8 of the un-embedding GEMM and of the logits)
            perm, n_drum = _drum_split(s)
            h = self.graph_decoder(s, out_perm=perm, out_dtype=torch.bfloat16 if bf16 else torch.float32)
            t = MAX_SIMU_TOKENS - 1
            return self._folded_heads_split(h, self.chord_decoder.weight.view(t, 2, half, d),
                                            self.chord_decoder.bias.view(t, 2, half), s.is_drum, perm, n_drum, bf16)
        h = self.graph_decoder(s)
        # chord_decoder emits, per token slot, a pitch half and a duration half (model.py:549-567). Applying the
        # two row-subsets of its weight separately gives the same numbers without slicing a [N, 15, d] activation
        # (whose backward would zero-fill and copy two full-size tensors).
        t = MAX_SIMU_TOKENS - 1
        w = self.chord_decoder.weight.view(t, 2, half, d)
        b = self.chord_decoder.bias.view(t, 2, half)
        if fold:
            return self._folded_heads(h, w, b, s.is_drum, bf16)
        h_pitch = ops.tc_linear(h, w[:, 0].reshape(t * half, d), b[:, 0].reshape(-1), out_bf16=bf16)
        h_dur = ops.tc_linear(h, w[:, 1].reshape(t * half, d), b[:, 1].reshape(-1), out_bf16=bf16)
        h_pitch = self.dropout_layer(h_pitch).view(-1, t, half)
        h_dur = self.dropout_layer(h_dur).view(-1, t, half)
        # both pitch heads on every node, then select per node: no compaction, no host sync
        is_drum = s.is_drum.view(-1, 1, 1)
        if h.is_cuda:
            # un-embedding heads (131 / 99 outputs) on the tcgen05 GEMM: output width padded to a multiple of 64 with
            # zero weight rows and a -inf bias, so the padded logits vanish from every softmax downstream
            lazy_bf16 = bf16 and not self.materialize_logits    # loss-only: bf16 logits, as under autocast
            drums = _padded_head(self.drums_pitch_emb, h_pitch, lazy_bf16)
            others = _padded_head(self.non_drums_pitch_emb, h_pitch, lazy_bf16)
            dur_pad = _padded_head(self.dur_emb, h_dur, lazy_bf16)
            if not self.materialize_logits:
                # training loops that only need the loss: skip assembling [N, 15, 230] (a select + a concatenation of
                # GB-sized tensors and their backward); vae_losses selects per node at the level of the NLL instead
                return LogitParts(drums, others, dur_pad, s.is_drum)
            pitch_pad = torch.where(is_drum, drums, others)
            pitch, dur = pitch_pad[..., :N_PITCH_TOKENS], dur_pad[..., :N_DUR_TOKENS]
            c_logits = torch.cat((pitch, dur), dim=-1)
            c_logits._parts = (pitch_pad, dur_pad)   # lets the loss skip re-slicing (padding columns are -inf)
            return c_logits
        pitch = torch.where(is_drum, self.drums_pitch_emb(h_pitch), self.non_drums_pitch_emb(h_pitch))
        dur = self.dur_emb(h_dur)
        c_logits = torch.cat((pitch, dur), dim=-1)
        c_logits._parts = (pitch, dur)          # lets the loss skip re-slicing the concatenation
        return c_logits

    def _folded_heads(self, h, w, b, is_drum, bf16: bool) -> "LogitParts":
        """With the dropout between chord_decoder and the un-embedding heads inactive (p = 0 as in training.json, or
        eval) the two Linears compose: logits[n, t] = h[n] @ (W_head W_cd[t, half])^T + (W_head b_cd[t, half] + b_head).
        One GEMM h [N, d] x [15 * 512, d]^T then yields, per token slot, [drum-pitch 192 | non-drum-pitch 192 |
        duration 128] logits (131 / 131 / 99 real columns, the rest -inf) — same function as model.py:549-576 without
        the [N, 15, d] intermediate or the three head GEMMs over N * 15 rows. Autograd reaches the four Linears through
        the small composition einsums."""
        t = w.size(0)
        with torch.autocast(device_type=h.device.type, enabled=False):
            composed = self._composed_heads(w, b)
            blocks_w, blocks_b = [c[0] for c in composed], [c[1] for c in composed]
            widths = [x.size(1) for x in blocks_w]
            w_all = torch.cat(blocks_w, dim=1)                                          # [t, 512, d]
            b_all = torch.cat(blocks_b, dim=1)
            cols = w_all.size(1)
            out = ops.tc_linear(h.float(), w_all.view(t * cols, -1), b_all.view(-1), out_bf16=bf16)
        out = out.view(-1, t, cols)
        c0, c1 = widths[0], widths[0] + widths[1]
        parts = LogitParts(out[..., :c0], out[..., c0:c1], out[..., c1:], is_drum)
        parts.combined, parts.widths = out, widths
        return parts


    def _composed_heads(self, w, b):
        """Per head (drum pitch, non-drum pitch, duration): W_head W_cd[t, half] [t, C64, d] and its bias [t, C64],
        C padded to a multiple of 64 with zero rows / -inf bias. The compositions are [C, half] x [half, t d] products of
        the parameters: on the tcgen05 GEMM in its fp32-grade TF32x3 mode (the library's SIMT sgemm took 0.1-0.17 ms
        for each of them and for each of their gradients)."""
        out = []
        t, _, half, d = w.shape
        for lin, which in ((self.drums_pitch_emb, 0), (self.non_drums_pitch_emb, 0), (self.dur_emb, 1)):
            hw, hb = lin.weight.float(), lin.bias.float()
            pad = (-hw.size(0)) % 64
            hw_p = F.pad(hw, (0, 0, 0, pad))                                               # [C64, half], zero rows
            w_t = w[:, which].float().permute(0, 2, 1).reshape(t * d, half)                 # [(t, d), half]
            wc = ops.tc_linear(hw_p, w_t, None, precision="fp32").view(-1, t, d).permute(1, 0, 2)   # [t, C64, d]
            bc = torch.einsum("ch,th->tc", hw, b[:, which].float()) + hb                   # [t, C]
            out.append((wc, F.pad(bc, (0, pad), value=float("-inf"))))
        return out

    def _folded_heads_split(self, h, w, b, is_drum, perm, n_drum: int, bf16: bool) -> "LogitParts":
        """_folded_heads on rows sorted drum-first (``h`` = node features in ``perm`` order): the drum block goes through
        [drum-pitch 192 | duration 128], the other block through [non-drum-pitch 192 | duration 128] — the pitch head a
        node does not use (model.py:571-574 selects by is_drum) is never computed."""
        t = w.size(0)
        with torch.autocast(device_type=h.device.type, enabled=False):
            (wd, bd), (wo, bo), (wu, bu) = self._composed_heads(w, b)
            cols = wd.size(1) + wu.size(1)
            w_d, b_d = torch.cat((wd, wu), dim=1).view(t * cols, -1), torch.cat((bd, bu), dim=1).view(-1)
            w_o, b_o = torch.cat((wo, wu), dim=1).view(t * cols, -1), torch.cat((bo, bu), dim=1).view(-1)
            out_d, out_o = ops.split_rows_linear(h, n_drum, w_d, b_d, w_o, b_o, out_bf16=bf16)
        parts = LogitParts(None, None, None, is_drum)
        parts.split = (out_d.view(-1, t, cols), out_o.view(-1, t, cols), perm, (wd.size(1), wu.size(1)))
        return parts


class LogitParts:
    """Content logits kept as the three head outputs (drum-pitch, non-drum-pitch, duration; width padded with -inf)
    plus the per-node drum flag — what ``c_logits`` is assembled from. ``dense()`` builds the reference's
    ``[N, 15, 230]`` tensor (model.py:571-576)."""

    def __init__(self, drums, others, dur, is_drum):
        self.drums, self.others, self.dur, self.is_drum = drums, others, dur, is_drum
        self.combined, self.widths = None, None     # set when the three heads are column blocks of one matrix
        # (drum rows [n_drum, t, wp + wu], other rows [.., t, wp + wu], perm, (wp, wu)): rows sorted drum-first, each
        # block holding its own pitch head and the duration head
        self.split = None

    def dense(self) -> torch.Tensor:
        if self.split is not None:
            out_d, out_o, perm, (wp, _) = self.split
            rows = torch.cat((out_d, out_o), dim=0).float()
            rows = torch.cat((rows[..., :N_PITCH_TOKENS], rows[..., wp:wp + N_DUR_TOKENS]), dim=-1)
            return torch.empty_like(rows).index_copy_(0, perm, rows)
        pitch = torch.where(self.is_drum.view(-1, 1, 1), self.drums, self.others)
        return torch.cat((pitch[..., :N_PITCH_TOKENS], self.dur[..., :N_DUR_TOKENS]), dim=-1).float()


def _padded_head(lin: nn.Linear, h: torch.Tensor, out_bf16: bool = False) -> torch.Tensor:
    """lin(h) for h [n, t, k] with the output width padded up to a multiple of 64: [n, t, n_pad], padding = -inf."""
    n_out, k = lin.weight.shape
    pad = (-n_out) % 64
    w = F.pad(lin.weight, (0, 0, 0, pad))
    b = F.pad(lin.bias, (0, pad), value=float("-inf"))
    out = ops.tc_linear(h.reshape(-1, k), w, b, out_bf16=out_bf16)
    return out.view(*h.shape[:-1], n_out + pad)


class Decoder(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        self.__dict__.update(kwargs)
        self.lin_decoder = nn.Linear(self.d, 2 * self.d)
        self.batch_norm = nn.BatchNorm1d(num_features=2 * self.d)
        self.dropout = nn.Dropout(p=self.dropout)
        self.s_decoder = StructureDecoder(**kwargs)
        self.c_decoder = ContentDecoder(**kwargs)
        self.sigmoid_thresh = 0.5

    def _structure_from_binary(self, s_tensor) -> Graph:
        """bool [B, n_bars, 4, 32] -> batched graph on the model's device (model.py:596-607)."""
        return graphs_from_tensor(s_tensor, device=next(self.parameters()).device)

    def _binary_from_logits(self, s_logits):
        s = torch.sigmoid(s_logits) >= self.sigmoid_thresh
        empty = ~s.flatten(-2).any(dim=-1)
        s[..., 0, 0] |= empty                                      # fake activation (model.py:617-621)
        return s

    def _structure_from_logits(self, s_logits) -> Graph:
        return self._structure_from_binary(self._binary_from_logits(s_logits))

    def forward(self, z, s=None):
        z = self.dropout(F.relu(self.batch_norm(self.lin_decoder(z))))
        z_s, z_c = z[:, : self.d], z[:, self.d:]
        s_logits = self.s_decoder(z_s)
        if s is None:
            s = self._structure_from_logits(s_logits.detach())
        return s_logits, self.c_decoder(z_c, s)


class VAE(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        self.encoder = Encoder(**kwargs)
        self.decoder = Decoder(**kwargs)

    def forward(self, graph, noise: Optional[torch.Tensor] = None):
        mu, log_var = self.encoder(graph)
        eps = torch.randn_like(mu) if noise is None else noise
        z = torch.exp(0.5 * log_var) * eps + mu
        return self.decoder(z, graph), mu, log_var
