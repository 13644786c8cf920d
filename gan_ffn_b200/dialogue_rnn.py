"""The DialogueRNN classifier head that ``GAN_FFN_DialogueRNN`` puts on top of the fused features
(reference model.py:828-1062 ``DialogueRNNCell`` / ``DialogueRNN`` / ``BiModel``, attention modules model.py:22-37 and
:136-201).

Scope: the *fusion* part of ``GAN_FFN_DialogueRNN`` runs on the sm_100a kernels; this
head is a per-time-step GRU recurrence that stays on stock PyTorch, fed by the fused features.  It is written from
the reference's behaviour, not from its code: same submodule / parameter names and shapes (so ``state_dict`` keys are
interchangeable), same arithmetic, but

  * party selection is a ``gather`` on the speaker index instead of a Python loop over the batch,
  * the global-state history is a preallocated ``(S, B, D_g)`` buffer instead of a tensor re-concatenated per step,
  * sequence reversal is one index ``gather`` instead of a per-dialogue flip + ``pad_sequence``,
  * the second-level matching attention over all time steps is one batched contraction instead of a loop over ``t``.

``tests/test_dialogue_rnn.py`` holds it to the unmodified reference (fixtures written by ``oracle/make_golden.py``).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


_SIDE_STREAMS = {}   # device index -> side stream of BiModel.forward (module level: modules must stay picklable)


class SimpleAttention(nn.Module):
    """reference model.py:22-37: softmax over time of a learned scalar score, weighted sum of the memory."""

    def __init__(self, input_dim):
        super().__init__()
        self.input_dim = input_dim
        self.scalar = nn.Linear(input_dim, 1, bias=False)

    def forward(self, M, x=None):
        alpha = F.softmax(self.scalar(M), dim=0).permute(1, 2, 0)          # (B, 1, T)
        return torch.bmm(alpha, M.transpose(0, 1))[:, 0, :], alpha


class MatchingAttention(nn.Module):
    """reference model.py:136-201.  ``M`` (T, B, mem_dim) memory, ``x`` (B, cand_dim) query, ``mask`` (B, T)."""

    def __init__(self, mem_dim, cand_dim, alpha_dim=None, att_type="general2"):
        super().__init__()
        assert att_type != "concat" or alpha_dim is not None
        assert att_type != "dot" or mem_dim == cand_dim
        self.mem_dim, self.cand_dim, self.att_type = mem_dim, cand_dim, att_type
        if att_type == "general":
            self.transform = nn.Linear(cand_dim, mem_dim, bias=False)
        if att_type == "general2":
            self.transform = nn.Linear(cand_dim, mem_dim, bias=True)
            torch.nn.init.normal_(self.transform.weight, std=0.01)
        elif att_type == "concat":
            self.transform = nn.Linear(cand_dim + mem_dim, alpha_dim, bias=False)
            self.vector_prod = nn.Linear(alpha_dim, 1, bias=False)

    def forward(self, M, x, mask=None):
        if mask is None:
            mask = torch.ones(M.size(1), M.size(0), dtype=M.dtype, device=M.device)
        Mb = M.transpose(0, 1)                                               # (B, T, mem)
        if self.att_type == "dot":
            alpha = F.softmax(torch.bmm(x.unsqueeze(1), Mb.transpose(1, 2)), dim=2)
        elif self.att_type == "general":
            alpha = F.softmax(torch.bmm(self.transform(x).unsqueeze(1), Mb.transpose(1, 2)), dim=2)
        elif self.att_type == "general2":
            alpha = self._general2(self.transform(x).unsqueeze(1), Mb, mask)
        else:
            x_ = x.unsqueeze(1).expand(-1, M.size(0), -1)
            mx_a = torch.tanh(self.transform(torch.cat([Mb, x_], 2)))
            alpha = F.softmax(self.vector_prod(mx_a), 1).transpose(1, 2)
        return torch.bmm(alpha, Mb)[:, 0, :], alpha

    @staticmethod
    def _general2(xq, Mb, mask):
        """xq (B, Q, mem) transformed queries -> alpha (B, Q, T): tanh scores on the masked memory, softmax over
        time, masked and renormalised (model.py:176-190)."""
        m = mask.unsqueeze(1)                                                # (B, 1, T)
        scores = torch.bmm(xq, (Mb * mask.unsqueeze(2)).transpose(1, 2)) * m
        a = F.softmax(torch.tanh(scores), dim=2) * m
        return a / a.sum(dim=2, keepdim=True)

    def all_steps(self, M, mask):
        """``forward(M, M[t], mask)`` for every t at once (BiModel's second-level attention, model.py:1046-1051).
        Returns (T, B, mem_dim) pooled memories and the (B, T, T) weights."""
        assert self.att_type == "general2"
        Mb = M.transpose(0, 1)
        alpha = self._general2(self.transform(Mb), Mb, mask)                 # (B, T, T)
        return torch.bmm(alpha, Mb).transpose(0, 1), alpha


class DialogueRNNCell(nn.Module):
    """reference model.py:828-931: global / party / emotion GRU cells of one time step."""

    def __init__(self, D_m, D_g, D_p, D_e, listener_state=False, context_attention="simple", D_a=100, dropout=0.5):
        super().__init__()
        self.D_m, self.D_g, self.D_p, self.D_e = D_m, D_g, D_p, D_e
        self.listener_state = listener_state
        self.g_cell = nn.GRUCell(D_m + D_p, D_g)
        self.p_cell = nn.GRUCell(D_m + D_g, D_p)
        self.e_cell = nn.GRUCell(D_p, D_e)
        if listener_state:
            self.l_cell = nn.GRUCell(D_m + D_p, D_p)
        self.dropout = nn.Dropout(dropout)
        if context_attention == "simple":
            self.attention = SimpleAttention(D_g)
        else:
            self.attention = MatchingAttention(D_g, D_m, D_a, context_attention)

    @staticmethod
    def _select(X, idx):
        """X (B, party, D), idx (B,) -> X[b, idx[b]]."""
        return X.gather(1, idx.view(-1, 1, 1).expand(-1, 1, X.size(2)))[:, 0, :]

    def forward(self, U, qmask, g_hist, q0, e0):
        """U (B, D_m); qmask (B, party); g_hist (t, B, D_g) (t may be 0); q0 (B, party, D_p); e0 (B, D_e) or empty."""
        B, party = qmask.size(0), qmask.size(1)
        idx = torch.argmax(qmask, 1)
        q0_sel = self._select(q0, idx)
        g_prev = g_hist[-1] if g_hist.size(0) else U.new_zeros(B, self.D_g)
        g_ = self.dropout(self.g_cell(torch.cat([U, q0_sel], dim=1), g_prev))
        if g_hist.size(0) == 0:
            c_, alpha = U.new_zeros(B, self.D_g), None
        else:
            c_, alpha = self.attention(g_hist, U)
        U_c = torch.cat([U, c_], dim=1).unsqueeze(1).expand(-1, party, -1)
        qs_ = self.p_cell(U_c.reshape(-1, self.D_m + self.D_g), q0.reshape(-1, self.D_p)).view(B, party, self.D_p)
        qs_ = self.dropout(qs_)
        if self.listener_state:
            U_ = U.unsqueeze(1).expand(-1, party, -1).reshape(-1, self.D_m)
            ss_ = self._select(qs_, idx).unsqueeze(1).expand(-1, party, -1).reshape(-1, self.D_p)
            ql_ = self.l_cell(torch.cat([U_, ss_], 1), q0.reshape(-1, self.D_p)).view(B, party, self.D_p)
            ql_ = self.dropout(ql_)
        else:
            ql_ = q0
        qm = qmask.unsqueeze(2)
        q_ = ql_ * (1 - qm) + qs_ * qm
        e_prev = e0 if e0.numel() else U.new_zeros(B, self.D_e)
        e_ = self.dropout(self.e_cell(self._select(q_, idx), e_prev))
        return g_, q_, e_, alpha


class DialogueRNN(nn.Module):
    """reference model.py:933-978: the cell unrolled over the dialogue."""

    def __init__(self, D_m, D_g, D_p, D_e, listener_state=False, context_attention="simple", D_a=100, dropout=0.5):
        super().__init__()
        self.D_m, self.D_g, self.D_p, self.D_e = D_m, D_g, D_p, D_e
        self.dropout = nn.Dropout(dropout)
        self.dialogue_cell = DialogueRNNCell(D_m, D_g, D_p, D_e, listener_state, context_attention, D_a, dropout)

    def forward(self, U, qmask):
        """U (S, B, D_m), qmask (S, B, party) -> emotions (S, B, D_e), list of attention weights per step."""
        S, B = U.size(0), U.size(1)
        g_list, e_list, alpha = [], [], []
        g_hist = U.new_zeros(0, B, self.D_g)
        q_ = U.new_zeros(B, qmask.size(2), self.D_p)
        e_ = U.new_zeros(0)
        for t in range(S):
            g_, q_, e_, alpha_ = self.dialogue_cell(U[t], qmask[t], g_hist, q_, e_)
            g_list.append(g_)
            g_hist = torch.stack(g_list, 0)
            e_list.append(e_)
            if alpha_ is not None:
                alpha.append(alpha_[:, 0, :])
        e = torch.stack(e_list, 0) if e_list else U.new_zeros(0)
        return e, alpha


class BiModel(nn.Module):
    """reference model.py:981-1062: forward + backward DialogueRNN over each dialogue's real length, matching attention
    over time, ReLU linear, log-softmax classifier."""

    def __init__(self, D_m, D_g, D_p, D_e, D_h, n_classes=7, listener_state=False, context_attention="simple", D_a=100,
                 dropout_rec=0.5, dropout=0.5):
        super().__init__()
        self.D_m, self.D_g, self.D_p, self.D_e, self.D_h = D_m, D_g, D_p, D_e, D_h
        self.n_classes = n_classes
        self.dropout = nn.Dropout(dropout)
        self.dropout_rec = nn.Dropout(dropout + 0.15)
        self.dialog_rnn_f = DialogueRNN(D_m, D_g, D_p, D_e, listener_state, context_attention, D_a, dropout_rec)
        self.dialog_rnn_r = DialogueRNN(D_m, D_g, D_p, D_e, listener_state, context_attention, D_a, dropout_rec)
        self.linear = nn.Linear(2 * D_e, 2 * D_h)
        self.smax_fc = nn.Linear(2 * D_h, n_classes)
        self.matchatt = MatchingAttention(2 * D_e, 2 * D_e, att_type="general2")

    @staticmethod
    def _reverse_seq(X, mask, max_len=None):
        """X (S, B, D), mask (B, S): each dialogue's first ``len`` steps reversed, zero beyond; output length
        max(len) (what flipping each prefix and ``pad_sequence`` gives, model.py:1019-1031).  ``max_len``: the longest
        dialogue when the caller knows it on the host (the loader's ``lengths``) -- avoids reading it back from the
        device, which also makes the head recordable into a CUDA graph."""
        lens = mask.sum(1).int()                                             # (B,)
        L = int(max_len) if max_len is not None else (int(lens.max().item()) if lens.numel() else 0)
        t = torch.arange(L, device=X.device).unsqueeze(1)                    # (L, 1)
        src = lens.unsqueeze(0).long() - 1 - t                               # (L, B)
        valid = src >= 0
        idx = src.clamp(min=0).unsqueeze(2).expand(-1, -1, X.size(2))
        return X.gather(0, idx) * valid.unsqueeze(2).to(X.dtype)

    def forward(self, U, qmask, umask, att2=True, max_len=None):
        # The two directions are independent until the concatenation: on a CUDA device the reverse direction runs on
        # a side stream (fork / join by events, so it is recordable into a CUDA graph; autograd replays each backward
        # op on its forward stream, so the two backward recurrences overlap as well).  Every time step is a chain of
        # small dependent kernels, so one direction alone leaves the device mostly idle.
        side = None
        if U.is_cuda and max_len is not None:
            side = _SIDE_STREAMS.get(U.device.index)
            if side is None:
                side = torch.cuda.Stream(device=U.device)
                _SIDE_STREAMS[U.device.index] = side
            main = torch.cuda.current_stream(U.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                rev_U = self._reverse_seq(U, umask, max_len)
                rev_qmask = self._reverse_seq(qmask, umask, max_len)
                emotions_b, alpha_b = self.dialog_rnn_r(rev_U, rev_qmask)
                emotions_b = self._reverse_seq(emotions_b, umask, max_len)
                emotions_b = self.dropout_rec(emotions_b)
        emotions_f, alpha_f = self.dialog_rnn_f(U, qmask)
        emotions_f = self.dropout_rec(emotions_f)
        if side is not None:
            main.wait_stream(side)
            for tns in (U, qmask, umask):
                tns.record_stream(side)          # allocated on the caller's stream, read on the side stream
            for tns in [emotions_b] + [a for a in alpha_b if torch.is_tensor(a)]:
                tns.record_stream(main)          # allocated on the side stream, read on the caller's
        else:
            rev_U = self._reverse_seq(U, umask, max_len)
            rev_qmask = self._reverse_seq(qmask, umask, max_len)
            emotions_b, alpha_b = self.dialog_rnn_r(rev_U, rev_qmask)
            emotions_b = self._reverse_seq(emotions_b, umask, max_len)
            emotions_b = self.dropout_rec(emotions_b)
        emotions = torch.cat([emotions_f, emotions_b], dim=-1)
        if att2:
            att_emotions, a = self.matchatt.all_steps(emotions, umask)
            alpha = [a[:, t, :] for t in range(a.size(1))]
            hidden = F.relu(self.linear(att_emotions))
        else:
            alpha = []
            hidden = F.relu(self.linear(emotions))
        hidden = self.dropout(hidden)
        log_prob = F.log_softmax(self.smax_fc(hidden), 2)
        return log_prob, alpha, alpha_f, alpha_b
