"""CPU baseline port: the reference's PyTorch-CPU execution of the hot path, restated.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/okge_oracle.py). The reference is pure Python on top
of ATen and cannot travel to the GPU box (``/root/reference`` does not exist there), so ``bench.py``'s
``cpu_baseline`` leg and ``--impl reference`` arm time THIS module on the box's host cores: it issues
the same ATen call sequence the reference issues (nn.Embedding lookups, the 4-``mm`` ComplEx form,
``torch.cat`` of the po/sp blocks, dense ``[B, N]`` fp32 labels, BCEWithLogitsLoss(sum) / log_softmax +
KLDivLoss(sum), autograd backward, dense torch.optim.Adagrad with the inherited eps = 1e-8, and the
per-prefix Python loop of ``compute_metrics``), with ``torch.set_num_threads(os.cpu_count())`` like
``openkge/trainer.py:136``. Parity status: PINNED against tests/golden/*.npz
(tests/test_oracle_golden.py::test_torch_cpu_port_*).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


class PortModel(torch.nn.Module):
    """LookupDistmult / LookupComplex / UnigramPoolingComplex / LSTMComplex / LSTMDistmult forward exactly as the
    reference composes it (openkge/model.py:52-77, 181-278, 455-480, 762-786, 963-987), dropout = 0."""

    def __init__(self, kind: str, scorer: str, params: Dict[str, np.ndarray], pool: str = "sum",
                 batchnorm: bool = False, min_size: int = 2):
        super().__init__()
        self.kind, self.scorer, self.pool, self.min_size = kind, scorer, pool, min_size
        self.entity_embedding = torch.nn.Embedding.from_pretrained(torch.tensor(params["entity_embedding.weight"]),
                                                                   freeze=False, padding_idx=0)
        self.relation_embedding = torch.nn.Embedding.from_pretrained(torch.tensor(params["relation_embedding.weight"]),
                                                                     freeze=False, padding_idx=0)
        self.entity_batchnorm = self.relation_batchnorm = None
        if kind in ("unigram", "lstm"):
            self.register_buffer("entity_token_ids", torch.tensor(params["entity_token_ids"]).long())
            self.register_buffer("relation_token_ids", torch.tensor(params["relation_token_ids"]).long())
            if kind == "lstm":                                               # openkge/model.py:947-948
                d = self.entity_embedding.weight.size(1)
                for which in ("entity", "relation"):
                    enc = torch.nn.LSTM(input_size=d, hidden_size=d, batch_first=True)
                    enc.load_state_dict({k: torch.tensor(params[f"{which}_encoder_in.{k}"]) for k in
                                         ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")})
                    setattr(self, f"{which}_encoder_in", enc)
            if batchnorm:
                d = self.entity_embedding.weight.size(1)
                self.entity_batchnorm = torch.nn.BatchNorm1d(d, momentum=0.1, eps=1e-5)
                self.relation_batchnorm = torch.nn.BatchNorm1d(d, momentum=0.1, eps=1e-5)
                for which in ("entity", "relation"):
                    bn = getattr(self, f"{which}_batchnorm")
                    bn.load_state_dict({k: torch.tensor(params[f"{which}_batchnorm.{k}"]) for k in
                                        ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")})

    # -- embedders ---------------------------------------------------------------------------
    def _encode(self, which: str, ids: torch.Tensor) -> torch.Tensor:
        emb = getattr(self, f"{which}_embedding")
        if self.kind == "lookup":
            return emb(ids.reshape(-1).long())                                   # model.py:457-458
        tok = F.embedding(ids.reshape(-1).long(), getattr(self, f"{which}_token_ids"))   # :762-763
        embedded = emb(tok.long())                                               # [n, L, D], :767
        if self.kind == "lstm":                                                  # :963-987
            output, _ = getattr(self, f"{which}_encoder_in")(embedded)
            enc = output[range(0, tok.size(0)), (tok > 0).long().sum(1) - 1, :]
            bn = getattr(self, f"{which}_batchnorm")
            return bn(enc) if bn is not None else enc
        if self.pool == "max":
            enc, _ = embedded.max(dim=1)
        elif self.pool == "mean":
            enc = embedded.sum(dim=1) / ((tok > 0).float().sum(1, keepdim=True) + 1e-12)
        else:
            enc = embedded.sum(dim=1)
        bn = getattr(self, f"{which}_batchnorm")
        if bn is not None:
            enc = bn(enc.contiguous())
        return enc

    def all_entities(self) -> torch.Tensor:
        n = self.entity_embedding.weight.size(0) if self.kind == "lookup" else self.entity_token_ids.size(0)
        if self.kind == "lookup" and not self.training:
            return self.entity_embedding.weight[self.min_size:].contiguous()     # _get_all, :512-514
        if self.kind in ("unigram", "lstm") and not self.training:
            # precompute_embeddings_from_tokens (:670-712): every row in 4,096-row chunks under no_grad, sliced [2:]
            with torch.no_grad():
                chunks = [self._encode("entity", torch.arange(i, min(i + 4096, n))) for i in range(0, n, 4096)]
            return torch.cat(chunks)[self.min_size:]
        return self._encode("entity", torch.arange(self.min_size, n))            # precompute_batch_shared_inputs

    # -- scorers -----------------------------------------------------------------------------
    def _score(self, subj, rel, obj, sp: bool) -> torch.Tensor:
        if self.scorer == "distmult":                                            # :268-272
            return (subj * rel).mm(obj.t()) if sp else (rel * obj).mm(subj.t())
        r1, r2 = (t.contiguous() for t in rel.chunk(2, dim=1))                   # :203-205
        s1, s2 = (t.contiguous() for t in subj.chunk(2, dim=1))
        o1, o2 = (t.contiguous() for t in obj.chunk(2, dim=1))
        if sp:                                                                   # :206-209
            return (s1 * r1).mm(o1.t()) + (s2 * r1).mm(o2.t()) + (s1 * r2).mm(o2.t()) - (s2 * r2).mm(o1.t())
        return (o1 * r1).mm(s1.t()) + (o2 * r1).mm(s2.t()) + (o2 * r2).mm(s1.t()) - (o1 * r2).mm(s2.t())

    def forward(self, po: Optional[Tuple[torch.Tensor, torch.Tensor]], sp: Optional[Tuple[torch.Tensor, torch.Tensor]],
                candidate_ids: Optional[torch.Tensor] = None):
        """AddLossModule scoring (trainer.py:69-91): candidates once (all entities, or the batch-shared ids through
        precompute_batch_shared_inputs, :80-82), po block then sp block, concatenated."""
        E = self.all_entities() if candidate_ids is None else self._encode("entity", candidate_ids)
        outs = []
        if po is not None:
            rel, obj = self._encode("relation", po[0]), self._encode("entity", po[1])
            outs.append(self._score(E, rel, obj, sp=False))
        if sp is not None:
            subj, rel = self._encode("entity", sp[0]), self._encode("relation", sp[1])
            outs.append(self._score(subj, rel, E, sp=True))
        return torch.cat(outs)


def loss_sum(scores: torch.Tensor, labels: torch.Tensor, kind: str = "bce", smoothing: float = 0.0) -> torch.Tensor:
    """trainer.py:93-106."""
    if kind == "kl":
        return F.kl_div(F.log_softmax(scores, dim=1).view(-1), labels.view(-1), reduction="sum")
    if smoothing > 0:
        labels = (labels + 1.0 / labels.size(-1)) * (1 - smoothing)
    return F.binary_cross_entropy_with_logits(scores.view(-1), labels.view(-1), reduction="sum")


def dense_labels(pos_ptr: np.ndarray, pos_idx: np.ndarray, n_cols: int) -> torch.Tensor:
    """The dense fp32 [B, N] label tensor of the reference collate (dataset.py:873, 921)."""
    B = len(pos_ptr) - 1
    y = torch.zeros(B, n_cols)
    rows = np.repeat(np.arange(B), np.diff(pos_ptr))
    y[torch.from_numpy(rows), torch.from_numpy(np.asarray(pos_idx, np.int64))] = 1
    return y


def make_adagrad(model: torch.nn.Module, lr: float, weight_decay: float) -> torch.optim.Optimizer:
    """What OptimRegime effectively builds (utils/optim.py:29, 143-145): Adagrad with Adam's eps = 1e-8."""
    return torch.optim.Adagrad(model.parameters(), lr=lr, eps=1e-8, weight_decay=weight_decay)


def train_step(model: PortModel, opt, po, sp, labels: torch.Tensor, kind: str = "bce", smoothing: float = 0.0,
               candidate_ids: Optional[torch.Tensor] = None):
    """Trainer.compute_one_batch(training=True) (trainer.py:205-246)."""
    model.train()
    opt.zero_grad()
    scores = model(po, sp, candidate_ids)
    loss = loss_sum(scores, labels, kind, smoothing)
    (loss.sum() / (labels.size(0) * labels.size(1))).backward()
    opt.step()
    return loss.detach().item(), scores


def compute_metrics(filter_mask: torch.Tensor, label_ids: Sequence[Sequence[torch.Tensor]], predictions: torch.Tensor):
    """OneToNMentionRelationDataset.compute_metrics (dataset.py:423-453): the per-prefix Python loop with
    repeat / masked_fill_(-1e8) / lt / eq / sum. Returns (sum of 1/(rank+1), #answers, ranks)."""
    mrr_sum, count, ranks_all = 0.0, 0, []
    for pf, labels, pred in zip(filter_mask, label_ids, predictions):
        n = len(labels)
        rep = pred.unsqueeze(0).repeat(n, 1)
        frep = pf.unsqueeze(0).repeat(n, 1)
        true = torch.Tensor([pred[l.long()].max(0)[0] for l in labels])
        rep.masked_fill_(frep, -1e8)
        greater = (true.view(n, -1) < rep).long().sum(1)
        equal = (true.view(n, -1) == rep).long().sum(1)
        ranks = greater + equal // 2
        mrr_sum += (1.0 / (ranks + 1).float()).sum().item()
        count += n
        ranks_all.append(ranks)
    return mrr_sum, count, torch.cat(ranks_all) if ranks_all else torch.zeros(0, dtype=torch.long)


def eval_step(model: PortModel, po, sp, labels, filter_mask, label_ids, kind="bce"):
    """compute_one_batch(training=False) (trainer.py:259-272): forward + loss + compute_metrics."""
    model.eval()
    with torch.no_grad():
        scores = model(po, sp)
        loss = loss_sum(scores, labels, kind)
    mrr_sum, count, ranks = compute_metrics(filter_mask, label_ids, scores)
    return loss.item(), mrr_sum, count, ranks


def set_threads() -> int:
    n = os.cpu_count() or 1
    torch.set_num_threads(n)                     # openkge/trainer.py:136
    return n
