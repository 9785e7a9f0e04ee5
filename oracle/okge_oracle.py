"""CPU oracle: a numpy restatement of the reference's algorithm for the OpenKGE hot path.

TEST INFRASTRUCTURE ONLY. Nothing in the product package imports this module; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do,
and there only as the checker / the timed CPU baseline, never as the thing shipped.

Parity status: PINNED. The reference (samuelbroscheit/open_knowledge_graph_embeddings) ships no
tests or golden vectors of its own, so the pins are outputs of the unmodified
reference run in the build container: ``tests/golden/*.npz`` written by
``tests/golden/make_golden.py``; ``tests/test_oracle_golden.py`` checks every function below against
them (scores / loss / gradients / post-step weights to fp32 round-off, rank counts bit-exact).

Every function cites the reference file:line it follows (paths relative to the reference root).
All arithmetic is float32 like the reference unless a dtype is passed.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

PAD, UNK, BOS, EOS = 0, 1, 2, 3  # openkge/index_mapper.py:14
MASK_FILL = np.float32(-1e8)     # openkge/dataset.py:440


# ---------------------------------------------------------------------------------------------
# ragged-list wire format — utils/misc.py:56-89
# ---------------------------------------------------------------------------------------------

def pack_list_of_lists(lol) -> List[int]:
    """utils/misc.py:56-70. [[5],[6,7],[8]] -> [5,6,8,9,0,5,6,7,8]: absolute offsets of each
    sub-list inside the packed array, a 0 terminator, then the values."""
    offsets, values = [0], []
    for item in lol:
        if isinstance(item, (list, tuple)):
            values.extend(item)
        else:
            values.append(item)
        offsets.append(len(values))
    offsets.append(-len(offsets) - 1)           # becomes the 0 terminator after the shift below
    shift = len(offsets)
    return [o + shift for o in offsets] + values


def unpack_list_of_lists(ents: Sequence[int]) -> Tuple[List[List[int]], List[int]]:
    """utils/misc.py:72-89. Inverse of pack_list_of_lists: (list of sub-lists, flat values)."""
    ents = list(int(x) for x in ents)
    out, end, all_begin, all_end = [], -1, -1, -1
    for off in ents:
        if all_begin == -1:
            all_begin = off
        if off == 0:
            break
        if end == -1:
            end = off
            continue
        begin, end = end, off
        all_end = off
        out.append(ents[begin:end])
    return out, ents[all_begin:all_end]


# ---------------------------------------------------------------------------------------------
# embedders
# ---------------------------------------------------------------------------------------------

def build_token_id_rows(id_to_tokens: Sequence[Sequence[int]], max_len: int) -> np.ndarray:
    """openkge/model.py:576-595. Row i = the LAST max_len tokens of entity/relation i, left-aligned,
    zero (PAD) padded. int64 like the reference buffer."""
    rows = np.zeros((len(id_to_tokens), max_len), dtype=np.int64)
    for i, toks in enumerate(id_to_tokens):
        t = list(toks)[-max_len:]
        rows[i, : len(t)] = t
    return rows


def lookup_encode(weight: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """openkge/model.py:455-458 with dropout = 0, no batch norm / projection / normalize."""
    return weight[np.asarray(ids, dtype=np.int64).reshape(-1)]


def unigram_pool_encode(weight: np.ndarray, token_rows: np.ndarray, ids: np.ndarray, pool: str = "sum") -> np.ndarray:
    """openkge/model.py:762-774. All L slots are gathered, PAD slots included (W[0] is an ordinary
    trained-from-init row, :633-634); mean divides by (#tok > 0) + 1e-12."""
    tok = token_rows[np.asarray(ids, dtype=np.int64).reshape(-1)]          # [n, L]   :763
    emb = weight[tok]                                                       # [n, L, D] :767
    if pool == "max":
        return emb.max(axis=1)                                              # :768-769
    if pool == "mean":
        lengths = (tok > 0).astype(weight.dtype).sum(axis=1, keepdims=True)  # :771
        return emb.sum(axis=1, dtype=weight.dtype) / (lengths + weight.dtype.type(1e-12))  # :772
    return emb.sum(axis=1, dtype=weight.dtype)                              # :774


def unigram_pool_backward(grad_out: np.ndarray, weight: np.ndarray, token_rows: np.ndarray, ids: np.ndarray,
                          pool: str = "sum") -> np.ndarray:
    """Autograd of unigram_pool_encode w.r.t. the token table. Token 0 is padding_idx and never
    receives gradient (openkge/model.py:597-608)."""
    tok = token_rows[np.asarray(ids, dtype=np.int64).reshape(-1)]
    n, L = tok.shape
    grad_w = np.zeros_like(weight)
    if pool == "max":
        emb = weight[tok]
        arg = emb.argmax(axis=1)                                            # first maximum, like torch.max
        for l in range(L):
            sel = (arg == l)
            contrib = np.where(sel, grad_out, 0).astype(weight.dtype)
            keep = tok[:, l] != 0
            np.add.at(grad_w, tok[keep, l], contrib[keep])
        return grad_w
    g = grad_out
    if pool == "mean":
        lengths = (tok > 0).astype(weight.dtype).sum(axis=1, keepdims=True)
        g = grad_out / (lengths + weight.dtype.type(1e-12))
    for l in range(L):
        keep = tok[:, l] != 0
        np.add.at(grad_w, tok[keep, l], g[keep])
    return grad_w


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def lstm_last_state_encode(weight: np.ndarray, token_rows: np.ndarray, ids: np.ndarray, w_ih: np.ndarray,
                           w_hh: np.ndarray, b_ih: np.ndarray, b_hh: np.ndarray):
    """LSTMRelationEmbedder._encode / _encode_tokens (openkge/model.py:963-987): ids -> token rows (:957-961) ->
    embeddings [n, L, D] -> single-layer torch.nn.LSTM (gate order i, f, g, o; h0 = c0 = 0) -> the output at the LAST
    REAL token of every row, ``last_state = (tokens > 0).sum(1) - 1`` (:970; -1 wraps to the final step like the
    reference's advanced indexing). Returns (encoded [n, D] float32, cache for the backward)."""
    ids = np.asarray(ids, np.int64).reshape(-1)
    tok = np.asarray(token_rows)[ids].astype(np.int64)                   # [n, L]
    n, L = tok.shape
    D = w_hh.shape[1]
    x = np.asarray(weight, np.float64)[tok]                              # [n, L, D]; PAD slots gather row 0 too
    w_ih, w_hh = np.asarray(w_ih, np.float64), np.asarray(w_hh, np.float64)
    bias = np.asarray(b_ih, np.float64) + np.asarray(b_hh, np.float64)
    last = ((tok > 0).sum(1) - 1) % L
    h, c = np.zeros((n, D)), np.zeros((n, D))
    steps, out = [], np.zeros((n, D))
    for t in range(L):
        pre = x[:, t] @ w_ih.T + h @ w_hh.T + bias
        i, f, g, o = _sigmoid(pre[:, :D]), _sigmoid(pre[:, D:2 * D]), np.tanh(pre[:, 2 * D:3 * D]), _sigmoid(pre[:, 3 * D:])
        c_new = f * c + i * g
        h_new = o * np.tanh(c_new)
        steps.append((h, c, i, f, g, o, c_new))
        h, c = h_new, c_new
        out[last == t] = h[last == t]
    return out.astype(np.float32), (tok, x, last, steps, w_ih, w_hh, weight.shape)


def lstm_last_state_backward(grad_out: np.ndarray, cache):
    """Back-propagation through time of ``lstm_last_state_encode``: (grad of the token table [V, D] with the PAD row
    left at zero (padding_idx = 0, openkge/model.py:597-608), dW_ih, dW_hh, db_ih, db_hh)."""
    tok, x, last, steps, w_ih, w_hh, wshape = cache
    n, L = tok.shape
    D = w_hh.shape[1]
    go = np.asarray(grad_out, np.float64)
    dh, dc = np.zeros((n, D)), np.zeros((n, D))
    dW_ih, dW_hh, db = np.zeros_like(w_ih), np.zeros_like(w_hh), np.zeros(4 * D)
    dtable = np.zeros(wshape, np.float64)
    for t in range(L - 1, -1, -1):
        h_prev, c_prev, i, f, g, o, c_new = steps[t]
        dh = dh + np.where((last == t)[:, None], go, 0.0)
        tc = np.tanh(c_new)
        do = dh * tc
        dc = dc + dh * o * (1 - tc * tc)
        dpre = np.concatenate([dc * g * i * (1 - i), dc * c_prev * f * (1 - f), dc * i * (1 - g * g), do * o * (1 - o)], 1)
        dW_ih += dpre.T @ x[:, t]
        dW_hh += dpre.T @ h_prev
        db += dpre.sum(0)
        dx = dpre @ w_ih
        np.add.at(dtable, tok[:, t], dx)
        dh = dpre @ w_hh
        dc = dc * f
    dtable[0] = 0.0
    f32 = np.float32
    return dtable.astype(f32), dW_ih.astype(f32), dW_hh.astype(f32), db.astype(f32), db.astype(f32)


def batchnorm_train(x: np.ndarray, gamma: np.ndarray, beta: np.ndarray, eps: float = 1e-5):
    """torch.nn.BatchNorm1d in training mode (openkge/model.py:610-615, 779-780): batch statistics over
    the rows, biased variance for normalisation. Returns (y, cache) with what backward needs."""
    mean = x.mean(axis=0, dtype=np.float64)
    var = x.astype(np.float64).var(axis=0)
    inv = 1.0 / np.sqrt(var + eps)
    xhat = (x - mean) * inv
    y = (xhat * gamma + beta).astype(x.dtype)
    return y, (xhat, inv, gamma, mean, var)


def batchnorm_train_backward(grad_y: np.ndarray, cache) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    xhat, inv, gamma, _, _ = cache
    n = grad_y.shape[0]
    g = grad_y.astype(np.float64)
    dgamma = (g * xhat).sum(axis=0)
    dbeta = g.sum(axis=0)
    dxhat = g * gamma
    dx = (inv / n) * (n * dxhat - dxhat.sum(axis=0) - xhat * (dxhat * xhat).sum(axis=0))
    return dx.astype(grad_y.dtype), dgamma.astype(grad_y.dtype), dbeta.astype(grad_y.dtype)


def batchnorm_eval(x, gamma, beta, running_mean, running_var, eps: float = 1e-5):
    return ((x - running_mean) / np.sqrt(running_var + eps) * gamma + beta).astype(x.dtype)


# ---------------------------------------------------------------------------------------------
# scorers — openkge/model.py:181-278
# ---------------------------------------------------------------------------------------------

def complex_prefix_score(subj: np.ndarray, rel: np.ndarray, obj: np.ndarray, sp: bool) -> np.ndarray:
    """ComplexRelationScorer._score(prefix=True), the reference's own 4-product form
    (openkge/model.py:200-215); halves by chunk(2, dim=1): first D/2 real, last D/2 imaginary."""
    h = rel.shape[1] // 2
    r1, r2 = rel[:, :h], rel[:, h:]
    s1, s2 = subj[:, :h], subj[:, h:]
    o1, o2 = obj[:, :h], obj[:, h:]
    if sp:   # :206-209
        return (s1 * r1) @ o1.T + (s2 * r1) @ o2.T + (s1 * r2) @ o2.T - (s2 * r2) @ o1.T
    # po   :212-215
    return (o1 * r1) @ s1.T + (o2 * r1) @ s2.T + (o2 * r2) @ s1.T - (o1 * r2) @ s2.T


def distmult_prefix_score(subj: np.ndarray, rel: np.ndarray, obj: np.ndarray, sp: bool) -> np.ndarray:
    """DistmultRelationScorer._score(prefix=True), openkge/model.py:268-272."""
    return (subj * rel) @ obj.T if sp else (rel * obj) @ subj.T


FOLD_COMPLEX_SP, FOLD_COMPLEX_PO, FOLD_DISTMULT = 0, 1, 2


def fold_query(kind: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """The same prefix scores written as one row vector q with score = q @ E^T:
    COMPLEX_SP a = subj, b = rel; COMPLEX_PO a = obj, b = rel; DISTMULT a = subj|obj, b = rel."""
    if kind == FOLD_DISTMULT:
        return a * b
    h = a.shape[1] // 2
    a1, a2, b1, b2 = a[:, :h], a[:, h:], b[:, :h], b[:, h:]
    if kind == FOLD_COMPLEX_SP:
        return np.concatenate([a1 * b1 - a2 * b2, a2 * b1 + a1 * b2], axis=1)
    return np.concatenate([a1 * b1 + a2 * b2, a2 * b1 - a1 * b2], axis=1)


def fold_query_backward(kind: int, a: np.ndarray, b: np.ndarray, gq: np.ndarray):
    if kind == FOLD_DISTMULT:
        return gq * b, gq * a
    h = a.shape[1] // 2
    a1, a2, b1, b2, g1, g2 = a[:, :h], a[:, h:], b[:, :h], b[:, h:], gq[:, :h], gq[:, h:]
    if kind == FOLD_COMPLEX_SP:
        ga = np.concatenate([g1 * b1 + g2 * b2, -g1 * b2 + g2 * b1], axis=1)
        gb = np.concatenate([g1 * a1 + g2 * a2, -g1 * a2 + g2 * a1], axis=1)
    else:
        ga = np.concatenate([g1 * b1 - g2 * b2, g1 * b2 + g2 * b1], axis=1)
        gb = np.concatenate([g1 * a1 + g2 * a2, g1 * a2 - g2 * a1], axis=1)
    return ga, gb


# ---------------------------------------------------------------------------------------------
# losses — openkge/trainer.py:91-106
# ---------------------------------------------------------------------------------------------

def dense_labels(pos_ptr: np.ndarray, pos_idx: np.ndarray, n_cols: int, dtype=np.float32) -> np.ndarray:
    y = np.zeros((len(pos_ptr) - 1, n_cols), dtype=dtype)
    for b in range(len(pos_ptr) - 1):
        y[b, pos_idx[pos_ptr[b]:pos_ptr[b + 1]]] = 1
    return y


def smooth_labels(labels: np.ndarray, eps: float) -> np.ndarray:
    """openkge/trainer.py:103-105: y <- (y + 1/N) * (1 - eps)."""
    if eps <= 0:
        return labels
    return ((labels + labels.dtype.type(1.0 / labels.shape[-1])) * labels.dtype.type(1 - eps)).astype(labels.dtype)


def bce_with_logits_sum(scores: np.ndarray, labels: np.ndarray) -> float:
    """BCEWithLogitsLoss(reduction='sum') (openkge/trainer.py:106) = sum softplus(s) - s*y."""
    s = scores.astype(np.float64)
    return float((np.maximum(s, 0) + np.log1p(np.exp(-np.abs(s))) - s * labels).sum())


def bce_with_logits_grad(scores: np.ndarray, labels: np.ndarray) -> np.ndarray:
    s = scores.astype(np.float64)
    return (1.0 / (1.0 + np.exp(-s)) - labels).astype(scores.dtype)


def log_softmax_rows(scores: np.ndarray) -> np.ndarray:
    s = scores.astype(np.float64)
    m = s.max(axis=1, keepdims=True)
    return s - (m + np.log(np.exp(s - m).sum(axis=1, keepdims=True)))


def kl_log_softmax_sum(scores: np.ndarray, labels: np.ndarray) -> float:
    """KLDivLoss(reduction='sum')(log_softmax(scores, 1), labels) (openkge/trainer.py:99-100, 106);
    with 0/1 labels xlogy(y, y) = 0, so this is -sum_{y=1} log_softmax."""
    lp = log_softmax_rows(scores)
    y = labels.astype(np.float64)
    ent = np.where(y > 0, y * np.log(np.where(y > 0, y, 1.0)), 0.0)
    return float((ent - y * lp).sum())


def kl_log_softmax_grad(scores: np.ndarray, labels: np.ndarray) -> np.ndarray:
    lp = log_softmax_rows(scores)
    y = labels.astype(np.float64)
    return (y.sum(axis=1, keepdims=True) * np.exp(lp) - y).astype(scores.dtype)


# ---------------------------------------------------------------------------------------------
# filtered ranking — openkge/dataset.py:423-453
# ---------------------------------------------------------------------------------------------

def rank_counts(scores: np.ndarray, ans_row, alt_ptr, alt_idx, filt_ptr, filt_idx):
    """Inner quantities of compute_metrics for Q ranked answers in CSR form:
    true = max over the answer's alternative mentions of the UNMASKED row (:436-438);
    masked row = filter positions set to -1e8 (:440); greater = #(true < masked) (:441-443);
    equal = #(true == masked) (:444). float32 comparisons, integer counts."""
    scores = np.asarray(scores, dtype=np.float32)
    Q = len(ans_row)
    true = np.empty(Q, np.float32)
    greater = np.empty(Q, np.int64)
    equal = np.empty(Q, np.int64)
    for j in range(Q):
        b = int(ans_row[j])
        row = scores[b]
        t = row[alt_idx[alt_ptr[j]:alt_ptr[j + 1]]].max()
        masked = row.copy()
        masked[filt_idx[filt_ptr[b]:filt_ptr[b + 1]]] = MASK_FILL
        true[j] = t
        greater[j] = int((t < masked).sum())
        equal[j] = int((t == masked).sum())
    return true, greater, equal


class Meter:
    """utils/metrics.py:4-30 AccumulateMeter: count-weighted running average."""

    def __init__(self):
        self.avg, self.count = 0.0, 0

    def update(self, val, n=1):
        self.avg = (self.avg * self.count + val * n) / (self.count + n)
        self.count += n


def compute_metrics(scores: np.ndarray, ans_row, alt_ptr, alt_idx, filt_ptr, filt_idx) -> Dict[str, Meter]:
    """compute_metrics (openkge/dataset.py:423-453): rank = greater + equal // 2 (:445); per PREFIX the
    means of 1/(rank+1), rank, rank<50/10/3/1 are accumulated weighted by that prefix's answer count
    (:446-452)."""
    _, greater, equal = rank_counts(scores, ans_row, alt_ptr, alt_idx, filt_ptr, filt_idx)
    ranks = greater + equal // 2
    return metrics_from_ranks(ranks, ans_row)


def metrics_from_ranks(ranks: np.ndarray, ans_row) -> Dict[str, Meter]:
    res = OrderedDict((k, Meter()) for k in ("loss", "h1", "h3", "h10", "h50", "mrr", "mr"))
    ans_row = np.asarray(ans_row)
    ranks = np.asarray(ranks)
    # the reference walks prefixes in batch order; answers of a prefix are contiguous in ans_row
    start = 0
    while start < len(ans_row):
        end = start
        while end < len(ans_row) and ans_row[end] == ans_row[start]:
            end += 1
        r = ranks[start:end]
        n = end - start
        # the reference evaluates these in float32 tensors and takes .item()
        res["mrr"].update(float((np.float32(1.0) / (r + 1).astype(np.float32)).sum(dtype=np.float32)) / n, n)
        res["mr"].update(float(r.sum()) / n, n)
        res["h50"].update(float((r < 50).sum()) / n, n)
        res["h10"].update(float((r < 10).sum()) / n, n)
        res["h3"].update(float((r < 3).sum()) / n, n)
        res["h1"].update(float((r < 1).sum()) / n, n)
        start = end
    return res


# ---------------------------------------------------------------------------------------------
# optimizers — utils/optim.py:139-160, 194-201 + torch.optim formulas
# ---------------------------------------------------------------------------------------------

def adagrad_step(param, grad, state_sum, lr, eps=1e-8, weight_decay=0.0, lr_decay=0.0, step=1):
    """torch.optim.Adagrad dense step as the reference effectively runs it: the regime rebuilds the
    optimizer on Adam(lr=0)'s param_groups (utils/optim.py:29, 143-145), hence eps = 1e-8.
    Returns (param, state_sum) updated; float32 arithmetic."""
    f = param.dtype.type
    g = grad + f(weight_decay) * param if weight_decay != 0 else grad
    clr = f(lr / (1 + (step - 1) * lr_decay))
    state_sum = state_sum + g * g
    std = np.sqrt(state_sum) + f(eps)
    return param - clr * (g / std), state_sum


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, step, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """torch.optim.Adam (no amsgrad) single-tensor formula."""
    f = param.dtype.type
    g = grad + f(weight_decay) * param if weight_decay != 0 else grad
    exp_avg = exp_avg + (g - exp_avg) * f(1 - beta1)
    exp_avg_sq = exp_avg_sq * f(beta2) + (g * g) * f(1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = np.sqrt(exp_avg_sq) / f(np.sqrt(bc2)) + f(eps)
    return param - f(lr / bc1) * (exp_avg / denom), exp_avg, exp_avg_sq


# ---------------------------------------------------------------------------------------------
# collate — openkge/dataset.py:724-940 (1-vs-all mode, use_batch_shared_entities = False)
# ---------------------------------------------------------------------------------------------

def collate_full(rows: np.ndarray, seen_entities: np.ndarray, all_splits_entities: np.ndarray,
                 entity_vocab_size: int, entity_vocab_offset: int, is_training_data: bool):
    """Restatement of OneToNMentionRelationDataset_collate_func for the 1-vs-all branch
    (openkge/dataset.py:869-876, 885-935), emitting the sparse form of the same tensors:
    rows are grouped po (slot 0) first then sp (slot 2) (:885-932); label columns are entity id -
    offset (:921); eval additionally yields the per-answer alternative lists (:923-926) and the
    all-splits filter set (:927). Returns a dict."""
    groups = {0: [], 2: []}
    for r in rows:
        a, b, ts, te, as_, ae, slot = (int(x) for x in r)
        lol, flat = unpack_list_of_lists(seen_entities[ts:te])                 # :779-780
        groups[slot].append(((a, b), lol, flat, [int(x) for x in all_splits_entities[as_:ae]]))
    n_cols = entity_vocab_size - entity_vocab_offset
    out = {"po": np.zeros((len(groups[0]), 2), np.int32), "sp": np.zeros((len(groups[2]), 2), np.int32)}
    pos_ptr, pos_idx = [0], []
    filt_ptr, filt_idx = [0], []
    ans_row, alt_ptr, alt_idx = [], [0], []
    b_off = 0
    for slot, key in ((0, "po"), (2, "sp")):
        for i, (pref, lol, flat, allsp) in enumerate(groups[slot]):
            out[key][i] = pref
            cols = sorted(set(e - entity_vocab_offset for e in flat))           # label_tensor[...] = 1, :921
            pos_idx.extend(cols)
            pos_ptr.append(len(pos_idx))
            if not is_training_data:
                for alt in lol:                                                 # :923-926
                    ans_row.append(b_off)
                    alt_idx.extend(e - entity_vocab_offset for e in alt)
                    alt_ptr.append(len(alt_idx))
                fcols = sorted(set(e - entity_vocab_offset for e in allsp))     # :927
                filt_idx.extend(fcols)
                filt_ptr.append(len(filt_idx))
            b_off += 1
    out["pos_ptr"] = np.asarray(pos_ptr, np.int32)
    out["pos_idx"] = np.asarray(pos_idx, np.int32)
    out["normalizer_metric"] = float(len(pos_idx))                              # labels.sum(), :934
    out["normalizer_loss"] = b_off * n_cols                                     # :935
    if not is_training_data:
        out["filt_ptr"] = np.asarray(filt_ptr, np.int32)
        out["filt_idx"] = np.asarray(filt_idx, np.int32)
        out["ans_row"] = np.asarray(ans_row, np.int32)
        out["alt_ptr"] = np.asarray(alt_ptr, np.int32)
        out["alt_idx"] = np.asarray(alt_idx, np.int32)
    return out


# ---------------------------------------------------------------------------------------------
# whole-step oracle: AddLossModule.forward + backward (openkge/trainer.py:48-113, 181-246)
# ---------------------------------------------------------------------------------------------

class OracleModel:
    """Forward / backward of the model families on the hot path in numpy, 1-vs-all mode,
    dropout = 0: LookupDistmult, LookupComplex (openkge/model.py:1006-1014),
    UnigramPoolingComplex (:1016-1019) and LSTMComplex / LSTMDistmult (:1026-1034), the token models with optional
    batch norm (normalize='batchnorm')."""

    def __init__(self, kind: str, scorer: str, params: Dict[str, np.ndarray], pool: str = "sum",
                 batchnorm: bool = False, min_size: int = 2):
        assert kind in ("lookup", "unigram", "lstm") and scorer in ("complex", "distmult")
        self.kind, self.scorer, self.pool, self.batchnorm, self.min_size = kind, scorer, pool, batchnorm, min_size
        self.p = {k: np.array(v, copy=True) for k, v in params.items()}

    # -- encoders -------------------------------------------------------------------------
    def _encode(self, which: str, ids: np.ndarray, training: bool, tape: Optional[list]):
        w = self.p[f"{which}_embedding.weight"]
        if self.kind == "lookup":
            x = lookup_encode(w, ids)
            if tape is not None:
                tape.append(("lookup", which, np.asarray(ids, np.int64).reshape(-1), None))
            return x
        rows = self.p[f"{which}_token_ids"]
        lstm_cache = None
        if self.kind == "lstm":
            e = f"{which}_encoder_in."
            x, lstm_cache = lstm_last_state_encode(w, rows, ids, self.p[e + "weight_ih_l0"], self.p[e + "weight_hh_l0"],
                                                   self.p[e + "bias_ih_l0"], self.p[e + "bias_hh_l0"])
        else:
            x = unigram_pool_encode(w, rows, ids, self.pool)
        cache = None
        if self.batchnorm:
            g, b = self.p[f"{which}_batchnorm.weight"], self.p[f"{which}_batchnorm.bias"]
            if training:
                x, cache = batchnorm_train(x, g, b)
                n = x.shape[0]
                mean, var = cache[3], cache[4]
                rm, rv = self.p[f"{which}_batchnorm.running_mean"], self.p[f"{which}_batchnorm.running_var"]
                self.p[f"{which}_batchnorm.running_mean"] = (0.9 * rm + 0.1 * mean).astype(rm.dtype)
                self.p[f"{which}_batchnorm.running_var"] = (0.9 * rv + 0.1 * var * n / max(n - 1, 1)).astype(rv.dtype)
                self.p[f"{which}_batchnorm.num_batches_tracked"] = self.p[f"{which}_batchnorm.num_batches_tracked"] + 1
            else:
                x = batchnorm_eval(x, g, b, self.p[f"{which}_batchnorm.running_mean"],
                                   self.p[f"{which}_batchnorm.running_var"])
        if tape is not None:
            tape.append((self.kind, which, np.asarray(ids, np.int64).reshape(-1), (cache, lstm_cache)))
        return x

    def all_entities(self, training: bool, tape=None, candidate_ids=None):
        """Candidate matrix: every real entity (1-vs-all, openkge/dataset.py:872) or, in batch-shared mode,
        precompute_batch_shared_inputs(candidate_ids) (openkge/trainer.py:80-82)."""
        n_ent = (self.p["entity_embedding.weight"].shape[0] if self.kind == "lookup"
                 else self.p["entity_token_ids"].shape[0])
        ids = np.arange(self.min_size, n_ent) if candidate_ids is None else np.asarray(candidate_ids).reshape(-1)
        return self._encode("entity", ids, training, tape)

    def _score(self, subj, rel, obj, sp):
        fn = complex_prefix_score if self.scorer == "complex" else distmult_prefix_score
        return fn(subj, rel, obj, sp)

    # -- forward only ---------------------------------------------------------------------
    def scores(self, po_rel, po_obj, sp_subj, sp_rel, training: bool = False) -> np.ndarray:
        """AddLossModule.forward scoring part (openkge/trainer.py:69-91): candidate matrix once, po block
        then sp block, concatenated."""
        E = self.all_entities(training)
        rel_po = self._encode("relation", po_rel, training, None)
        obj_po = self._encode("entity", po_obj, training, None)
        s_po = self._score(E, rel_po, obj_po, sp=False)
        subj_sp = self._encode("entity", sp_subj, training, None)
        rel_sp = self._encode("relation", sp_rel, training, None)
        s_sp = self._score(subj_sp, rel_sp, E, sp=True)
        return np.concatenate([s_po, s_sp], axis=0)

    def operands(self, po_rel, po_obj, sp_subj, sp_rel, training: bool = False):
        """(Q, E): the folded query rows (po block first) and the candidate matrix whose product is
        ``scores`` — used by the tests for the norm-wise tolerance |ds| <= tol * ||q|| * ||e||.
        Call on a scratch copy when training=True and batch norm is on (running stats are updated)."""
        E = self.all_entities(training)
        rel_po = self._encode("relation", po_rel, training, None)
        obj_po = self._encode("entity", po_obj, training, None)
        subj_sp = self._encode("entity", sp_subj, training, None)
        rel_sp = self._encode("relation", sp_rel, training, None)
        kind_po = FOLD_COMPLEX_PO if self.scorer == "complex" else FOLD_DISTMULT
        kind_sp = FOLD_COMPLEX_SP if self.scorer == "complex" else FOLD_DISTMULT
        return np.concatenate([fold_query(kind_po, obj_po, rel_po), fold_query(kind_sp, subj_sp, rel_sp)], 0), E

    # -- forward + backward -----------------------------------------------------------------
    def loss_and_grads(self, po_rel, po_obj, sp_subj, sp_rel, pos_ptr, pos_idx, loss: str = "bce",
                       smoothing: float = 0.0, candidate_ids=None):
        """Returns (scores, loss_sum, grads dict) where grads are d(loss_sum / (B*N)) / d(param), i.e. what
        Trainer.compute_one_batch back-propagates (openkge/trainer.py:217-234)."""
        tape: list = []
        E = self.all_entities(True, tape, candidate_ids)
        rel_po = self._encode("relation", po_rel, True, tape)
        obj_po = self._encode("entity", po_obj, True, tape)
        subj_sp = self._encode("entity", sp_subj, True, tape)
        rel_sp = self._encode("relation", sp_rel, True, tape)
        kind_po = FOLD_COMPLEX_PO if self.scorer == "complex" else FOLD_DISTMULT
        kind_sp = FOLD_COMPLEX_SP if self.scorer == "complex" else FOLD_DISTMULT
        q_po = fold_query(kind_po, obj_po, rel_po)
        q_sp = fold_query(kind_sp, subj_sp, rel_sp)
        scores = np.concatenate([self._score(E, rel_po, obj_po, False), self._score(subj_sp, rel_sp, E, True)], 0)
        B, N = scores.shape
        y = dense_labels(pos_ptr, pos_idx, N)
        if loss == "bce":
            y = smooth_labels(y, smoothing)
            loss_sum = bce_with_logits_sum(scores, y)
            dS = bce_with_logits_grad(scores, y)
        else:
            loss_sum = kl_log_softmax_sum(scores, y)
            dS = kl_log_softmax_grad(scores, y)
        dS = (dS.astype(np.float64) / (B * N))
        Q = np.concatenate([q_po, q_sp], 0).astype(np.float64)
        dQ = (dS @ E.astype(np.float64)).astype(np.float32)
        dE = (dS.T @ Q).astype(np.float32)
        b_po = len(np.asarray(po_rel).reshape(-1))
        g_obj_po, g_rel_po = fold_query_backward(kind_po, obj_po, rel_po, dQ[:b_po])
        g_subj_sp, g_rel_sp = fold_query_backward(kind_sp, subj_sp, rel_sp, dQ[b_po:])
        out_grads = [dE, g_rel_po, g_obj_po, g_subj_sp, g_rel_sp]
        grads = {k: np.zeros_like(v) for k, v in self.p.items()
                 if k.endswith("embedding.weight") or k.endswith("batchnorm.weight") or k.endswith("batchnorm.bias")
                 or "_encoder_in." in k}
        for (enc, which, ids, cache), g in zip(tape, out_grads):
            key = f"{which}_embedding.weight"
            if enc == "lookup":
                np.add.at(grads[key], ids, g)
            else:
                bn_cache, lstm_cache = cache
                if bn_cache is not None:
                    g, dgamma, dbeta = batchnorm_train_backward(g, bn_cache)
                    grads[f"{which}_batchnorm.weight"] += dgamma
                    grads[f"{which}_batchnorm.bias"] += dbeta
                if enc == "lstm":
                    dtab, dwi, dwh, dbi, dbh = lstm_last_state_backward(g, lstm_cache)
                    e = f"{which}_encoder_in."
                    grads[key] += dtab
                    grads[e + "weight_ih_l0"] += dwi
                    grads[e + "weight_hh_l0"] += dwh
                    grads[e + "bias_ih_l0"] += dbi
                    grads[e + "bias_hh_l0"] += dbh
                else:
                    grads[key] += unigram_pool_backward(g, self.p[key], self.p[f"{which}_token_ids"], ids, self.pool)
        return scores, loss_sum, grads
