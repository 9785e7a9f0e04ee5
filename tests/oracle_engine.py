"""CPU stand-in for the native op module, built on the oracle (numpy) and plain torch, used ONLY by the
tests to drive the host-side partition / reduction logic of ``sharded.py`` under gloo on a machine without
a GPU. Same function names and argument conventions as ``open_knowledge_graph_embeddings_b200.kernels``."""
import numpy as np
import torch

from oracle import okge_oracle as O


def _np(t):
    return t.detach().cpu().numpy()


def gather_rows(table, ids):
    return table[ids.reshape(-1).long()].clone()


def scatter_add_rows(grad, ids, grad_table, skip_id=-1):
    ids = ids.reshape(-1).long()
    keep = ids != skip_id
    grad_table.index_add_(0, ids[keep], grad[keep])


def gather_pool_fwd(tok_table, id_rows, ids, mode, id_start=0, n=None):
    rows = _np(id_rows).astype(np.int64)
    sel = np.arange(id_start, id_start + (n if n is not None else rows.shape[0] - id_start)) if ids is None else _np(ids)
    return torch.from_numpy(O.unigram_pool_encode(_np(tok_table), rows, sel, mode))


def gather_pool_bwd(grad_out, tok_table, id_rows, ids, mode, grad_tok_table, id_start=0):
    rows = _np(id_rows).astype(np.int64)
    sel = np.arange(id_start, id_start + grad_out.size(0)) if ids is None else _np(ids)
    grad_tok_table += torch.from_numpy(O.unigram_pool_backward(_np(grad_out), _np(tok_table), rows, sel, mode))


def bn_col_sums(a, x=None, mean=None, invstd=None):
    ad = a.double()
    second = ad * ad if x is None else ad * ((x.double() - mean.double()) * invstd.double())
    return torch.stack([ad.sum(0), second.sum(0)])


def bn_normalize(x, mean, invstd, gamma, beta):
    return (((x.double() - mean.double()) * invstd.double()) * gamma.double() + beta.double()).float()


def bn_normalize_bwd(dy, x, mean, invstd, coef, gamma):
    xhat = (x.double() - mean.double()) * invstd.double()
    return (gamma.double() * invstd.double() * (dy.double() - coef[0].double() - xhat * coef[1].double())).float()


def fold_query(kind, a, b):
    return torch.from_numpy(O.fold_query(kind, _np(a), _np(b)))


def fold_query_bwd(kind, a, b, gq):
    ga, gb = O.fold_query_backward(kind, _np(a), _np(b), _np(gq))
    return torch.from_numpy(ga), torch.from_numpy(gb)


TF32_RAW_OPERAND_SCALE = 1.0


def ColMajor(x):
    """row-major x[K, rows] used as the operand x^T (the native module reads it MN-major without a transpose)."""
    return x.t()


def quantize(x, split=False, out=None, fixed_scale=None):
    """The native module turns fp32 matrices into fp16 operands here; the oracle engine contracts in fp64 and keeps them
    as they are (a plain tensor stands in for the operand object). ``out`` is refreshed in place like the native one."""
    if out is not None:
        out.copy_(x)
        return out
    return x.clone()


def gather_rows_f16(op, ids):
    return gather_rows(op, ids)


def gemm_nt(a, b, alpha=1.0, alpha_dev=None, out=None, splits=None):
    r = (alpha * (a.double() @ b.double().t())).float()
    if out is not None:
        out.copy_(r)
        return out
    return r


def score_store(q, e, split=False):
    return (q.double() @ e.double().t()).float()


def _dense(ptr, idx, n):
    """Dense labels of a CSR whose negative columns mean "not in this column block" (sharded.restrict_csr)."""
    ptr, idx = _np(ptr).astype(np.int64), _np(idx).astype(np.int64)
    rows = np.repeat(np.arange(len(ptr) - 1), np.diff(ptr))
    y = np.zeros((len(ptr) - 1, n), np.float64)
    y[rows[idx >= 0], idx[idx >= 0]] = 1.0
    return torch.from_numpy(y)


def score_bce(q, e, pos_ptr, pos_idx, y_base=0.0, y_pos=1.0, want_dS=True, n_cols_dev=None):
    s = q.double() @ e.double().t()
    y = _dense(pos_ptr, pos_idx, e.size(0)) * (y_pos - y_base) + y_base
    loss = (torch.nn.functional.softplus(s) - s * y).sum().reshape(1)
    dS = (torch.sigmoid(s) - y).float()
    return loss, dS if want_dS else None


def score_lse(q, e, pos_ptr, pos_idx):
    s = q.double() @ e.double().t()
    rows = torch.repeat_interleave(torch.arange(q.size(0)), (pos_ptr[1:] - pos_ptr[:-1]).long())
    own = pos_idx >= 0
    pos = torch.where(own, s[rows, pos_idx.long().clamp(min=0)], torch.zeros((), dtype=s.dtype))
    return torch.logsumexp(s, dim=1).float(), pos.float()


def score_softmax_grad(q, e, pos_ptr, pos_idx, row_lse, row_weight):
    s = q.double() @ e.double().t()
    y = _dense(pos_ptr, pos_idx, e.size(0))
    dS = (row_weight.double()[:, None] * torch.exp(s - row_lse.double()[:, None]) - y).float()
    return dS


def adagrad_dense(param, grad, state_sum, clr, eps, weight_decay):
    p, s = O.adagrad_step(_np(param), _np(grad), _np(state_sum), clr, eps, weight_decay)
    param.copy_(torch.from_numpy(p))
    state_sum.copy_(torch.from_numpy(s))


def gemm_adagrad(a, b, param, state_sum, clr, eps, weight_decay, alpha=1.0, alpha_dev=None, extra_map=None, extra=None,
                 shadow=None):
    g = gemm_nt(a, b, alpha=alpha)
    if alpha_dev is not None:
        g = g * float(alpha_dev)
    if extra_map is not None:
        has = extra_map >= 0
        g[has] += extra[extra_map[has].long()]
    adagrad_dense(param, g, state_sum, clr, eps, weight_decay)
    if shadow is not None:
        shadow.copy_(param)


def row_slots_build(ids, slot_map, skip_id=-1):
    for i, r in enumerate(ids.reshape(-1).tolist()):
        if r != skip_id:
            slot_map[r] = max(int(slot_map[r]), i)


def row_slots_accumulate(grad, ids, slot_map, extra, skip_id=-1):
    ids = ids.reshape(-1).long()
    keep = ids != skip_id
    extra.index_add_(0, slot_map[ids[keep]].long(), grad[keep])


def row_slots_clear(ids, slot_map, skip_id=-1):
    ids = ids.reshape(-1).long()
    slot_map[ids[ids != skip_id]] = -1


def rank_true_score(sel, ans_row, alt_ptr, alt_pos, true):
    for j in range(ans_row.numel()):
        for a in range(int(alt_ptr[j]), int(alt_ptr[j + 1])):
            if alt_pos[a] >= 0:
                true[j] = max(float(true[j]), float(sel[int(ans_row[j]), int(alt_pos[a])]))


def score_rank(q, e, thresh, greater, equal, split=False):
    s = score_store(q, e)
    greater += (thresh[:, None] < s).sum(1).int()
    equal += (thresh[:, None] == s).sum(1).int()


def rank_filter_correct(sel, ans_row, filt_ptr, filt_pos, thresh, greater, equal, add_mask_terms=True):
    for j in range(ans_row.numel()):
        b = int(ans_row[j])
        t = thresh[j]
        for f in range(int(filt_ptr[b]), int(filt_ptr[b + 1])):
            p = int(filt_pos[f])
            if p >= 0:
                v = sel[b, p]
                greater[j] -= int(t < v)
                equal[j] -= int(t == v)
        if add_mask_terms:
            nf = int(filt_ptr[b + 1] - filt_ptr[b])
            greater[j] += nf * int(t < -1e8)
            equal[j] += nf * int(t == -1e8)
