"""On-device evaluation metrics: the scores stay on the GPU instead of travelling to scikit-learn.

Mirrors the reference's evaluator extensions, which call scikit-learn per class and average (macro):
  training/extensions/roc_auc_evaluator.py:87-119   roc_auc_score(t, y, average='macro')
  training/extensions/prc_auc_evaluator.py:87-125   per class: auc(recall, precision) of precision_recall_curve, mean
  training/extensions/f1_evaluator.py:47-72         f1_score(t, round(y), average='macro')
  training/extensions/acc_evaluator.py:46-77        per class accuracy of round(y), mean
`scores` (n, K) are probabilities (or any monotone scores for the AUCs), `labels` (n, K) are 0/1.  Ties are handled exactly
like scikit-learn's curves (one curve point per distinct score).  Everything is tensor algebra on the scores' device
(one sort per call); no host synchronisation until the caller reads the result."""
import torch


def _curves(scores, labels):
    """Per column: cumulative true / false positives at the END of every run of equal scores (descending order).
    Returns (tp, fp, is_end, P, N) with shapes (n, K), (n, K), (n, K) bool, (K,), (K,)."""
    s, order = torch.sort(scores, dim=0, descending=True, stable=True)
    y = torch.gather(labels.to(torch.float64), 0, order)
    tp = torch.cumsum(y, dim=0)
    fp = torch.cumsum(1.0 - y, dim=0)
    is_end = torch.ones_like(s, dtype=torch.bool)
    is_end[:-1] = s[1:] != s[:-1]
    return tp, fp, is_end, tp[-1], fp[-1]


def roc_auc(scores, labels):
    """Macro-averaged ROC AUC over the K columns (trapezoid over the distinct-threshold ROC curve)."""
    tp, fp, is_end, P, N = _curves(scores, labels)
    # trapezoid over curve points = sum over runs of (fp_end - fp_prev_end) * (tp_end + tp_prev_end) / 2
    big = torch.zeros_like(tp)
    tp_e = torch.where(is_end, tp, big)
    fp_e = torch.where(is_end, fp, big)
    # previous run end: cumulative max works because tp, fp are non-decreasing
    tp_prev = torch.cat([torch.zeros_like(tp[:1]), torch.cummax(tp_e, dim=0).values[:-1]])
    fp_prev = torch.cat([torch.zeros_like(fp[:1]), torch.cummax(fp_e, dim=0).values[:-1]])
    area = torch.where(is_end, (fp - fp_prev) * (tp + tp_prev) * 0.5, big).sum(dim=0)
    auc = area / (P * N)
    if bool(((P == 0) | (N == 0)).any()):
        raise ValueError("Only one class present in y_true. ROC AUC score is not defined in that case.")
    return auc.mean()


def prc_auc(scores, labels):
    """Mean over columns of auc(recall, precision) of scikit-learn's precision_recall_curve (trapezoid, with the final
    (recall 0, precision 1) point)."""
    tp, fp, is_end, P, _ = _curves(scores, labels)
    prec = tp / (tp + fp)
    rec = tp / P
    big = torch.zeros_like(tp)
    # curve points ordered by DEcreasing threshold: (rec_k, prec_k) at run ends; preceded by (0, 1)
    rec_e = torch.where(is_end, rec, big)
    rec_prev = torch.cat([torch.zeros_like(rec[:1]), torch.cummax(rec_e, dim=0).values[:-1]])
    # precision at the previous run end: carry forward the last run-end value
    idx = torch.arange(tp.shape[0], device=tp.device).unsqueeze(1).expand_as(tp)
    last_end = torch.cummax(torch.where(is_end, idx, torch.full_like(idx, -1)), dim=0).values
    prev_end = torch.cat([torch.full_like(last_end[:1], -1), last_end[:-1]])
    prec_prev = torch.where(prev_end >= 0, torch.gather(prec, 0, prev_end.clamp(min=0)), torch.ones_like(prec))
    area = torch.where(is_end, (rec - rec_prev) * (prec + prec_prev) * 0.5, big).sum(dim=0)
    return area.mean()


def _confusion(scores, labels):
    pred = torch.round(scores)
    y = labels.to(scores.dtype)
    tp = ((pred == 1) & (y == 1)).sum(dim=0).to(torch.float64)
    fp = ((pred == 1) & (y == 0)).sum(dim=0).to(torch.float64)
    fn = ((pred == 0) & (y == 1)).sum(dim=0).to(torch.float64)
    tn = ((pred == 0) & (y == 0)).sum(dim=0).to(torch.float64)
    return tp, fp, fn, tn


def _macro(scores, labels, per_class):
    """scikit-learn's `average='macro'` as the reference's evaluators get it: over the K label columns for a multi-label
    indicator (K > 1), over the two CLASSES {0, 1} for a single binary column (K == 1, (n,1) arrays are read as 1-D)."""
    tp, fp, fn, tn = _confusion(scores, labels)
    pos = per_class(tp, fp, fn)
    if scores.shape[1] > 1:
        return pos.mean()
    neg = per_class(tn, fn, fp)              # class 0 as the positive class
    return (pos + neg).mean() * 0.5


def _safe_div(a, b):
    return torch.where(b > 0, a / b.clamp(min=1), torch.zeros_like(a))


def f1(scores, labels):
    """Macro F1 of round(scores) (a class without predicted or true members scores 0, as scikit-learn does)."""
    return _macro(scores, labels, lambda tp, fp, fn: _safe_div(2 * tp, 2 * tp + fp + fn))


def accuracy(scores, labels):
    tp, fp, fn, tn = _confusion(scores, labels)
    return ((tp + tn) / (tp + fp + fn + tn)).mean()


def precision(scores, labels):
    return _macro(scores, labels, lambda tp, fp, fn: _safe_div(tp, tp + fp))


def recall(scores, labels):
    return _macro(scores, labels, lambda tp, fp, fn: _safe_div(tp, tp + fn))


def evaluate(scores, labels):
    """All six numbers of the reference's evaluator stack in one pass over device-resident predictions."""
    return dict(roc_auc=roc_auc(scores, labels), prc_auc=prc_auc(scores, labels), f1=f1(scores, labels),
                accuracy=accuracy(scores, labels), precision=precision(scores, labels), recall=recall(scores, labels))
