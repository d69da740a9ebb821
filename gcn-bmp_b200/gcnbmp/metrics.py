"""On-device evaluation metrics: the scores stay on the GPU instead of travelling to scikit-learn.

Mirrors the reference's evaluator extensions, which call scikit-learn per class and average (macro):
  training/extensions/roc_auc_evaluator.py:87-119   roc_auc_score(t, y, average='macro')
  training/extensions/prc_auc_evaluator.py:87-125   per class: auc(recall, precision) of precision_recall_curve, mean
  training/extensions/f1_evaluator.py:47-72         f1_score(t, round(y), average='macro')
  training/extensions/acc_evaluator.py:46-77        per class accuracy of round(y), mean
  training/extensions/{precision,recall}_evaluator.py:47-75   precision_score / recall_score(average='binary', pos_label=1)
  training/multilabel_extensions/*_evaluator.py     the same scores per label column (average='binary'), then the mean
`scores` (n, K) are probabilities (or any monotone scores for the AUCs), `labels` (n, K) are 0/1.  Ties are handled exactly
like scikit-learn's curves (one curve point per distinct score).  `mask` (n, K) bool, optional: entries to keep -- the
per-column form of the evaluators' `ignore_labels` (every column is scored over its own non-ignored rows, as if the rows were
dropped column by column before calling scikit-learn).  Everything is tensor algebra on the scores' device (one sort per
call); no host synchronisation until the caller reads the result."""
import torch


def _curves(scores, labels, mask=None):
    """Per column: cumulative true / false positives at the END of every run of equal scores (descending order).
    Returns (tp, fp, is_end, P, N) with shapes (n, K), (n, K), (n, K) bool, (K,), (K,).  Masked-out entries count neither as
    positives nor as negatives: the curve points they add repeat the previous point (zero area)."""
    s, order = torch.sort(scores, dim=0, descending=True, stable=True)
    y = torch.gather(labels.to(torch.float64), 0, order)
    w = torch.ones_like(y) if mask is None else torch.gather(mask.to(torch.float64), 0, order)
    tp = torch.cumsum(y * w, dim=0)
    fp = torch.cumsum((1.0 - y) * w, dim=0)
    is_end = torch.ones_like(s, dtype=torch.bool)
    is_end[:-1] = s[1:] != s[:-1]
    return tp, fp, is_end, tp[-1], fp[-1]


def roc_auc(scores, labels, mask=None):
    """Macro-averaged ROC AUC over the K columns (trapezoid over the distinct-threshold ROC curve)."""
    tp, fp, is_end, P, N = _curves(scores, labels, mask)
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


def prc_auc(scores, labels, mask=None):
    """Mean over columns of auc(recall, precision) of scikit-learn's precision_recall_curve (trapezoid, with the final
    (recall 0, precision 1) point)."""
    tp, fp, is_end, P, _ = _curves(scores, labels, mask)
    prec = torch.where(tp + fp > 0, tp / (tp + fp).clamp(min=1), torch.ones_like(tp))
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


def _confusion(scores, labels, mask=None):
    pred = torch.round(scores)
    y = labels.to(scores.dtype)
    keep = torch.ones_like(pred, dtype=torch.bool) if mask is None else mask.to(torch.bool)
    tp = ((pred == 1) & (y == 1) & keep).sum(dim=0).to(torch.float64)
    fp = ((pred == 1) & (y == 0) & keep).sum(dim=0).to(torch.float64)
    fn = ((pred == 0) & (y == 1) & keep).sum(dim=0).to(torch.float64)
    tn = ((pred == 0) & (y == 0) & keep).sum(dim=0).to(torch.float64)
    return tp, fp, fn, tn


def _macro(scores, labels, per_class, mask=None):
    """f1_evaluator.py:47-72 (`average='macro'`): over the K label columns for a multi-label indicator (K > 1; the multilabel
    evaluator scores every column with average='binary' and takes the mean -- the same number), over the two CLASSES {0, 1}
    for a single binary column (K == 1, (n,1) arrays are read as 1-D)."""
    tp, fp, fn, tn = _confusion(scores, labels, mask)
    pos = per_class(tp, fp, fn)
    if scores.shape[1] > 1:
        return pos.mean()
    neg = per_class(tn, fn, fp)              # class 0 as the positive class
    return (pos + neg).mean() * 0.5


def _binary(scores, labels, per_class, mask=None):
    """`average='binary', pos_label=1` ({precision,recall}_evaluator.py:47-75): the positive class only; per column and then
    the mean for K > 1 (multilabel_extensions/{precision,recall}_evaluator.py:71-82)."""
    tp, fp, fn, _ = _confusion(scores, labels, mask)
    return per_class(tp, fp, fn).mean()


def _safe_div(a, b):
    return torch.where(b > 0, a / b.clamp(min=1), torch.zeros_like(a))


def f1(scores, labels, mask=None):
    """Macro F1 of round(scores) (a class without predicted or true members scores 0, as scikit-learn does)."""
    return _macro(scores, labels, lambda tp, fp, fn: _safe_div(2 * tp, 2 * tp + fp + fn), mask)


def accuracy(scores, labels, mask=None):
    tp, fp, fn, tn = _confusion(scores, labels, mask)
    return ((tp + tn) / (tp + fp + fn + tn)).mean()


def precision(scores, labels, mask=None):
    """Positive-class precision tp / (tp + fp) of round(scores) (0 when nothing is predicted positive)."""
    return _binary(scores, labels, lambda tp, fp, fn: _safe_div(tp, tp + fp), mask)


def recall(scores, labels, mask=None):
    """Positive-class recall tp / (tp + fn) of round(scores)."""
    return _binary(scores, labels, lambda tp, fp, fn: _safe_div(tp, tp + fn), mask)


def evaluate(scores, labels, mask=None):
    """All six numbers of the reference's evaluator stack in one pass over device-resident predictions."""
    return dict(roc_auc=roc_auc(scores, labels, mask), prc_auc=prc_auc(scores, labels, mask), f1=f1(scores, labels, mask),
                accuracy=accuracy(scores, labels, mask), precision=precision(scores, labels, mask),
                recall=recall(scores, labels, mask))
