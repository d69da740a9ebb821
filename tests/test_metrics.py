"""On-device evaluation metrics against scikit-learn called the way the reference's evaluators call it
(training/extensions/*_evaluator.py): macro averages over the label columns, ties included."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gcn-bmp_b200"))
sk = pytest.importorskip("sklearn.metrics")


def _load():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bmp_metrics", os.path.join(ROOT, "gcn-bmp_b200", "gcnbmp", "metrics.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _reference(y, t):
    k = y.shape[1]
    prc = []
    for c in range(k):
        p, r, _ = sk.precision_recall_curve(t[:, c], y[:, c], pos_label=1)
        prc.append(sk.auc(r, p))
    yr = np.round(y)
    return dict(roc_auc=sk.roc_auc_score(t, y, average="macro"), prc_auc=float(np.mean(prc)),
                f1=sk.f1_score(t, yr, average="macro", zero_division=0),
                accuracy=float(np.mean([sk.accuracy_score(t[:, c], yr[:, c]) for c in range(k)])),
                precision=float(np.mean([sk.precision_score(t[:, c], yr[:, c], average="binary", pos_label=1, zero_division=0) for c in range(k)])),
                recall=float(np.mean([sk.recall_score(t[:, c], yr[:, c], average="binary", pos_label=1, zero_division=0) for c in range(k)])))


@pytest.mark.parametrize("n,k,ties", [(500, 1, False), (2000, 7, False), (3000, 5, True), (64, 86, True)])
def test_metrics_match_sklearn(n, k, ties):
    M = _load()
    rng = np.random.default_rng(n + k)
    t = (rng.random((n, k)) < 0.3).astype(np.int64)
    t[0], t[1] = 1, 0                                   # both classes present in every column
    y = 1.0 / (1.0 + np.exp(-(rng.standard_normal((n, k)) + 1.5 * t - 0.7)))
    if ties:
        y = np.round(y, 1)                              # heavy ties: one curve point per distinct score
    ref = _reference(y, t)
    got = M.evaluate(torch.tensor(y), torch.tensor(t))
    for key, v in ref.items():
        assert abs(float(got[key]) - v) <= 1e-9, (key, float(got[key]), v)


def test_roc_auc_single_class_raises_like_sklearn():
    M = _load()
    with pytest.raises(ValueError):
        M.roc_auc(torch.rand(10, 2, dtype=torch.float64), torch.zeros(10, 2, dtype=torch.int64))


@pytest.mark.parametrize("n,k,ties", [(400, 1, False), (600, 5, True), (128, 86, False)])
def test_metrics_with_ignored_labels_match_sklearn_per_column(n, k, ties):
    """ignore_labels per label column: every column is scored over its own kept rows (the 86-class KAIST arrays)."""
    M = _load()
    rng = np.random.default_rng(7 * n + k)
    t = (rng.random((n, k)) < 0.3).astype(np.int64)
    t[0], t[1] = 1, 0
    y = 1.0 / (1.0 + np.exp(-(rng.standard_normal((n, k)) + 1.5 * t - 0.7)))
    if ties:
        y = np.round(y, 1)
    keep = rng.random((n, k)) > 0.2
    keep[:2] = True
    keep[np.argmax(y, axis=0), np.arange(k)] = False        # the top-scored entry of every column is ignored
    want = {key: [] for key in ("roc_auc", "prc_auc", "f1", "accuracy", "precision", "recall")}
    for c in range(k):
        r = _reference(y[keep[:, c], c:c + 1], t[keep[:, c], c:c + 1])
        if k > 1:      # multilabel evaluators: F1 of the positive class per column
            r["f1"] = sk.f1_score(t[keep[:, c], c], np.round(y[keep[:, c], c]), average="binary", zero_division=0)
        for key in want:
            want[key].append(r[key])
    got = M.evaluate(torch.tensor(y), torch.tensor(np.where(keep, t, 0)), torch.tensor(keep))
    for key, v in want.items():
        assert abs(float(got[key]) - float(np.mean(v))) <= 1e-9, (key, float(got[key]), float(np.mean(v)))
