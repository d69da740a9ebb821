"""Per-epoch evaluation driver: the reference registers one BatchEvaluator subclass per metric and split
(train_binary.py:566-634: Accuracy / ROC-AUC / PRC-AUC / F1 on the train and the validation iterator), each of which runs
the WHOLE split through the predictor again, moves the logits to the host and calls scikit-learn
(training/extensions/batch_evaluator.py:49-101).  Here a split is scored ONCE: logits stay on the device, sigmoid as
batch_evaluator.py:51-53,84, and all metrics come from gcnbmp.metrics on the device-resident scores."""
import torch

from . import metrics as M
from .train import PairTrainer


class BatchEvaluator(object):
    """`evaluate(...)` -> {'<name>/accuracy': .., '<name>/roc_auc': .., '<name>/prc_auc': .., '<name>/f1': .., ..} like the
    observation keys the reference's reporter emits ('train_acc/main/accuracy', 'val_roc/main/roc_auc', ...)."""

    def __init__(self, model, name="val", batch=4096, ignore_labels=-1, raise_value_error=True,
                 which=("accuracy", "roc_auc", "prc_auc", "f1", "precision", "recall")):
        self.trainer = model if isinstance(model, PairTrainer) else PairTrainer(model, chunk=batch, optimizer=False)
        self.name, self.ignore, self.raise_value_error, self.which = name, ignore_labels, raise_value_error, tuple(which)

    def _report(self, logits, labels):
        dev = logits.device
        t = (labels if isinstance(labels, torch.Tensor) else torch.as_tensor(labels)).to(dev)
        t = t.reshape(logits.shape)
        y = torch.sigmoid(logits.double())
        mask = None
        if self.ignore is not None and bool((t == self.ignore).any()):
            # the evaluators drop ignored entries before calling scikit-learn; with several label columns (the 86-class KAIST
            # set) every column keeps its own non-ignored rows (gcnbmp.metrics `mask`)
            mask = t != self.ignore
            t = torch.where(mask, t, torch.zeros_like(t))
        fns = dict(accuracy=M.accuracy, roc_auc=M.roc_auc, prc_auc=M.prc_auc, f1=M.f1, precision=M.precision, recall=M.recall)
        out = {}
        for k in self.which:
            try:
                out["%s/%s" % (self.name, k)] = float(fns[k](y, t, mask))
            except ValueError:
                if self.raise_value_error:
                    raise
                out["%s/%s" % (self.name, k)] = float("nan")      # raise_value_error=False: warn-and-skip in the reference
        return out

    def evaluate(self, atoms_1, adjs_1, atoms_2, adjs_2, labels):
        """Arrays in the reference layout (what concat_mols hands the predictor)."""
        return self._report(self.trainer.predict(atoms_1, adjs_1, atoms_2, adjs_2), labels)

    def evaluate_indexed(self, table_atoms, table_adjs, idx_1, idx_2, labels):
        """Index pairs into a device-resident drug table: every drug of the split is encoded once."""
        return self._report(self.trainer.predict_indexed(table_atoms, table_adjs, idx_1, idx_2), labels)
