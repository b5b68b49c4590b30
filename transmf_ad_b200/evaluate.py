"""Validation / test step and metrics on the B200 path -- the reference's ``val_step`` and metric glue
(kfold_train_adversarial.py:144-187; utils/utils.py:44-51) without ignite.

``val_step`` is the reference function verbatim in behaviour: ``model.eval()``, ``torch.no_grad()``, forward, CE loss; the
model's eval forward runs the inference path (BatchNorm folded into the conv operands, no statistics, nothing saved).
``EvalAccumulator`` replaces the ignite metrics: arg-max labels / positive-class softmax probability / confusion counts come
from one ``tmf_eval_head`` launch per batch and stay on the device until ``compute()``; ROC-AUC is the rank statistic
(Mann-Whitney U with average ranks for ties), which is what sklearn's ``roc_auc_score`` (behind ignite's ``ROC_AUC``) computes.
"""
from __future__ import annotations

import torch

from . import _lib as L


def eval_head(logits, labels=None, counts=None):
    """-> (pred int64 [B], prob_last fp32 [B]); accumulates {TN, FP, FN, TP} into ``counts`` (4 x int64, device) if given."""
    logits = logits.detach().float().contiguous()
    B, C = logits.shape
    pred = torch.empty(B, dtype=torch.int64, device=logits.device)
    prob = torch.empty(B, dtype=torch.float32, device=logits.device)
    lab = None if labels is None else labels.detach().to(torch.int64).contiguous()
    L.call("tmf_eval_head", L.ptr(logits), L.ptr(lab), L.ptr(pred), L.ptr(prob), L.ptr(counts), B, C)
    return pred, prob


def val_step(model, batch, device, criterion=None):
    """reference kfold_train_adversarial.py:144-161 (``batch`` is a dict with 'MRI', 'PET', 'label'; MRI-only models take
    'MRI').  Returns the reference's output dict plus 'pred' and 'prob' (metric inputs)."""
    criterion = criterion or torch.nn.CrossEntropyLoss()
    model.eval()
    with torch.no_grad():
        mri = batch["MRI"].to(device, non_blocking=True)
        label = batch["label"].to(device, non_blocking=True)
        if "PET" in batch:
            out = model(mri, batch["PET"].to(device, non_blocking=True))
        else:
            out = model(mri)
        logits = out[0] if isinstance(out, tuple) else out
        loss = criterion(logits, label)
        pred, prob = eval_head(logits)
    return {"label": label, "logits": logits, "loss": loss, "pred": pred, "prob": prob}


def cal_confusion_metrics(c_matrix):
    """reference utils/utils.py:44-51 on c_matrix[label][pred]."""
    TP, FN, FP, TN = c_matrix[1][1], c_matrix[1][0], c_matrix[0][1], c_matrix[0][0]
    precision = TP / (TP + FP)
    recall = TP / (TP + FN)
    f1 = 2 * precision * recall / (precision + recall)
    sen = TP / (TP + FN)
    spe = TN / (FP + TN)
    return sen, spe, f1


def roc_auc(prob, label):
    """Area under the ROC curve = P(score_pos > score_neg) + 0.5 P(tie): average-rank Mann-Whitney statistic."""
    prob, label = prob.double().flatten(), label.flatten()
    n_pos, n_neg = int((label == 1).sum()), int((label == 0).sum())
    if n_pos == 0 or n_neg == 0:
        return float("nan")
    order = torch.argsort(prob)
    sp = prob[order]
    ranks = torch.arange(1, len(sp) + 1, dtype=torch.float64, device=prob.device)
    # average ranks over ties
    uniq, inv, cnt = torch.unique_consecutive(sp, return_inverse=True, return_counts=True)
    sums = torch.zeros(len(uniq), dtype=torch.float64, device=prob.device).index_add_(0, inv, ranks)
    avg = (sums / cnt.double())[inv]
    r = torch.empty_like(avg)
    r[order] = avg
    u = float(r[label == 1].sum()) - n_pos * (n_pos + 1) / 2.0
    return u / (n_pos * n_neg)


class EvalAccumulator:
    """Accuracy / confusion matrix / ROC-AUC / mean loss over an evaluation run (reference :178-187)."""

    def __init__(self, device):
        self.counts = torch.zeros(4, dtype=torch.int64, device=device)
        self.probs, self.labels, self.losses, self.n = [], [], [], 0

    def update(self, out):
        eval_head(out["logits"], out["label"], self.counts)      # confusion counts on the device (pred / prob recomputed: 1 launch)
        self.probs.append(out["prob"])
        self.labels.append(out["label"])
        self.losses.append(out["loss"].detach() * out["label"].numel())
        self.n += out["label"].numel()

    def compute(self):
        tn, fp, fn, tp = [int(v) for v in self.counts.tolist()]          # the only host synchronisation
        c = [[tn, fp], [fn, tp]]
        res = {"accuracy": (tn + tp) / max(self.n, 1), "confusion": c,
               "auc": roc_auc(torch.cat(self.probs), torch.cat(self.labels)),
               "loss": float(torch.stack(self.losses).sum()) / max(self.n, 1)}
        try:
            res["sen"], res["spe"], res["f1"] = cal_confusion_metrics(c)
        except ZeroDivisionError:
            res["sen"] = res["spe"] = res["f1"] = float("nan")
        return res
