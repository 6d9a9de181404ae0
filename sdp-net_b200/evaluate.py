"""Evaluation loop helpers right behind the forward: loss / top-1 accumulated ON
the device with one kernel per batch and no `.item()` until the end, then reduced across ranks.
Mirrors model_test.py:76-85 (CE + BCE-with-logits + accuracy) and `track_accuracy`
(training_utilities.py:50-88: SUM all_reduce of correct and total)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops


class EvalMeter:
    def __init__(self, device="cuda", label_smoothing: float = 0.0):
        self.acc = torch.zeros(5, dtype=torch.float64, device=device)   # ce_sum, bce_sum, correct, rows, bad labels
        self.label_smoothing = float(label_smoothing)
        self.classes = None

    def update(self, logits: torch.Tensor, labels: torch.Tensor) -> None:
        self.classes = logits.shape[1]
        ops.eval_metrics(logits, labels.to(torch.int64).contiguous(), self.acc, self.label_smoothing)

    def reset(self) -> None:
        self.acc.zero_()

    def synchronize(self, group=None) -> torch.Tensor:
        """Global sums over the ranks.  The local accumulator is left untouched (a clone is reduced), so `result()`
        can be read as often as wanted between `update()`s -- the running readout of model_test.py:84."""
        total = self.acc.clone()
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
        return total

    def result(self, group=None) -> dict:
        """One host sync: per-sample mean CE, per-element mean BCE (the reduction both reference losses use
        within a batch, here over the whole evaluation), top-1 accuracy.  Raises if any label was outside
        [0, classes) (what nn.CrossEntropyLoss does upstream; its ignore_index -100 is skipped)."""
        ce, bce, correct, rows, bad = (float(v) for v in self.synchronize(group).cpu())
        if bad > 0:
            raise ValueError(f"{int(bad)} label(s) outside [0, {self.classes}) reached EvalMeter.update")
        if rows == 0:
            return {"cross_entropy": 0.0, "bce_with_logits": 0.0, "accuracy": 0.0, "samples": 0}
        return {"cross_entropy": ce / rows, "bce_with_logits": bce / (rows * (self.classes or 1)),
                "accuracy": correct / rows, "samples": int(rows)}


@torch.no_grad()
def evaluate(model, batches, num_registers: int = 3, label_smoothing: float = 0.0, group=None) -> dict:
    """`for images, labels in batches` -> metrics; images/labels are moved to the model's device."""
    dev = model._device()
    meter = EvalMeter(dev, label_smoothing)
    for images, labels in batches:
        logits = model(images.to(dev, non_blocking=True), num_registers)
        meter.update(logits, labels.to(dev, non_blocking=True))
    return meter.result(group)
