"""Result meters returned by the rank / loss path — the boundary types of ``utils/metrics.py``.

``MetricResult`` is what ``Trainer.compute_one_batch`` and ``compute_metrics`` hand back
(utils/metrics.py:43-89): an ordered mapping loss, h1, h3, h10, h50, mrr, mr of count-weighted
running averages (``AccumulateMeter``, utils/metrics.py:4-40)."""
from __future__ import annotations

from collections import OrderedDict


class AccumulateMeter:
    def __init__(self, greater_is_better: bool = True, print_precision: int = 4):
        self.greater_is_better = greater_is_better
        self.print_precision = print_precision
        self.reset()

    def reset(self) -> None:
        self.avg, self.val, self.count = 0.0, 0.0, 0

    def update(self, val, n=1) -> None:
        total = self.count + n
        self.avg = (self.avg * self.count + val * n) / total
        self.val, self.count = val, total

    def __add__(self, other: "AccumulateMeter") -> "AccumulateMeter":
        if other.count > 0:
            self.update(other.avg, other.count)
        return self

    def avg_better_than(self, other: "AccumulateMeter") -> bool:
        return self.avg > other.avg if self.greater_is_better else self.avg < other.avg

    def avg_better_than_float(self, value: float) -> bool:
        return self.avg > value if self.greater_is_better else self.avg < value

    def __repr__(self) -> str:
        return f"{self.avg:.{self.print_precision}f}"


class MetricResult(OrderedDict):
    KEYS = ("loss", "h1", "h3", "h10", "h50", "mrr", "mr")

    def __init__(self):
        super().__init__()
        self["loss"] = AccumulateMeter(greater_is_better=False, print_precision=7)
        for k in self.KEYS[1:]:
            self[k] = AccumulateMeter()

    @property
    def metrics(self):
        return list(self.values())

    @property
    def averages(self) -> str:
        return "  ".join(f"{k}: {v}" for k, v in self.items())

    @property
    def averages_dict(self):
        return {k: v.avg for k, v in self.items()}

    def __add__(self, other: "MetricResult") -> "MetricResult":
        for mine, theirs in zip(self.values(), other.values()):
            if isinstance(theirs, AccumulateMeter):
                mine += theirs
        return self

    def reset(self) -> None:
        for m in self.values():
            m.reset()

    def __repr__(self) -> str:
        return "".join(f"{k}: {v.avg}\n" for k, v in self.items())
