"""MetricsLogger — append-only text log of per-epoch metrics, same row format as the reference
(/root/reference/src/checkpoint_utils/MetricsLogger.py:5-70): `epoch, timestamp, metrics..., checkpoint file`, keys that
contain "ber" in scientific notation, a header block when epoch 0 is logged with a config."""
import os
from datetime import datetime
from typing import Any, Dict, Optional


class MetricsLogger:
    def __init__(self, log_dir: str = "checkpoints", filename: str = "training_metrics.txt"):
        os.makedirs(log_dir, exist_ok=True)
        self.log_dir, self.log_file = log_dir, os.path.join(log_dir, filename)
        self.best_ber = float("inf")

    @staticmethod
    def _cell(key: str, value: float) -> str:
        return format(value, ".6e" if "ber" in key.lower() else ".6f")

    def log(self, epoch: int, metrics: Dict[str, float], checkpoint_filename: str, config: Optional[Dict[str, Any]] = None):
        stamp = f"{datetime.now():%Y-%m-%d %H:%M:%S}"
        out = []
        starts_file = epoch == 0 and config is not None          # a fresh run rewrites the file and puts a header block first
        if starts_file:
            out += [f"# Training started: {stamp}",
                    "# Config: " + ", ".join(f"{k}={v}" for k, v in config.items()),
                    "# Columns: Epoch, Timestamp, " + ", ".join(metrics) + ", Checkpoint_File",
                    "-" * 120]
        out.append(", ".join([f"{epoch:4d}", stamp, *(self._cell(k, v) for k, v in metrics.items()), checkpoint_filename]))
        with open(self.log_file, "w" if starts_file else "a") as f:
            f.write("\n".join(out) + "\n")

    def is_best(self, ber: float) -> bool:
        better = ber < self.best_ber
        if better:
            self.best_ber = ber
        return better
