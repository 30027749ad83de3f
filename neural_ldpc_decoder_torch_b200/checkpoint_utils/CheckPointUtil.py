"""CheckPointUtil — same file formats and method signatures as the reference
(/root/reference/src/checkpoint_utils/CheckPointUtil.py:8-159): a checkpoint is the dict
{'model_state_dict', 'optimizer_state_dict'?, 'epoch'?, <metrics flattened>, 'config'?} written with torch.save, and
`load` is a strict `load_state_dict`.  Because the decoder mirrors emit the reference's state_dict keys (dense structure
buffers included, synthesised on demand), files written here load in the reference and vice versa.

One deliberate difference: the per-tensor text dump (`save_weights(as_txt=True)`) writes the learned parameters and the
small structure matrices, but skips the two (E*Z)^2 lifting matrices (2 x 40 MB of zeros and ones per BG2 model; the
reference spends minutes in np.savetxt on them) unless `dump_lifting_matrices=True`."""
import os
from datetime import datetime
from typing import Any, Dict, Optional

import numpy as np
import torch


def _txt_name(param_name: str) -> str:
    return param_name.translate(str.maketrans("./", "__")) + ".txt"


def _write_matrix(path: str, t: torch.Tensor) -> None:
    """np.savetxt of one tensor; tensors of rank > 2 are flattened to [shape[0], -1] with the original shape in the header"""
    arr = t.detach().cpu().numpy()
    if arr.ndim <= 2:
        np.savetxt(path, arr)
    else:
        np.savetxt(path, arr.reshape(arr.shape[0], -1), header=f"Original shape: {arr.shape}\nReshaped to 2D for savetxt")


class CheckPointUtil:
    def __init__(self, checkpoint_dir: str = "checkpoints", dump_lifting_matrices: bool = False):
        self.checkpoint_dir, self.dump_lifting_matrices = checkpoint_dir, dump_lifting_matrices
        os.makedirs(self.checkpoint_dir, exist_ok=True)

    def _in_dir(self, name: str) -> str:
        return os.path.join(self.checkpoint_dir, name)

    def save(self, filepath: str, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None,
             epoch: Optional[int] = None, metrics: Optional[Dict[str, float]] = None,
             config: Optional[Dict[str, Any]] = None) -> str:
        # key order as the reference writes it: model, optimizer, epoch, the metrics flattened into the top level, config
        optional = (("optimizer_state_dict", None if optimizer is None else optimizer.state_dict()), ("epoch", epoch))
        data = {"model_state_dict": model.state_dict(), **{k: v for k, v in optional if v is not None}}
        data.update(metrics or {})
        if config is not None:
            data["config"] = config
        torch.save(data, self._in_dir(filepath))
        return self._in_dir(filepath)

    def save_weights(self, filepath: str, model: torch.nn.Module, as_txt: bool = False) -> str:
        weights_path = self._in_dir(filepath if filepath.endswith(".pth") else filepath + ".pth")
        state = model.state_dict()
        torch.save(state, weights_path)
        if not as_txt:
            return weights_path
        txt_dir = self._in_dir(filepath.replace(".pth", "") + "_weights_txt")
        os.makedirs(txt_dir, exist_ok=True)
        lines = [f"# Model weights saved at: {datetime.now():%Y-%m-%d %H:%M:%S}",
                 f"# Total parameters: {sum(p.numel() for p in model.parameters())}",
                 "# Format: Each parameter saved in separate .txt file", "-" * 80, "Parameter_Name, Shape, Filename"]
        for name, t in state.items():
            derived = name.rsplit(".", 1)[-1].startswith("Lift_Matrix") and not self.dump_lifting_matrices
            target = "<skipped: derived from the base graph>" if derived else _txt_name(name)
            if not derived:
                _write_matrix(os.path.join(txt_dir, target), t)
            lines.append(f"{name}, {list(t.shape)}, {target}")
        with open(os.path.join(txt_dir, "index.txt"), "w") as f:
            f.write("\n".join(lines) + "\n")
        return weights_path

    def load(self, filepath: str, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None,
             device: Optional[torch.device] = None) -> Dict[str, Any]:
        path = filepath if os.path.isabs(filepath) else self._in_dir(filepath)
        ckpt = torch.load(path, **({} if device is None else {"map_location": device}))
        model.load_state_dict(ckpt["model_state_dict"])        # strict, like the reference
        if optimizer is not None and "optimizer_state_dict" in ckpt:
            optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        return ckpt
