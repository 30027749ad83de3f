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


class CheckPointUtil:
    def __init__(self, checkpoint_dir: str = "checkpoints", dump_lifting_matrices: bool = False):
        self.checkpoint_dir = checkpoint_dir
        self.dump_lifting_matrices = dump_lifting_matrices
        os.makedirs(checkpoint_dir, exist_ok=True)

    def save(self, filepath: str, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None,
             epoch: Optional[int] = None, metrics: Optional[Dict[str, float]] = None,
             config: Optional[Dict[str, Any]] = None) -> str:
        path = os.path.join(self.checkpoint_dir, filepath)
        data = {'model_state_dict': model.state_dict()}
        if optimizer is not None:
            data['optimizer_state_dict'] = optimizer.state_dict()
        if epoch is not None:
            data['epoch'] = epoch
        if metrics is not None:
            data.update(metrics)
        if config is not None:
            data['config'] = config
        torch.save(data, path)
        return path

    def save_weights(self, filepath: str, model: torch.nn.Module, as_txt: bool = False) -> str:
        pth = filepath if filepath.endswith('.pth') else filepath + '.pth'
        weights_path = os.path.join(self.checkpoint_dir, pth)
        state = model.state_dict()
        torch.save(state, weights_path)
        if as_txt:
            base = filepath.replace('.pth', '')
            txt_dir = os.path.join(self.checkpoint_dir, f"{base}_weights_txt")
            os.makedirs(txt_dir, exist_ok=True)
            index = os.path.join(txt_dir, "index.txt")
            with open(index, 'w') as f:
                f.write(f"# Model weights saved at: {datetime.now().strftime('%Y-%m-%d %H:%M:%S')}\n")
                f.write(f"# Total parameters: {sum(p.numel() for p in model.parameters())}\n")
                f.write("# Format: Each parameter saved in separate .txt file\n")
                f.write("-" * 80 + "\n")
                f.write("Parameter_Name, Shape, Filename\n")
                for name, t in state.items():
                    if name.split('.')[-1].startswith("Lift_Matrix") and not self.dump_lifting_matrices:
                        f.write(f"{name}, {list(t.shape)}, <skipped: derived from the base graph>\n")
                        continue
                    fname = name.replace('.', '_').replace('/', '_') + ".txt"
                    arr = t.cpu().numpy()
                    if arr.ndim > 2:
                        shape = arr.shape
                        np.savetxt(os.path.join(txt_dir, fname), arr.reshape(shape[0], -1),
                                   header=f"Original shape: {shape}\nReshaped to 2D for savetxt")
                    else:
                        np.savetxt(os.path.join(txt_dir, fname), arr)
                    f.write(f"{name}, {list(t.shape)}, {fname}\n")
        return weights_path

    def load(self, filepath: str, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None,
             device: Optional[torch.device] = None) -> Dict[str, Any]:
        path = filepath if os.path.isabs(filepath) else os.path.join(self.checkpoint_dir, filepath)
        checkpoint = torch.load(path, map_location=device) if device is not None else torch.load(path)
        model.load_state_dict(checkpoint['model_state_dict'])
        if optimizer is not None and 'optimizer_state_dict' in checkpoint:
            optimizer.load_state_dict(checkpoint['optimizer_state_dict'])
        return checkpoint
