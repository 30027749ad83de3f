#!/usr/bin/env python
"""kernel-level breakdown of one training step (torch profiler): python tools/prof_train_kernels.py [B]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.argv = ["prof_train.py"] + sys.argv[1:]
exec(compile(open(os.path.join(HERE, "prof_train.py")).read(), "prof_train.py", "exec"))   # set up + two warm steps  # noqa: S102
x, y = gen(B)   # noqa: F821
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    outs = model(x, target_iter=list(range(T)))   # noqa: F821
    loss = crit(outs, y, coeff_param=list(range(T)))   # noqa: F821
    loss.backward()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
print(prof.key_averages(group_by_input_shape=True, group_by_stack_n=6).table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=40, max_shapes_column_width=60, max_src_column_width=110))
