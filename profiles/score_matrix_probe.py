#!/usr/bin/env python
"""Two launches of score_matrix_kernel<float> at 16384 envs x 64 x 64 (the ncu target of profiles/r2_capture.sh)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uavenv_b200 as ub  # noqa: E402

env = ub.UAVEnvBatched(16384, config=ub.Config(NUM_UAVS=64, NUM_TARGETS=64), seed=1)
env.reset()
pf = torch.empty(16384, 64, 64, device="cuda"); pd = torch.empty_like(pf)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(3):
    e0.record()
    env._lib.uavenv_score_matrix(env._h, C.c_void_p(pf.data_ptr()), C.c_void_p(pd.data_ptr()), env._stream())
    e1.record()
    torch.cuda.synchronize()
print("score_matrix_kernel<float>: %.3f ms, %.3g pairs/s" % (e0.elapsed_time(e1), 16384 * 4096 / (e0.elapsed_time(e1) * 1e-3)))
env.close()
