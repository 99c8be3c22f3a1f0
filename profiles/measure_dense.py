#!/usr/bin/env python
"""Timing of the hand-written dense-layer kernel (csrc/policy_dense.cu) at the shapes of one PPO minibatch step
(131072 windows = 655360 token rows): CUDA events, L2 flushed between launches; algorithmic bytes = A read + D written
(+ aux read for the ReLU-backward epilogue); the weight (<= 96 KB) is resident.  One JSON object per line."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uavenv_b200  # noqa: E402,F401
from target_allocation_ppo_transformer_b200 import _capi  # noqa: E402

PK = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
PEAK = json.load(open(PK))["hbm_gbs"] if os.path.isfile(PK) else 6650.0
L = _capi.load_policy()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for M, N, K, act in ((655360, 384, 128, 0), (655360, 128, 128, 0), (655360, 256, 128, 1), (655360, 128, 256, 0), (655360, 128, 384, 0),
                     (655360, 256, 128, 2), (131072, 256, 128, 0), (131072, 128, 128, 0), (131072, 64, 128, 1), (16384, 128, 128, 0)):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.1).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    aux = torch.relu(torch.randn(M, N, device="cuda")).to(torch.bfloat16) if act == 2 else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)

    def call():
        rc = L.uavpolicy_selftest_dense(C.c_void_p(a.data_ptr()), K, C.c_void_p(w.data_ptr()), C.c_void_p(bias.data_ptr()) if act != 2 else None,
                                        C.c_void_p(aux.data_ptr()) if aux is not None else None, N, C.c_void_p(out.data_ptr()), M, N, K, act, None)
        assert rc == 0
    for _ in range(3):
        call()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    torch.cuda.synchronize()
    for e0, e1 in ev:
        flush.fill_(1)
        e0.record(); call(); e1.record()
    torch.cuda.synchronize()
    ms = sorted(e0.elapsed_time(e1) for e0, e1 in ev)[len(ev) // 2]
    algo = M * K * 2 + M * N * 2 * (2 if act == 2 else 1)
    print(json.dumps({"kernel": "dense_kernel", "M": M, "N": N, "K": K, "act": ["identity", "relu", "drelu"][act], "us": 1e3 * ms,
                      "algo_GBps": algo / (ms * 1e-3) / 1e9, "frac_of_measured_hbm": algo / (ms * 1e-3) / 1e9 / PEAK,
                      "tflops": 2.0 * M * N * K / (ms * 1e-3) / 1e12}))
