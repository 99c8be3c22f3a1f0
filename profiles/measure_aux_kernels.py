#!/usr/bin/env python
"""Timing of the auxiliary kernels (score matrix, GAE + normalisation, reset) with CUDA events; writes one JSON
object per line.  Run on the GPU box:  python profiles/measure_aux_kernels.py > gpurun_out/aux.jsonl"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uavenv_b200 as ub  # noqa: E402

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.isfile(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, warm=3, iters=10, flush=None):
    for _ in range(warm):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    torch.cuda.synchronize()
    for a, b in ev:
        if flush is not None:
            flush.fill_(1)
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / iters


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for B, N, M in ((4096, 30, 10), (16384, 64, 64), (1024, 256, 256)):
        env = ub.UAVEnvBatched(B, config=ub.Config(NUM_UAVS=N, NUM_TARGETS=M), seed=1)
        ms_reset = timeit(lambda: env.reset(full_reset=True), 1, 3)
        for dt, nm in ((torch.float32, "f32"), (torch.float64, "f64")):
            pf = torch.empty(B, N, M, dtype=dt, device="cuda"); pd = torch.empty_like(pf)
            import ctypes as C
            fn = env._lib.uavenv_score_matrix if dt == torch.float32 else env._lib.uavenv_score_matrix_f64
            call = lambda: fn(env._h, C.c_void_p(pf.data_ptr()), C.c_void_p(pd.data_ptr()), env._stream())
            ms = timeit(call, 2, 5, flush)
            pairs = B * N * M
            print(json.dumps({"kernel": "score_matrix_kernel<%s>" % nm, "envs": B, "N": N, "M": M, "ms": ms,
                              "pairs_per_sec": pairs / (ms * 1e-3),
                              "out_GBps": pairs * 2 * pf.element_size() / (ms * 1e-3) / 1e9}))
        print(json.dumps({"kernel": "reset_kernel(full: 2 scenes/env)", "envs": B, "N": N, "M": M, "ms": ms_reset,
                          "scenes_per_sec": 2 * B / (ms_reset * 1e-3)}))
        env.close()
    for T, B in ((128, 16384), (256, 65536), (2048, 64), (300, 1)):
        r = torch.randn(T, B, device="cuda"); v = torch.randn(T, B, device="cuda")
        d = torch.rand(T, B, device="cuda") < 0.02
        lv = torch.randn(B, device="cuda")
        ms = timeit(lambda: ub.compute_gae(r, v, d, lv, 0.998, 0.95, normalize=True), 3, 10, flush)
        # algorithmic bytes: read r,v (4+4) + done (1) once, write ret, adv (4+4), normalise adv in place (4+4)
        algo = T * B * 25
        print(json.dumps({"kernel": "gae_kernel+normalize_kernel", "T": T, "B": B, "ms": ms,
                          "elements_per_sec": T * B / (ms * 1e-3), "algo_GBps": algo / (ms * 1e-3) / 1e9,
                          "frac_of_measured_hbm": algo / (ms * 1e-3) / 1e9 / PEAK}))


if __name__ == "__main__":
    main()
