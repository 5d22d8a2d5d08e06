#!/usr/bin/env python
"""Times agnn_gemm_pair (CTA pairs, 256 x 256 tiles) against the single-CTA grouped kernel on the large products of a
step (fp16 hi / lo operand pairs, fp32 parity mode).  CUDA events, L2 flushed between runs, median of 20."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from analysisgnn_b200 import _lib, linalg  # noqa: E402

DEV = "cuda:0"
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=DEV)


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    ms = []
    for _ in range(n):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def main():
    lib = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    for m, n, k, b_mn in [(50000, 256, 2560, False), (50000, 2560, 256, True), (50000, 512, 256, True),
                          (50000, 256, 512, False), (50000, 256, 256, False), (200000, 256, 2560, False)]:
        x = linalg.split_f16(torch.randn(m, k, device=DEV))
        w = linalg.split_f16(torch.randn(k, n, device=DEV) * 0.1 if b_mn else torch.randn(n, k, device=DEV) * 0.1)
        out = torch.empty(m, n, device=DEV)
        layout = _lib.MN_MAJOR if b_mn else _lib.K_MAJOR

        def pair():
            _lib.check(lib.agnn_gemm_pair(layout, m, n, k, x.hi.data_ptr(), x.lo.data_ptr(), x.hi.stride(0), x.amax.data_ptr(),
                                          w.hi.data_ptr(), w.lo.data_ptr(), w.hi.stride(0), w.amax.data_ptr(),
                                          out.data_ptr(), out.stride(0), None, 0, None, st))

        def single():
            linalg._gemm(x, _lib.K_MAJOR, w, layout, m, n, k, None, 0, out)

        linalg.GEMM_PAIR = False
        t1, t2 = timeit(single), timeit(pair)
        fl = 2.0 * m * n * k
        print(f"M={m} N={n} K={k} B={'MN' if b_mn else 'K'}: single {t1 * 1e3:7.1f} us ({fl / t1 / 1e9:6.1f} TF/s)  "
              f"pair {t2 * 1e3:7.1f} us ({fl / t2 / 1e9:6.1f} TF/s)  x{t1 / t2:.2f}", flush=True)


if __name__ == "__main__":
    main()
