"""Instruction-throughput micro-benchmarks on the SM (per-SM per-clock rates) -- evidence for softmax design."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hallo_b200 import lib  # noqa: E402

NAMES = {0: "ex2.approx.ftz.f32", 1: "ex2.approx.ftz.f16x2", 2: "ex2.approx.ftz.bf16x2", 3: "fma.rn.f32 (3-reg)",
         4: "max.f32", 5: "cvt.rn.f16x2.f32", 6: "add.f32"}
h = lib.load()
scratch = torch.zeros(16, device="cuda")
iters = 4096
sms = torch.cuda.get_device_properties(0).multi_processor_count
clk = 1.965e9
for mode, name in NAMES.items():
    for _ in range(2):
        n = h.hallo_b200_ubench_exp(C.c_int(mode), C.c_int(iters), C.c_void_p(scratch.data_ptr()), lib.current_stream())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = h.hallo_b200_ubench_exp(C.c_int(mode), C.c_int(iters), C.c_void_p(scratch.data_ptr()), lib.current_stream())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ops = n * iters * 8
    per_sm_clk = ops / (ms * 1e-3) / sms / clk
    print(f"{name:26s} {ms:8.3f} ms  {per_sm_clk:7.1f} thread-instr / clk / SM (at 1965 MHz)", flush=True)
