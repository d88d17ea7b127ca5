"""In-kernel timeline of hallo_b200_gemm (diagnostic): runs a GEMM shape on the -DHB_GEMM_TRACE build of the library
(tools/libhb_gemm_trace.so, built by `python tools/gemm_trace.py --build`), reads back the per-CTA event records and
prints where the producer, the MMA issuer and the two epilogue groups spend their clocks per tile.
    python tools/gemm_trace.py --build        (here: cross-compiles the traced library)
    python tools/gemm_trace.py                (on the GPU box)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRACE_LIB = os.path.join(ROOT, "tools", "libhb_gemm_trace.so")

if "--build" in sys.argv:
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
           "-DHB_GEMM_TRACE", "-o", TRACE_LIB, os.path.join(ROOT, "hallo_b200", "csrc", "lib.cu")]
    print(" ".join(cmd))
    subprocess.check_call(cmd, cwd=os.path.join(ROOT, "hallo_b200", "csrc"))
    sys.exit(0)

os.environ["HALLO_B200_LIB"] = TRACE_LIB
sys.path.insert(0, ROOT)
import ctypes as C  # noqa: E402

import torch  # noqa: E402

from hallo_b200 import lib, ops  # noqa: E402

SLOTS = 256
NAMES = {1: "start", 2: "prod.tile_begin", 3: "prod.tile_issued", 10: "mma.tile_begin", 11: "mma.acc_free", 12: "mma.first_stage",
         13: "mma.last_stage", 20: "epi.tile_begin", 21: "epi.acc_ready", 22: "epi.res_wait", 23: "epi.res_ready",
         24: "epi.panel_written", 25: "epi.group_synced", 26: "epi.store_issued+prev_read", 27: "epi.unit_loaded",
         28: "epi.unit_biased", 14: "mma.stage_wait_clocks"}


def run(M, N, K, geglu=False, residual=True, label="", use_bias=True, verbose=True):
    dev = "cuda"
    a = torch.randn(M, K, device=dev, dtype=torch.float16)
    w = torch.randn(N, K, device=dev, dtype=torch.float16) * 0.05
    bias = torch.randn(N, device=dev, dtype=torch.float16)
    n_out = N // 2 if geglu else N
    out = torch.empty(M, n_out, device=dev, dtype=torch.float16)
    res = torch.randn(M, n_out, device=dev, dtype=torch.float16) if residual else None
    fn = lambda: ops.gemm(a, w, out, bias=bias if use_bias else None, residual=res, geglu=geglu)
    h = lib.load()
    h.hallo_b200_gemm_trace_buffer.argtypes = [C.c_void_p]
    h.hallo_b200_gemm_trace_buffer(None)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    # flush L2 so that the traced launch sees what a step sees (operands from HBM)
    junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    junk.fill_(1)
    buf = torch.zeros(148 * 4 * SLOTS * 2, dtype=torch.int64, device=dev)
    h.hallo_b200_gemm_trace_buffer(C.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    h.hallo_b200_gemm_trace_buffer(None)
    rec = buf.cpu().view(148, 4, SLOTS, 2)
    print(f"\n=== gemm M{M} N{N} K{K} geglu={geglu} residual={residual} bias={use_bias} {label}: {e0.elapsed_time(e1) * 1e3:.1f} us (cold L2)")
    import collections
    agg = collections.defaultdict(list)
    for cta in range(148):
        t0 = None
        for r in range(4):
            evs = [(int(x[0]) >> 32, int(x[0]) & 0xffffffff, int(x[1])) for x in rec[cta, r] if int(x[1]) != 0]
            if r == 0 and evs:
                t0 = evs[0][2]
            for e, i, t in evs:
                if e == 14:
                    agg[(r, "mma.stage_wait_clocks", "(sum over the tile's k-blocks)")].append(i)
            evs = [x for x in evs if x[0] != 14]
            for (e, i, t), (e2, i2, t2) in zip(evs, evs[1:]):
                agg[(r, NAMES.get(e, e), NAMES.get(e2, e2))].append(t2 - t)
            if verbose and cta in (0,) and evs and t0 is not None:
                line = " ".join(f"{NAMES.get(e, e).split('.')[-1]}[{i}]@{t - t0}" for e, i, t in evs[:26])
                print(f"  cta {cta} rec {r}: {line}")
    print("  mean clocks between consecutive events (count):")
    for (r, a_, b_), v in sorted(agg.items()):
        print(f"    rec{r} {a_:28s} -> {b_:28s} {sum(v) / len(v):9.0f}  x{len(v)}  (max {max(v)})")


if __name__ == "__main__":
    run(131072, 320, 320, label="L0 to_out / proj (x55 per step)")
    run(131072, 320, 320, residual=False, label="no residual", verbose=False)
    run(131072, 320, 320, use_bias=False, label="no bias", verbose=False)
    run(131072, 320, 320, residual=False, use_bias=False, label="plain", verbose=False)
    run(131072, 960, 320, residual=False, use_bias=False, label="L0 QKV")
    run(131072, 2560, 320, geglu=True, residual=False, label="L0 GEGLU", verbose=False)
    run(32768, 640, 640, label="L1 to_out", verbose=False)
