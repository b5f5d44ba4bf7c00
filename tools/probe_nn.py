"""phase timestamps of CTA (0,0) of one linear_kernel launch (build with -DB200NN_PROBE=1 -> lib/probe_nn.so)"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B200NN_LIB"] = os.path.join(ROOT, "vid2player3d_b200", "lib", "probe_nn.so")
import torch
from vid2player3d_b200 import nn
dev = "cuda:0"
names = ["start", "setup done", "first TMA issued", "producer done", "first full", "mma issued all", "tmem_full seen", "epilogue done", "end"]
for (M, K, N) in ((8192, 64, 64), (8192, 512, 75), (8192, 1024, 1024), (128, 64, 64)):
    a = nn.padded_bf16(M, K, dev); a.normal_()
    out = nn.padded_bf16(M, N, dev)
    lin = nn.Linear(a, torch.randn(N, K), torch.randn(N), out, M, act="relu")
    for _ in range(3):
        lin.run()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 16)()
    nn.lib().b200nn_probe_read(buf)
    t = list(buf)[:9]
    nc = (nn.padded_rows(M) // 128) * ((N + 127) // 128)
    cb = (C.c_ulonglong * (8 * nc))()
    nn.lib().b200nn_probe_read_ctas(cb, C.c_int(nc))
    v = list(cb)
    st = v[0::8]
    t00 = min(st)
    ends = [[v[8 * c + k] for c in range(nc)] for k in range(1, 6)]
    print(f"   {nc} CTAs: starts span {(max(st) - t00) / 1e3:.2f} us; epilogue warp ends (max over CTAs) " +
          " ".join(f"{(max(e) - t00) / 1e3:.2f}" for e in ends[:4]) + f" us; dealloc done max {(max(ends[4]) - t00) / 1e3:.2f} us")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): lin.run()
    e1.record(); torch.cuda.synchronize()
    print(f"   back-to-back launches: {e0.elapsed_time(e1) / 20 * 1e3:.2f} us per launch")
    print(f"M={M} K={K} N={N}: " + ", ".join(f"{n} +{(x - t[0]) / 1e3:.2f}us" for n, x in zip(names, t)))
