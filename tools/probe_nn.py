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
    cb = (C.c_ulonglong * (2 * nc))()
    nn.lib().b200nn_probe_read_ctas(cb, C.c_int(nc))
    st, en = list(cb)[0::2], list(cb)[1::2]
    t00 = min(st)
    print(f"   {nc} CTAs: starts span {(max(st) - t00) / 1e3:.2f} us, epilogue ends: min {(min(en) - t00) / 1e3:.2f} median {(sorted(en)[nc // 2] - t00) / 1e3:.2f} max {(max(en) - t00) / 1e3:.2f} us")
    print(f"M={M} K={K} N={N}: " + ", ".join(f"{n} +{(x - t[0]) / 1e3:.2f}us" for n, x in zip(names, t)))
