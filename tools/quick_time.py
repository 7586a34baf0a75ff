"""Quick per-stage timing of the commit path at the BASELINE shape (device-resident inputs, CUDA events)."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import leanmultisig_b200 as L
from leanmultisig_b200._lib import lib, check

n_vars = int(sys.argv[1]) if len(sys.argv) > 1 else 28
k, r = 7, 1
live = 1 << (n_vars - 1)
ctx = L.Context(0, 24)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
torch.cuda.set_stream(stream)
P = 0x7F000001
g = torch.Generator(device="cuda").manual_seed(0)
ev = (torch.randint(0, P, (live,), dtype=torch.int64, device="cuda", generator=g)).to(torch.int32)
h = 1 << (n_vars + r - k)
cols = 64
cw = torch.empty((h, cols), dtype=torch.int32, device="cuda")
layers = torch.empty((2 * h - 1, 8), dtype=torch.int32, device="cuda")
pt = torch.randint(0, P, (n_vars, 5), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
out = torch.empty(8, dtype=torch.int32, device="cuda")

def timeit(name, fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{name:28s} min {min(ts):8.3f} ms  med {sorted(ts)[len(ts)//2]:8.3f} ms", flush=True)
    return min(ts)

l = lib()
t_ntt = timeit("reorder_and_dft", lambda: check(l.lm_dev_reorder_and_dft(ctx.handle, ev.data_ptr(), n_vars, 1, k, r, cols, cw.data_ptr())))
t_mk = timeit("merkle_tree (64 of 128 live)", lambda: check(l.lm_dev_merkle_tree(ctx.handle, cw.data_ptr(), h, cols, 128, 64, layers.data_ptr())))
t_ev = timeit("mle_eval 2^%d" % n_vars, lambda: check(l.lm_dev_mle_eval(ctx.handle, ev.data_ptr(), n_vars, 1, live, pt.data_ptr(), out.data_ptr())))
st = torch.randint(0, P, (1 << 22, 16), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
t_p = timeit("poseidon1 compress 2^22", lambda: check(l.lm_dev_poseidon1(ctx.handle, st.data_ptr(), 1 << 22, 1)))
n_perm = h * 8 + h - 1
print(f"commit (ntt+merkle) {t_ntt + t_mk:.3f} ms -> {2**n_vars / (t_ntt + t_mk) / 1e6:.2f} Gelem/s; "
      f"merkle {n_perm / t_mk / 1e6:.2f} Gperm/s; raw perm {(1<<22) / t_p / 1e6:.2f} Gperm/s")
