"""Does an HBM-bound elementwise kernel overlap a tensor-bound conv kernel when they are launched on two streams?
(Decides whether the stream lanes of the sliding-window driver can hide the norm / gather / head passes.)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from brainseg_b200 import _lib as L
from brainseg_b200 import packing as P

dev = torch.device("cuda:0")
lib = L.lib()
N, D, cin, cout = 4, 128, 64, 64
x = torch.randn(N, D, D, D, cin, device=dev).to(torch.float16)
out = torch.empty(N, D, D, D, cout, device=dev, dtype=torch.float16)
w = torch.randn(cout, cin, 3, 3, 3, device=dev) / (27 * cin) ** 0.5
wp = P.pack_conv3_weight(w, cin, torch.float16)
bp = P.pad_bias(None, cout).to(dev)
plan = L.ConvPlan(kind=0, stride=1, N=N, D=D, H=D, W=D, cin=cin, in_ptr=x.data_ptr(), in_ctot=cin, cout=cout,
                  out_ptr=out.data_ptr(), out_ctot=cout, out_coff=0, weights=wp.data_ptr(), bias=bp.data_ptr(), act=1,
                  slope=0.01, stats=None, out_f16=1, in_f16=1, use_khshift=-1, max_ctas=0)
y = torch.randn(N, D, D, D, cout, device=dev).to(torch.float16)
ss = torch.ones(N, cout, 2, device=dev)
vox = D * D * D


def apply(stream):
    L.check(lib.bsg_norm_apply_lrelu(C.c_void_p(y.data_ptr()), vox, N, cout, cout, 0, C.c_void_p(ss.data_ptr()), 0.01, 1, 1,
                                     L.stream_ptr(stream)))


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
t_conv = timed(lambda: plan.run(None))
t_apply = timed(lambda: apply(None))


def both():
    ev = torch.cuda.Event()
    ev.record()
    s1.wait_event(ev)
    s2.wait_event(ev)
    plan.run(s1)
    apply(s2)
    a, b = torch.cuda.Event(), torch.cuda.Event()
    a.record(s1)
    b.record(s2)
    torch.cuda.current_stream().wait_event(a)
    torch.cuda.current_stream().wait_event(b)


def both_apply_first():
    ev = torch.cuda.Event()
    ev.record()
    s1.wait_event(ev)
    s2.wait_event(ev)
    apply(s2)
    plan.run(s1)
    a, b = torch.cuda.Event(), torch.cuda.Event()
    a.record(s1)
    b.record(s2)
    torch.cuda.current_stream().wait_event(a)
    torch.cuda.current_stream().wait_event(b)


t_both = timed(both)
t_both2 = timed(both_apply_first)
print(f"conv alone {t_conv:.3f} ms, norm apply alone {t_apply:.3f} ms, serial sum {t_conv + t_apply:.3f} ms")
print(f"two streams (conv launched first) {t_both:.3f} ms, (apply launched first) {t_both2:.3f} ms")
