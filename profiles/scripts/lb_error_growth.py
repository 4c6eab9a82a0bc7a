"""Diagnostic (GPU): how the large-batch step's deviation from the oracle grows over consecutive UNCORRECTED steps at the
real layer shapes (hidden 1024 x 1024, 8192 rows), per tensor, in both GEMM modes.  Prints max |got - ref| / max |ref|."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
import dqn_b200
from oracle import dqn_oracle as O
from test_gpu_large_batch import make

def rel(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    err = np.abs(got - ref)
    relerr = err / np.maximum(np.abs(ref), 1e-30)
    big = np.abs(ref) > 1e-3 * np.abs(ref).max()
    return "max|d|/max|ref| %.2e  p99.9 rel(|ref|>1e-3max) %.2e  frac>1e-5 %.4f" % (
        err.max() / np.abs(ref).max(), np.quantile(relerr[big], 0.999) if big.any() else 0.0, float(np.mean(relerr[big] > 1e-5)) if big.any() else 0.0)

for mode in ("fp32", "tc3xtf32"):
    for H, B in (((1024, 1024), 8192), ((256, 256), 512)):
        tr, ora = make(B=B, kind="adamw", gemm_mode=mode, N=20000, fill=20000, HID=H, seed=9)
        for step in range(4):
            ref = ora.step()
            tr.forward_backward(debug=True)
            got = tr.debug_read()
            print(f"== {mode} H={H[0]} B={B} step {step}: argmax agree {np.mean(got['max_actions'] == ref['max_actions']):.5f} loss rel {abs(got['loss'] - float(ref['loss'])) / abs(float(ref['loss'])):.2e}")
            for k in ("q", "targets"):
                print(f"   {k:8s} {rel(got[k], ref[k])}")
            for m in O.MODULES:
                print(f"   grad {m[-8:]:8s} w {rel(got['grads'][m]['w'], ref['grads'][m]['w'])}")
            tr.apply()
            p = tr.get_params()
            for m in O.MODULES:
                d = np.abs(p[m]["w"].astype(np.float64) - ora.params[m]["w"])
                print(f"   theta {m[-8:]:8s} w {rel(p[m]['w'], ora.params[m]['w'])}  n(|d|>1e-6) {int((d > 1e-6).sum())} max|d| {d.max():.2e}")
        tr.close()
