"""Per-tile cost of short-K GEMM streams by epilogue kind (diagnostic, run on a B200)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lirec_b200 import _ext, ops
_ext.require_device()
dev = "cuda"

def rnd(r, c, s=0.05):
    return (torch.randn(r, c, device=dev) * s).to(torch.bfloat16)

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def run(M, N, K, passes, kind):
    a, b = rnd(M, K), rnd(N, K)
    ps = [(_ext.operand(a), 0, 0, _ext.operand(b), 0, 0, K)] * passes
    kw = {}
    if kind == "f32":
        out = torch.empty(M, N, device=dev)
    elif kind == "split":
        out = torch.empty(M, 2 * N, device=dev, dtype=torch.bfloat16); kw = dict(out_kind=ops.OUT_SPLIT, out_lo_off=N)
    elif kind == "tanh_drop_split":
        out = torch.empty(M, 2 * N, device=dev, dtype=torch.bfloat16)
        kw = dict(out_kind=ops.OUT_SPLIT, out_lo_off=N, act=ops.ACT_TANH, post=ops.POST_DROPOUT,
                  drop=ops.dropout_desc(0.3, 1, 2, 0), bias=torch.randn(N, device=dev))
    elif kind == "split_t":
        MP = (M + 63) // 64 * 64
        out = torch.empty(2 * N, MP, device=dev, dtype=torch.bfloat16); kw = dict(out_kind=ops.OUT_SPLIT_T, out_ld_m=MP, out_lo_off=N)
    elif kind == "relu_f32":
        out = torch.empty(M, N, device=dev); kw = dict(act=ops.ACT_RELU, bias=torch.randn(N, device=dev))
    g = ops.gemm_problem(M, N, ps, out=out, **kw)
    ms = timeit(lambda: ops.gemm_grouped([g]))
    tiles = ((M + 255) // 256) * ((N + 255) // 256)
    rounds = (tiles + 73) // 74
    print("M=%-6d N=%-5d K=%-5d passes=%d %-16s: %.3f ms %7.1f TFLOP/s  tiles=%d rounds=%d  %.2f us/round (MMA floor %.2f us)" % (
        M, N, K, passes, kind, ms, 2.0 * M * N * K * passes / ms / 1e9, tiles, rounds, 1e3 * ms / rounds,
        passes * (K / 64) * 4 * 128 / 1.9e3))
    sys.stdout.flush()

if __name__ == "__main__":
    for kind in ["f32", "relu_f32", "split", "split_t", "tanh_drop_split"]:
        run(8448 * 8, 512, 512, 2, kind)
    for kind in ["f32", "split", "split_t", "tanh_drop_split"]:
        run(8448 * 8, 512, 128, 2, kind)
    for kind in ["f32", "split", "tanh_drop_split"]:
        run(8448 * 8, 512, 2048, 2, kind)
