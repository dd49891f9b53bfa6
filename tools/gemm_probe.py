"""GPU probe of the grouped tcgen05 GEMM: runs layout / pass / epilogue cases and prints error
statistics for each (diagnostic tool; the asserting versions live in tests/test_gemm_gpu.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lirec_b200 import _ext, ops

torch.manual_seed(0)
dev = "cuda"
_ext.require_device()


def rnd(r, c, scale=1.0):
    return (torch.randn(r, c, device=dev) * scale).to(torch.bfloat16)


def report(name, got, ref):
    got = got.double(); ref = ref.double()
    err = (got - ref).abs().max().item()
    den = ref.abs().max().item() + 1e-30
    bad = ((got - ref).abs() > 1e-2 * den).sum().item()
    print("%-46s max_abs_err %.3e  rel(maxnorm) %.3e  bad %d/%d  %s" % (
        name, err, err / den, bad, got.numel(), "OK" if err / den < 1e-4 else "FAIL"))
    if err / den >= 1e-4:
        idx = ((got - ref).abs() > 1e-2 * den).nonzero()
        if idx.numel():
            print("   first bad idx:", idx[:6].tolist(), " rows bad:", idx[:, 0].unique()[:12].tolist(),
                  " cols bad:", idx[:, 1].unique()[:12].tolist())
    sys.stdout.flush()


def case_kmajor(M, N, K, name):
    a, b = rnd(M, K), rnd(N, K)
    out = torch.full((M, N), float("nan"), device=dev)
    g = ops.gemm_problem(M, N, [(_ext.operand(a), 0, 0, _ext.operand(b), 0, 0, K)], out=out)
    ops.gemm_grouped([g]); torch.cuda.synchronize()
    report(name, out, a.double() @ b.double().t())


def case_mn_a(M, N, K, name):   # A stored [K, M] (MN-major), B K-major
    a, b = rnd(K, M), rnd(N, K)
    out = torch.full((M, N), float("nan"), device=dev)
    g = ops.gemm_problem(M, N, [(_ext.operand(a), 0, 0, _ext.operand(b), 0, 0, K)], a_mn_major=True, out=out)
    ops.gemm_grouped([g]); torch.cuda.synchronize()
    report(name, out, a.double().t() @ b.double().t())


def case_mn_b(M, N, K, name):   # A K-major, B stored [K, N] (MN-major)  (dgrad)
    KP = (K + 63) // 64 * 64        # A is zero-padded to a k-block multiple; B rows past K are TMA OOB zeros
    a, b = rnd(M, KP), rnd(K, N)
    a[:, K:] = 0
    out = torch.full((M, N), float("nan"), device=dev)
    g = ops.gemm_problem(M, N, [(_ext.operand(a), 0, 0, _ext.operand(b), 0, 0, KP)], b_mn_major=True, out=out)
    ops.gemm_grouped([g]); torch.cuda.synchronize()
    report(name, out, a[:, :K].double() @ b.double())


def case_mn_ab(M, N, K, name):  # both MN-major (wgrad): out = A^T B, A [K,M], B [K,N]
    MP = (M + 63) // 64 * 64        # dY is stored with its class dim padded (as the model does)
    a, b = rnd(K, MP), rnd(K, N)
    out = torch.full((M, N), float("nan"), device=dev)
    g = ops.gemm_problem(M, N, [(_ext.operand(a), 0, 0, _ext.operand(b), 0, 0, K)], a_mn_major=True,
                         b_mn_major=True, out=out)
    ops.gemm_grouped([g]); torch.cuda.synchronize()
    report(name, out, a[:, :M].double().t() @ b.double())


def case_split(M, N, K, name):
    x = torch.randn(M, K, device=dev)
    hi = x.to(torch.bfloat16); lo = (x - hi.float()).to(torch.bfloat16)
    a = torch.cat([hi, lo], 1).contiguous()
    b = rnd(N, K)
    out = torch.full((M, N), float("nan"), device=dev)
    oa = _ext.operand(a)
    g = ops.gemm_problem(M, N, [(oa, 0, 0, _ext.operand(b), 0, 0, K), (oa, 0, K, _ext.operand(b), 0, 0, K)], out=out)
    ops.gemm_grouped([g]); torch.cuda.synchronize()
    report(name + " (vs fp32 x)", out, x.double() @ b.double().t())


def case_epilogue(M, N, K, name):
    a, b = rnd(M, K, 0.2), rnd(N, K, 0.2)
    bias = torch.randn(N, device=dev)
    out = torch.zeros(M, 2 * N, device=dev, dtype=torch.bfloat16)
    g = ops.gemm_problem(M, N, [(_ext.operand(a), 0, 0, _ext.operand(b), 0, 0, K)], alpha=0.5, bias=bias,
                         act=ops.ACT_TANH, out=out, out_kind=ops.OUT_SPLIT, out_lo_off=N)
    ops.gemm_grouped([g]); torch.cuda.synchronize()
    ref = torch.tanh(0.5 * (a.double() @ b.double().t()) + bias.double())
    got = out[:, :N].double() + out[:, N:].double()
    report(name, got, ref)


def case_grouped(name):
    probs, checks = [], []
    for (M, N, K) in [(700, 512, 768), (130, 256, 512), (64, 101, 3072), (1000, 15, 1536)]:  # N=101/15: B rows OOB
        a, b = rnd(M, K), rnd(N, K)
        out = torch.full((M, N), float("nan"), device=dev)
        probs.append(ops.gemm_problem(M, N, [(_ext.operand(a), 0, 0, _ext.operand(b), 0, 0, K)], out=out))
        checks.append((out, a, b, (M, N, K)))
    ops.gemm_grouped(probs); torch.cuda.synchronize()
    for out, a, b, shp in checks:
        report(name + " " + str(shp), out, a.double() @ b.double().t())


def bench(M, N, K, iters=20):
    a, b = rnd(M, K), rnd(N, K)
    out = torch.empty(M, N, device=dev)
    g = ops.gemm_problem(M, N, [(_ext.operand(a), 0, 0, _ext.operand(b), 0, 0, K)], out=out)
    for _ in range(3):
        ops.gemm_grouped([g])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.gemm_grouped([g])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("bench M=%d N=%d K=%d: %.3f ms  %.1f TFLOP/s" % (M, N, K, ms, 2.0 * M * N * K / ms / 1e9))
    t0 = time.time()
    ref = a @ b.t()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        ref = a @ b.t()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("   cuBLAS bf16 same shape: %.3f ms  %.1f TFLOP/s" % (ms, 2.0 * M * N * K / ms / 1e9))


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    case_kmajor(128, 128, 64, "K-major 128x128x64 (1 tile, 1 k-block)")
    case_kmajor(128, 128, 256, "K-major 128x128x256")
    case_kmajor(300, 200, 320, "K-major ragged 300x200x320")
    case_kmajor(4096, 1024, 1024, "K-major 4096x1024x1024 (multi-tile persistent)")
    case_mn_b(128, 128, 64, "B MN-major 128x128x64")
    case_mn_b(300, 200, 101, "B MN-major ragged 300x200x101 (dgrad-like)")
    case_mn_a(128, 128, 64, "A MN-major 128x128x64")
    case_mn_a(200, 304, 136, "A MN-major ragged")
    case_mn_ab(128, 128, 64, "A,B MN-major 128x128x64")
    case_mn_ab(101, 3072, 777, "A,B MN-major wgrad-like 101x3072x777")
    case_mn_ab(512, 768, 1000, "A,B MN-major wgrad-like 512x768x1000")
    case_split(256, 256, 512, "hi/lo split A, 2 passes")
    case_epilogue(200, 256, 128, "epilogue alpha+bias+tanh -> split out")
    case_grouped("grouped")
    bench(8192, 3072, 6144)
    bench(8192, 512, 2048)
