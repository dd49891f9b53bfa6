"""Grouped tcgen05 GEMM vs fp64 matmul on identically rounded (bf16) operands — through the C ABI."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 2e-5   # fp32 accumulation of exact bf16 products: only summation-order error remains


def _rnd(r, c, scale=1.0):
    return (torch.randn(r, c, device="cuda") * scale).to(torch.bfloat16)


def _rel(got, ref):
    got, ref = got.double(), ref.double()
    return float((got - ref).abs().max() / (ref.abs().max() + 1e-30))


@pytest.fixture(autouse=True)
def _seed():
    torch.manual_seed(0)
    from lirec_b200 import _ext
    _ext.require_device()


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 200, 320), (1, 8, 64), (4096, 1024, 1024), (77, 3072, 3072)])
def test_k_major(M, N, K):
    from lirec_b200 import _ext, ops
    a, b = _rnd(M, K), _rnd(N, K)
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm_grouped([ops.gemm_problem(M, N, [(_ext.operand(a), 0, 0, _ext.operand(b), 0, 0, K)], out=out)])
    assert _rel(out, a.double() @ b.double().t()) < TOL


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 200, 101), (5000, 1536, 15)])
def test_b_mn_major_dgrad_layout(M, N, K):
    """dX = dY @ W with W[out=K, in=N] read in place (rows past K are TMA out-of-bounds zeros)."""
    from lirec_b200 import _ext, ops
    KP = (K + 63) // 64 * 64
    a, b = _rnd(M, KP), _rnd(K, N)
    a[:, K:] = 0
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm_grouped([ops.gemm_problem(M, N, [(_ext.operand(a), 0, 0, _ext.operand(b), 0, 0, KP)],
                                       b_mn_major=True, out=out)])
    assert _rel(out, a[:, :K].double() @ b.double()) < TOL


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (101, 3072, 777), (512, 768, 1000), (15, 1536, 333)])
def test_both_mn_major_wgrad_layout(M, N, K):
    """dW = dY^T X with dY [K, M] and X [K, N] read in place; the reduction tail is zero-filled."""
    from lirec_b200 import _ext, ops
    MP = (M + 63) // 64 * 64
    a, b = _rnd(K, MP), _rnd(K, N)
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm_grouped([ops.gemm_problem(M, N, [(_ext.operand(a), 0, 0, _ext.operand(b), 0, 0, K)],
                                       a_mn_major=True, b_mn_major=True, out=out)])
    assert _rel(out, a[:, :M].double().t() @ b.double()) < TOL


def test_split_passes_recover_fp32_operand():
    """hi/lo split A (2 passes over the same B) reproduces the fp32 product to ~2^-17."""
    from lirec_b200 import _ext, ops
    M, N, K = 256, 256, 512
    x = torch.randn(M, K, device="cuda")
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    a = torch.cat([hi, lo], 1).contiguous()
    b = _rnd(N, K)
    out = torch.empty(M, N, device="cuda")
    oa, ob = _ext.operand(a), _ext.operand(b)
    ops.gemm_grouped([ops.gemm_problem(M, N, [(oa, 0, 0, ob, 0, 0, K), (oa, 0, K, ob, 0, 0, K)], out=out)])
    assert _rel(out, x.double() @ b.double().t()) < 2e-5
    single = torch.empty(M, N, device="cuda")
    ops.gemm_grouped([ops.gemm_problem(M, N, [(oa, 0, 0, ob, 0, 0, K)], out=single)])
    assert _rel(single, x.double() @ b.double().t()) > 1e-4    # one bf16 pass is NOT parity-grade


def test_epilogue_bias_tanh_dropout_split_and_derivatives():
    from lirec_b200 import _ext, ops
    from oracle import dropout as od
    import numpy as np
    M, N, K, p, seed = 200, 256, 128, 0.3, 99
    a, b = _rnd(M, K, 0.2), _rnd(N, K, 0.2)
    bias = torch.randn(N, device="cuda")
    flag = (torch.arange(M, device="cuda") % 3 != 0).int()
    f2 = torch.zeros(M, 2 * N, device="cuda", dtype=torch.bfloat16)
    drop = ops.dropout_desc(p, seed, 4, 6)
    ops.gemm_grouped([ops.gemm_problem(M, N, [(_ext.operand(a), 0, 0, _ext.operand(b), 0, 0, K)], alpha=0.5,
                                       bias=bias, row_flag=flag, act=ops.ACT_TANH, post=ops.POST_DROPOUT, drop=drop,
                                       out=f2, out_kind=ops.OUT_SPLIT, out_lo_off=N)])
    keep = torch.from_numpy(od.keep_mask(seed, 4, np.arange(M), 6 + np.arange(N), p)).cuda().double()
    z = 0.5 * (a.double() @ b.double().t()) + bias.double() * flag.double().view(-1, 1)
    ref = torch.tanh(z) * keep / (1 - p)
    got = f2[:, :N].double() + f2[:, N:].double()
    assert _rel(got, ref) < 1e-5
    # DTANH: upstream gradient through dropout(tanh(.)) recomputed from the stored split feature
    g, w = _rnd(M, 64), _rnd(N, 64)
    dz = torch.zeros(M, 2 * N, device="cuda", dtype=torch.bfloat16)
    ops.gemm_grouped([ops.gemm_problem(M, N, [(_ext.operand(g), 0, 0, _ext.operand(w), 0, 0, 64)], post=ops.POST_DTANH,
                                       drop=drop, aux=f2, aux_lo_off=N, out=dz, out_kind=ops.OUT_SPLIT, out_lo_off=N)])
    ref = (g.double() @ w.double().t()) * keep / (1 - p) * (1 - torch.tanh(z) ** 2)
    assert _rel(dz[:, :N].double() + dz[:, N:].double(), ref) < 1e-4
    # DRELU: gradient through dropout(relu(.)) from the stored feature's sign
    r2 = torch.zeros(M, 2 * N, device="cuda", dtype=torch.bfloat16)
    ops.gemm_grouped([ops.gemm_problem(M, N, [(_ext.operand(a), 0, 0, _ext.operand(b), 0, 0, K)], bias=bias,
                                       act=ops.ACT_RELU, post=ops.POST_DROPOUT, drop=drop, out=r2,
                                       out_kind=ops.OUT_SPLIT, out_lo_off=N)])
    d = torch.zeros(M, 2 * N, device="cuda", dtype=torch.bfloat16)
    ops.gemm_grouped([ops.gemm_problem(M, N, [(_ext.operand(g), 0, 0, _ext.operand(w), 0, 0, 64)], post=ops.POST_DRELU,
                                       post_scale=1 / (1 - p), aux=r2, aux_lo_off=N, out=d, out_kind=ops.OUT_SPLIT,
                                       out_lo_off=N)])
    pre = a.double() @ b.double().t() + bias.double()
    ref = (g.double() @ w.double().t()) * keep / (1 - p) * (pre > 0)
    assert _rel(d[:, :N].double() + d[:, N:].double(), ref) < 1e-5


def test_grouped_launch_strided_and_transposed_outputs():
    from lirec_b200 import _ext, ops
    probs, checks, keep_alive = [], [], []          # problems hold raw pointers: keep operands alive
    for (M, N, K) in [(700, 512, 768), (130, 256, 512), (64, 101, 3072), (1000, 15, 1536)]:
        a, b = _rnd(M, K), _rnd(N, K)
        keep_alive += [a, b]
        out = torch.full((M, N), float("nan"), device="cuda")
        probs.append(ops.gemm_problem(M, N, [(_ext.operand(a), 0, 0, _ext.operand(b), 0, 0, K)], out=out))
        checks.append((out, a.double() @ b.double().t()))
    a, b = _rnd(200, 128), _rnd(96, 128)
    out_t = torch.full((96, 200), float("nan"), device="cuda")     # transposed store: out[n, m]
    probs.append(ops.gemm_problem(200, 96, [(_ext.operand(a), 0, 0, _ext.operand(b), 0, 0, 128)], out=out_t,
                                  out_ld_m=1, out_ld_n=200))
    checks.append((out_t.t(), a.double() @ b.double().t()))
    acc = torch.ones(200, 96, device="cuda")
    probs.append(ops.gemm_problem(200, 96, [(_ext.operand(a), 0, 0, _ext.operand(b), 0, 0, 128)], out=acc,
                                  accumulate=True))
    checks.append((acc, 1 + a.double() @ b.double().t()))
    ops.gemm_grouped(probs)
    for got, ref in checks:
        assert _rel(got, ref) < TOL


def test_argument_errors_are_reported():
    from lirec_b200 import _ext, ops
    a, b = _rnd(64, 100), _rnd(64, 100)          # ld = 100 elements: not 16-byte aligned rows
    out = torch.empty(64, 64, device="cuda")
    with pytest.raises(RuntimeError, match="16-byte"):
        ops.gemm_grouped([ops.gemm_problem(64, 64, [(_ext.operand(a), 0, 0, _ext.operand(b), 0, 0, 100)], out=out)])


def test_transposed_gradient_forms():
    """The three operand forms of backward (all with a K-major B, the CTA-pair kernel's fast path):
    dgrad  dX = dY W      : A = dY^T [feat, rows] read MN-major, B = W^T [in, out], transposed split store;
    wgrad  dW = dY^T X    : computed as X^T dY with A = X natural (MN-major), B = dY^T, transposed fp32 store;
    bgrad  db = dY^T 1    : A = dY^T K-major, B = a single row of ones (rows past it are TMA zero fill)."""
    from lirec_b200 import _ext, ops
    rows, out_f, in_f = 1000, 200, 328
    pitch = 1024
    dy = torch.randn(rows, out_f, device="cuda")
    dyT = torch.full((2 * 256, pitch), float("nan"), device="cuda", dtype=torch.bfloat16)   # hi rows 0.., lo rows 256..
    dyT[:256, :rows] = 0
    dyT[256:, :rows] = 0
    hi = dy.to(torch.bfloat16)
    dyT[:out_f, :rows] = hi.t()
    dyT[256:256 + out_f, :rows] = (dy - hi.float()).to(torch.bfloat16).t()
    w = _rnd(out_f, in_f)
    wT = torch.zeros(in_f, 256, device="cuda", dtype=torch.bfloat16)
    wT[:, :out_f] = w.t()
    o_dyT = _ext.operand(dyT[:, :rows])                       # view: cols = rows, ld = pitch
    # dgrad, transposed split output [2 * in_p, pitch]
    in_p = 384
    dxT = torch.zeros(2 * in_p, pitch, device="cuda", dtype=torch.bfloat16)
    ops.gemm_grouped([ops.gemm_problem(rows, in_f, [(o_dyT, 0, 0, _ext.operand(wT), 0, 0, 256),
                                                    (o_dyT, 0, 256, _ext.operand(wT), 0, 0, 256)],
                                       a_mn_major=True, out=dxT, out_kind=ops.OUT_SPLIT_T, out_ld_m=pitch,
                                       out_lo_off=in_p)])
    got = (dxT[:in_f, :rows].double() + dxT[in_p:in_p + in_f, :rows].double()).t()
    assert _rel(got, dy.double() @ w.double()) < 2e-5
    assert (dxT[:, rows:] == 0).all() and (dxT[in_f:in_p] == 0).all()         # nothing outside the valid block
    # wgrad, transposed fp32 store into dW[out_f, in_f]
    x = _rnd(rows, in_f)
    dW = torch.full((out_f, in_f), float("nan"), device="cuda")
    ops.gemm_grouped([ops.gemm_problem(in_f, out_f, [(_ext.operand(x), 0, 0, o_dyT, 0, 0, rows),
                                                     (_ext.operand(x), 0, 0, o_dyT, 256, 0, rows)],
                                       a_mn_major=True, out=dW, out_ld_m=1, out_ld_n=in_f)])
    assert _rel(dW, dy.double().t() @ x.double()) < 2e-5
    # bgrad against one row of ones
    ones = torch.ones(1, pitch, device="cuda", dtype=torch.bfloat16)
    db = torch.full((out_f,), float("nan"), device="cuda")
    ops.gemm_grouped([ops.gemm_problem(out_f, 1, [(o_dyT, 0, 0, _ext.operand(ones[:, :rows]), 0, 0, rows),
                                                  (o_dyT, 256, 0, _ext.operand(ones[:, :rows]), 0, 0, rows)],
                                       out=db, out_ld_m=1)])
    assert _rel(db, dy.double().sum(0)) < 2e-5


@pytest.mark.parametrize("rows,in_f,pitch", [(997, 328, 1024), (1000, 96, 1004), (61, 40, 64), (300, 256, 304)])
def test_transposed_split_store_edges(rows, in_f, pitch):
    """The shared-memory staged transposed hi/lo store (16-byte pieces of eight consecutive rows) at its
    edges: a row count that is not a multiple of 8, partial column chunks, a pitch that forbids vector
    stores — bit-identical values to the split of the fp32 product and nothing written outside."""
    from lirec_b200 import _ext, ops
    out_f = 128
    torch.manual_seed(rows)
    dy = torch.randn(rows, out_f, device="cuda")
    kp = (rows + 63) // 64 * 64
    dyT = torch.zeros(out_f, kp, device="cuda", dtype=torch.bfloat16)
    dyT[:, :rows] = dy.to(torch.bfloat16).t()
    w = _rnd(out_f, in_f)
    wT = w.t().contiguous()
    in_p = (in_f + 7) // 8 * 8 + 8
    dxT = torch.full((2 * in_p, pitch), 7.0, device="cuda", dtype=torch.bfloat16)
    ops.gemm_grouped([ops.gemm_problem(rows, in_f, [(_ext.operand(dyT[:, :rows]), 0, 0, _ext.operand(wT), 0, 0, out_f)],
                                       a_mn_major=True, out=dxT, out_kind=ops.OUT_SPLIT_T, out_ld_m=pitch,
                                       out_lo_off=in_p)])
    ref = dy.to(torch.bfloat16).double() @ w.double()                     # [rows, in_f]
    got = (dxT[:in_f, :rows].double() + dxT[in_p:in_p + in_f, :rows].double()).t()
    assert _rel(got, ref) < 2e-5
    assert (dxT[:, rows:] == 7).all() and (dxT[in_f:in_p] == 7).all() and (dxT[in_p + in_f:] == 7).all()


@pytest.mark.parametrize("M,N,K", [(20000, 512, 768), (530, 512, 2048), (64, 64, 64)])      # pair kernel, single-CTA
def test_sign_mask_side_output(M, N, K):
    """LIREC_POST_SIGN_MASK: next to relu(a b^T + bias) the epilogue writes [v > 0] as one bit per element
    (word n >> 5 of row m, bit n & 31) — the ReLU gate backward's scatter-reduce reads instead of the activations."""
    from lirec_b200 import _ext, ops
    a, b = _rnd(M, K), _rnd(N, K)
    bias = torch.randn(N, device="cuda") * 3
    out = torch.full((M, N), float("nan"), device="cuda")
    mask = torch.full((M + 3, N // 32 + 2), -1, dtype=torch.int32, device="cuda")     # wider pitch, guard rows
    ops.gemm_grouped([ops.gemm_problem(M, N, [(_ext.operand(a), 0, 0, _ext.operand(b), 0, 0, K)], bias=bias,
                                       act=ops.ACT_RELU, post=ops.POST_SIGN_MASK, aux=mask, out=out)])
    torch.cuda.synchronize()
    ref = torch.relu(a.double() @ b.double().t() + bias.double())
    assert _rel(out, ref) < TOL
    bits = (out > 0).view(M, N // 32, 32).to(torch.int64)
    words = (bits << torch.arange(32, device="cuda")).sum(-1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)
    assert torch.equal(mask[:M, :N // 32], words)
    assert bool((mask[M:] == -1).all()) and bool((mask[:, N // 32:] == -1).all())
