"""tcgen05 / TMEM / TMA GEMM (MMI_IMPL_TC) against fp64 matmul of the same bf16 inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def dev():
    from segmminterest_b200 import _lib
    assert torch.cuda.is_available()
    assert _lib.load().mmi_has_tc() == 1, "tcgen05 path unavailable on this device/build"
    return torch.device("cuda:0")


@pytest.mark.parametrize("M,N,K", [(1000, 512, 512), (4096, 3072, 512), (300, 64, 48), (5000, 512, 640), (128, 256, 64),
                                    (777, 1024, 3072), (2048, 2048, 512)])
def test_tc_gemm_nt(dev, M, N, K):
    from segmminterest_b200 import ops
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device=dev).bfloat16()
    W = torch.randn(N, K, device=dev).bfloat16()
    C = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
    ops.gemm(ops.GEMM_NT, ops.IMPL_TC, A, K, W, K, C, N, M, N, K)
    ref = A.double() @ W.double().T
    assert _rel(C, ref) < 4e-3
    Cf = torch.full((M, N), float("nan"), device=dev, dtype=torch.float32)
    ops.gemm(ops.GEMM_NT, ops.IMPL_TC, A, K, W, K, Cf, N, M, N, K)
    assert _rel(Cf, ref) < 1e-5  # fp32 accumulation order over K


@pytest.mark.parametrize("Mg,Ng,Kg", [(512, 512, 5000), (3072, 512, 4096), (64, 48, 300), (512, 640, 40960), (2048, 512, 999),
                                       (128, 128, 64)])
def test_tc_gemm_tn_weight_gradient(dev, Mg, Ng, Kg):
    """dW[Mg,Ng] += dY[Kg,Mg]^T X[Kg,Ng]: MN-major operands straight from row-major activations, split-K."""
    from segmminterest_b200 import ops
    torch.manual_seed(Mg + Ng + Kg)
    dY = torch.randn(Kg, Mg, device=dev).bfloat16()
    X = torch.randn(Kg, Ng, device=dev).bfloat16()
    C = torch.ones(Mg, Ng, device=dev, dtype=torch.float32)
    ops.gemm(ops.GEMM_TN, ops.IMPL_TC, dY, Mg, X, Ng, C, Ng, Mg, Ng, Kg, accumulate=True, split_k=0)
    ref = dY.double().T @ X.double()
    assert _rel(C - 1, ref) < 1e-5
    C2 = torch.zeros(Mg, Ng, device=dev, dtype=torch.float32)
    ops.gemm(ops.GEMM_TN, ops.IMPL_TC, dY, Mg, X, Ng, C2, Ng, Mg, Ng, Kg, accumulate=True, split_k=1)
    assert _rel(C2, ref) < 3e-4  # one long fp32 accumulation chain in TMEM (no split): error grows with K


def test_tc_gemm_column_slices_and_epilogue(dev):
    from segmminterest_b200 import ops
    torch.manual_seed(7)
    M, N, K, L = 1200, 512, 512, 40
    big = torch.randn(M, 3 * K, device=dev).bfloat16()       # A is a column slice of a wider tensor (lda = 3K)
    A = big[:, K:2 * K]
    W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev)
    pe = torch.randn(L, N, device=dev)
    C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    import ctypes
    Aview = big.view(-1)[K:]  # same storage, offset K elements
    ops.gemm(ops.GEMM_NT, ops.IMPL_TC, Aview, 3 * K, W, K, C, N, M, N, K, bias=b, act=ops.ACT_GELU, preact=pre, add=pe, add_mod=L, ld_add=N)
    z = A.double() @ W.double().T + b.double()
    ref = torch.nn.functional.gelu(z) + pe.double().repeat(M // L, 1)
    assert _rel(pre, z) < 4e-3 and _rel(C, ref) < 4e-3
    Z = torch.randn(M, N, device=dev).bfloat16()
    R = torch.randn(M, N, device=dev).bfloat16()
    ops.gemm(ops.GEMM_NT, ops.IMPL_TC, Aview, 3 * K, W, K, C, N, M, N, K, mul_gelu_grad=Z, add=R, add_mod=M, ld_add=N)
    zz = Z.double().requires_grad_(True)
    torch.nn.functional.gelu(zz).sum().backward()
    ref = (A.double() @ W.double().T) * zz.grad + R.double()
    assert _rel(C, ref) < 5e-3


def test_tc_gemm_matches_simt_bit_pattern_scale(dev):
    """Same bf16 inputs through both implementations: fp32 outputs agree to accumulation-order noise."""
    from segmminterest_b200 import ops
    torch.manual_seed(9)
    M, N, K = 2048, 512, 512
    A = torch.randn(M, K, device=dev).bfloat16()
    W = torch.randn(N, K, device=dev).bfloat16()
    c1 = torch.empty(M, N, device=dev)
    c2 = torch.empty(M, N, device=dev)
    ops.gemm(ops.GEMM_NT, ops.IMPL_TC, A, K, W, K, c1, N, M, N, K)
    ops.gemm(ops.GEMM_NT, ops.IMPL_SIMT, A, K, W, K, c2, N, M, N, K)
    assert _rel(c1, c2) < 2e-6


def _gelu_grad(z):
    zz = z.double().clone().requires_grad_(True)
    torch.nn.functional.gelu(zz).sum().backward()
    return zz.grad


@pytest.mark.parametrize("M,N,K", [(1000, 512, 512), (4133, 256, 512), (300, 3072, 512), (257, 128, 64), (2048, 512, 3072), (999, 192, 128)])
def test_tc_gemm_tma_epilogue_variants(dev, M, N, K):
    """The TMA-store / TMA-prefetch epilogue (bf16 out): plain, bias, residual add, GELU with saved z or saved
    gelu'(z), multiply by gelu'(z) or by a saved derivative -- including ragged M and N tails."""
    from segmminterest_b200 import ops
    torch.manual_seed(M * 7 + N + K)
    A = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(N, K, device=dev) * 0.06).bfloat16()
    b = torch.randn(N, device=dev)
    R = torch.randn(M, N, device=dev).bfloat16()
    Z = torch.randn(M, N, device=dev).bfloat16()
    base = A.double() @ W.double().T

    def run(**kw):
        C = torch.full((M + 3, N), 7.0, device=dev, dtype=torch.bfloat16)   # 3 guard rows: the TMA store must clip at M
        ops.gemm(ops.GEMM_NT, ops.IMPL_TC, A, K, W, K, C, N, M, N, K, **kw)
        torch.cuda.synchronize()
        assert bool((C[M:] == 7.0).all()), "store wrote past row M"
        return C[:M]

    tol = 5e-3
    assert _rel(run(), base) < tol
    assert _rel(run(bias=b), base + b.double()) < tol
    assert _rel(run(bias=b, add=R, add_mod=M, ld_add=N), base + b.double() + R.double()) < tol
    assert _rel(run(mul_gelu_grad=Z), base * _gelu_grad(Z)) < tol
    assert _rel(run(mul_gelu_grad=Z, mul_is_grad=True), base * Z.double()) < tol
    assert _rel(run(bias=b, act=ops.ACT_GELU), torch.nn.functional.gelu(base + b.double())) < tol
    for save_grad in (False, True):
        pre = torch.full((M + 3, N), 7.0, device=dev, dtype=torch.bfloat16)
        out = run(bias=b, act=ops.ACT_GELU, preact=pre, save_act_grad=save_grad)
        z = base + b.double()
        assert _rel(out, torch.nn.functional.gelu(z)) < tol
        assert _rel(pre[:M], _gelu_grad(z) if save_grad else z) < tol
        assert bool((pre[M:] == 7.0).all())


def test_tc_gemm_tma_epilogue_strided_outputs(dev):
    """Output / residual that are column slices of wider tensors (ldc != N), as the fused-projection buffers are."""
    from segmminterest_b200 import ops
    torch.manual_seed(11)
    M, N, K = 1500, 512, 512
    A = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    wide = torch.zeros(M, 3 * N, device=dev, dtype=torch.bfloat16)
    Rw = torch.randn(M, 2 * N, device=dev).bfloat16()
    Cview = wide.view(-1)[N:]
    Rview = Rw.view(-1)[N:]
    ops.gemm(ops.GEMM_NT, ops.IMPL_TC, A, K, W, K, Cview, 3 * N, M, N, K, add=Rview, add_mod=M, ld_add=2 * N)
    ref = A.double() @ W.double().T + Rw[:, N:].double()
    assert _rel(wide[:, N:2 * N], ref) < 5e-3
    assert float(wide[:, :N].abs().max()) == 0.0 and float(wide[:, 2 * N:].abs().max()) == 0.0
