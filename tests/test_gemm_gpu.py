"""GPU: the strided fp32 GEMM of the PointWiseMLP path against a float64 torch product (floating point:
tolerance 1e-5 relative to the result's max magnitude, written here)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(1000, 144, 75), (4096, 72, 144), (33, 5, 7), (257, 130, 40)])
def test_sgemm_row_major_times_transposed_weights(cuda, M, N, K):
    from closerlook3d_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    lda = K + 5
    a = torch.randn(M, lda, generator=g).to(cuda)
    w = torch.randn(N, K, generator=g).to(cuda)
    out = ops.sgemm(a, lda, 1, w, 1, K, M, N, K)  # out[m][n] = sum_k a[m][k] w[n][k]
    ref = (a[:, :K].double() @ w.double().t()).float()
    assert float((out[:, :N] - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))


def test_sgemm_plain_and_ldc(cuda):
    from closerlook3d_b200 import ops
    g = torch.Generator().manual_seed(3)
    M, N, K, ldc = 777, 72, 144, 80
    a = torch.randn(M, K, generator=g).to(cuda)
    b = torch.randn(K, N + 3, generator=g).to(cuda)  # only the first N columns are used
    out = ops.sgemm(a, K, 1, b, N + 3, 1, M, N, K, ldc=ldc)
    ref = (a.double() @ b[:, :N].double()).float()
    assert out.shape == (M, ldc)
    assert float((out[:, :N] - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("splitk", [1, 7, 64])
def test_sgemm_transposed_a_splitk(cuda, splitk):
    from closerlook3d_b200 import ops
    g = torch.Generator().manual_seed(11)
    P, M, N = 5000, 144, 75  # out (M x N) = ga^T (M x P) @ f (P x N): the weight-gradient shape
    ga = torch.randn(P, M, generator=g).to(cuda)
    f = torch.randn(P, N + 5, generator=g).to(cuda)
    out = ops.sgemm(ga, 1, M, f, N + 5, 1, M, N, P, splitk=splitk)
    ref = (ga.double().t() @ f[:, :N].double()).float()
    assert float((out[:, :N] - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))
