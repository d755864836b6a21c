"""Kernel-level parity on the B200, through the C ABI: the tcgen05 GEMM (every epilogue and tile
width) and the 197-token attention against plain fp32/fp64 torch references of the same op."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bf16(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).bfloat16().float().numpy()


GEMM_CASES = [
    # M, N, K, bn, epi
    (1, 768, 768, 32, 0), (7, 768, 768, 32, 0), (64, 2304, 768, 32, 0), (128, 256, 64, 32, 0), (129, 256, 64, 64, 0),
    (200, 768, 768, 128, 0), (333, 768, 256, 256, 0), (788, 768, 256, 192, 0), (1576, 2304, 768, 256, 0), (1576, 3072, 768, 192, 1),
    (64, 3072, 768, 32, 1), (64, 768, 3072, 32, 2), (1000, 768, 3072, 192, 2), (1000, 768, 3072, 256, 2), (8, 768, 768, 32, 5),
    (8, 6144, 768, 64, 4), (64, 6144, 768, 128, 4), (3, 6144, 768, 32, 4),
    # 7 = in-place f32 accumulate through the TMA reduce-add epilogue (ragged last tile: rows >= M are clipped by the map)
    (1000, 768, 3072, 256, 7), (1576, 768, 768, 256, 7), (333, 768, 768, 128, 7), (65, 768, 256, 64, 7), (12608, 768, 768, 256, 7),
]


@pytest.mark.parametrize("M,N,K,bn,epi", GEMM_CASES)
def test_gemm_against_torch(engine8, M, N, K, bn, epi):
    import torch
    rng = np.random.default_rng(M * 31 + N * 7 + K + bn + epi)
    A = rng.standard_normal((M, K), dtype=np.float32)
    Wt = rng.standard_normal((N, K), dtype=np.float32) * 0.05
    b = rng.standard_normal((N,), dtype=np.float32)
    R = rng.standard_normal((M, N), dtype=np.float32) if epi in (2, 7) else None
    got, am = engine8.test_gemm(epi, bn, A, Wt, b, R)
    ref = torch.from_numpy(_bf16(A)).double() @ torch.from_numpy(_bf16(Wt)).double().T + torch.from_numpy(b).double()
    if epi in (1, 5):
        ref = torch.nn.functional.gelu(ref)       # erf GELU, as the reference (ViT / BERT hidden_act="gelu")
    if epi in (2, 7):
        ref = ref + torch.from_numpy(R).double()
    ref = ref.numpy()
    err = np.abs(got - ref).max()
    if epi in (0, 1):     # bf16 output: half an ulp of the largest magnitude
        assert err <= 2.0 ** -8 * max(1.0, np.abs(ref).max()), err
    else:                 # fp32 output: fp32 accumulation over K
        assert err <= 1e-3, err
    if epi == 4:
        assert np.array_equal(am, np.argmax(got, axis=1))


def test_argmax_ties_resolve_to_lowest_index(engine8):
    """torch.argmax tie-break (generation/utils.py:2793): equal logits -> lowest index."""
    A = np.zeros((4, 768), np.float32)
    Wt = np.zeros((6144, 768), np.float32)
    b = np.zeros((6144,), np.float32)
    b[[100, 4000, 6000]] = 1.0
    _, am = engine8.test_gemm(4, 64, A, Wt, b)
    assert (am == 100).all()
    b[:] = 0.0
    _, am = engine8.test_gemm(4, 64, A, Wt, b)
    assert (am == 0).all()


@pytest.mark.parametrize("n", [1, 3])
def test_encoder_attention_against_torch(engine8, n):
    import torch
    rng = np.random.default_rng(n)
    qkv = rng.standard_normal((n * 197, 2304), dtype=np.float32)
    qkv[:, :768] *= 0.25      # the engine folds 1/sqrt(64) into q; feed pre-scaled queries
    got = engine8.test_encoder_attention(qkv)
    q = torch.from_numpy(_bf16(qkv)).double().view(n, 197, 3, 12, 64)
    Q, K, V = (q[:, :, i].transpose(1, 2) for i in range(3))
    ref = (torch.softmax(Q @ K.transpose(-1, -2), dim=-1) @ V).transpose(1, 2).reshape(n * 197, 768).numpy()
    assert np.abs(got - ref).max() < 2e-2


@pytest.mark.parametrize("M,N,K,epi", [(256, 768, 256, 0), (333, 768, 768, 0), (1576, 2304, 768, 0), (1576, 3072, 768, 1), (1000, 768, 3072, 2)])
def test_pair_gemm_cta_group2_against_torch(engine8, M, N, K, epi):
    """The cta_group::2 variant (two CTAs of a cluster share one 256 x 256 tile) against the same reference."""
    import torch
    rng = np.random.default_rng(M + N + K + epi)
    A = rng.standard_normal((M, K), dtype=np.float32)
    Wt = rng.standard_normal((N, K), dtype=np.float32) * 0.05
    b = rng.standard_normal((N,), dtype=np.float32)
    R = rng.standard_normal((M, N), dtype=np.float32) if epi == 2 else None
    engine8.set_option("gemm_pair", 1)
    try:
        got, _ = engine8.test_gemm(epi, 256, A, Wt, b, R)
    finally:
        engine8.set_option("gemm_pair", 0)
    ref = torch.from_numpy(_bf16(A)).double() @ torch.from_numpy(_bf16(Wt)).double().T + torch.from_numpy(b).double()
    if epi == 1:
        ref = torch.nn.functional.gelu(ref)
    if epi == 2:
        ref = ref + torch.from_numpy(R).double()
    err = np.abs(got - ref.numpy()).max()
    assert err <= (2.0 ** -8 * max(1.0, float(ref.abs().max())) if epi in (0, 1) else 1e-3), err


def test_encoder_attention_mma_sync_variant(engine8):
    """The warp-level mma.sync attention kernel (option attn_tc = 0) stays covered."""
    import torch
    rng = np.random.default_rng(9)
    qkv = rng.standard_normal((2 * 197, 2304), dtype=np.float32)
    qkv[:, :768] *= 0.25
    engine8.set_option("attn_tc", 0)
    try:
        got = engine8.test_encoder_attention(qkv)
    finally:
        engine8.set_option("attn_tc", 1)
    q = torch.from_numpy(_bf16(qkv)).double().view(2, 197, 3, 12, 64)
    Q, K, V = (q[:, :, i].transpose(1, 2) for i in range(3))
    ref = (torch.softmax(Q @ K.transpose(-1, -2), dim=-1) @ V).transpose(1, 2).reshape(2 * 197, 768).numpy()
    assert np.abs(got - ref).max() < 2e-2
