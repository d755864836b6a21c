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
    # more tiles than SMs and not a multiple of them: the stream-K work split (boundary tiles finished from a partial of the
    # neighbouring CTA), every epilogue family; M = 64 x 197 is the headline's encoder
    (12608, 2304, 768, 256, 0), (12608, 3072, 768, 256, 1), (12608, 768, 3072, 256, 7), (12608, 768, 3072, 128, 7),
    (20000, 768, 256, 256, 2), (9000, 768, 768, 64, 0),
    # bn = 0: the cluster K-split kernel of the large-batch decoder program (gemm_ksplit.cuh): four CTAs per 128 x 128 tile, partial
    # chunks exchanged through distributed shared memory, fixed-order sum, TMA reduce-add into the residual
    (512, 768, 3072, 0, 7), (512, 768, 768, 0, 7), (300, 768, 768, 0, 7), (130, 768, 3072, 0, 7), (1, 768, 768, 0, 7),
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
    if epi in (0, 1):     # bf16 output: half an ulp of the element's binade, plus the fp32 accumulation error that can flip a rounding
        ulp_half = 2.0 ** (np.floor(np.log2(np.maximum(np.abs(ref), 2.0 ** -20))) - 8)
        assert (np.abs(got - ref) <= 2.0 * ulp_half + 1e-4).all(), float((np.abs(got - ref) / ulp_half).max())
        assert err <= 2.0 ** -7 * max(1.0, np.abs(ref).max()), err
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


# ---- decoder stage kernels (decode_stages.cuh) against fp64 references of the same op ----

@pytest.fixture(scope="module")
def engine_dec(weights0):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from manga_ocr_b200.engine import Engine
    e = Engine(weights0, device=0, max_batch=70, max_length=300)
    yield e
    e.close()


def _attn_ref(q, k, v, n_keys):
    """softmax(q_h . K_h^T) V_h per row and head over the first n_keys[b] keys, fp64 (q is pre-scaled)."""
    import torch
    n = q.shape[0]
    out = np.zeros((n, 768))
    for b in range(n):
        nk = int(n_keys[b])
        Q = torch.from_numpy(_bf16(q[b])).double().view(12, 1, 64)
        K = torch.from_numpy(_bf16(k[b, :nk])).double().view(nk, 12, 64).transpose(0, 1)
        V = torch.from_numpy(_bf16(v[b, :nk])).double().view(nk, 12, 64).transpose(0, 1)
        out[b] = (torch.softmax(Q @ K.transpose(-1, -2), dim=-1) @ V).reshape(768).numpy()
    return out


@pytest.mark.parametrize("mode", [1, 3], ids=["four_warp", "warp_per_unit"])
@pytest.mark.parametrize("n_rows,pos", [(3, [0, 1, 5]), (5, [31, 32, 127, 128, 129]), (6, [30, 33, 63, 64, 65, 96]), (4, [255, 256, 257, 298]), (66, None)])
def test_decode_self_attention_against_torch(engine_dec, n_rows, pos, mode):
    """One query over 1..299 cached keys (BertSelfAttention with a cache, modeling_bert.py:143-207): every staged-block
    boundary (32-key warp ranges / chunks, 128-key blocks) and the append of this step's K/V row; both attention
    kernels (mode 3 = the warp-per-unit kernel of the large-batch program)."""
    rng = np.random.default_rng(n_rows)
    pos = np.asarray(pos if pos is not None else rng.integers(0, 299, n_rows), np.int32)
    n_ctx = int(pos.max()) + 1
    q = rng.standard_normal((n_rows, 768), dtype=np.float32) * 0.3
    k = rng.standard_normal((n_rows, n_ctx, 768), dtype=np.float32)
    v = rng.standard_normal((n_rows, n_ctx, 768), dtype=np.float32)
    nk = rng.standard_normal((n_rows, 768), dtype=np.float32)
    nv = rng.standard_normal((n_rows, 768), dtype=np.float32)
    ctx, k_row, v_row = engine_dec.test_decode_attention(mode, q, k, v, pos=pos, new_k=nk, new_v=nv)
    kk, vv = k.copy(), v.copy()
    for b in range(n_rows):
        kk[b, pos[b]] = nk[b]
        vv[b, pos[b]] = nv[b]
    ref = _attn_ref(q, kk, vv, pos + 1)
    assert np.abs(ctx - ref).max() < 2e-2, float(np.abs(ctx - ref).max())
    assert np.array_equal(k_row, _bf16(nk)) and np.array_equal(v_row, _bf16(nv))       # the cache append is exact


@pytest.mark.parametrize("mode", [0, 2, 4], ids=["f32_partial_query", "bf16_query", "warp_per_unit"])
def test_decode_cross_attention_against_torch(engine_dec, mode):
    """One query over the 197 encoder keys (BertSelfAttention as cross-attention, modeling_bert.py:210-284), no mask."""
    rng = np.random.default_rng(7 + mode)
    n = 9
    q = rng.standard_normal((n, 768), dtype=np.float32) * 0.3
    k = rng.standard_normal((n, 197, 768), dtype=np.float32)
    v = rng.standard_normal((n, 197, 768), dtype=np.float32)
    ctx = engine_dec.test_decode_attention(mode, q, k, v)
    ref = _attn_ref(q, k, v, np.full(n, 197))
    assert np.abs(ctx - ref).max() < 2e-2, float(np.abs(ctx - ref).max())


@pytest.mark.parametrize("kind,n_rows,N,K", [(0, 64, 2304, 768), (0, 5, 768, 768), (1, 64, 3072, 768), (1, 17, 3072, 768), (2, 64, 768, 768),
                                             (2, 64, 768, 3072), (2, 70, 768, 3072), (3, 64, 6144, 768), (3, 3, 6144, 768)])
def test_decode_stage_gemm_against_torch(engine_dec, kind, n_rows, N, K):
    """The small-M mma.sync GEMM stages (BertSelfAttention / BertIntermediate / BertOutput Linear layers at M = batch rows)."""
    import torch
    rng = np.random.default_rng(kind * 1000 + n_rows + N + K)
    A = rng.standard_normal((n_rows, K), dtype=np.float32)
    Wt = rng.standard_normal((N, K), dtype=np.float32) * 0.05
    b = rng.standard_normal((N,), dtype=np.float32)
    got, am = engine_dec.test_stage_gemm(kind, A, Wt, b)
    ref = torch.from_numpy(_bf16(A)).double() @ torch.from_numpy(_bf16(Wt)).double().T + torch.from_numpy(b).double()
    if kind == 1:
        ref = torch.nn.functional.gelu(ref)
    ref = ref.numpy()
    err = np.abs(got - ref).max()
    if kind in (0, 1):
        assert err <= 2.0 ** -8 * max(1.0, np.abs(ref).max()), err
    else:
        assert err <= 1e-3, err
    if kind == 3:
        assert np.array_equal(am, np.argmax(got, axis=1))


@pytest.mark.parametrize("kind", [4, 5], ids=["cluster_proj_ln", "split_k_plus_ln_stage"])
@pytest.mark.parametrize("n_rows,gelu,with_resid", [(64, False, True), (7, False, True), (33, True, False), (70, False, True)])
def test_decode_projection_layernorm_against_torch(engine_dec, kind, n_rows, gelu, with_resid):
    """LayerNorm([gelu](A W^T + b) + resid): BertSelfOutput / BertOutput (modeling_bert.py:287-298, 343-356) and the LM-head
    transform (:471-486) - in the 16-CTA cluster kernel (row statistics exchanged through distributed shared memory) and as
    split-K partials + the LayerNorm row stage; ragged last row group, rows beyond one cluster wave."""
    import torch
    rng = np.random.default_rng(100 * kind + n_rows)
    A = rng.standard_normal((n_rows, 768), dtype=np.float32)
    Wt = rng.standard_normal((768, 768), dtype=np.float32) * 0.05
    b = rng.standard_normal((768,), dtype=np.float32)
    R = rng.standard_normal((n_rows, 768), dtype=np.float32) + 0.5 if with_resid else None     # a non-zero row mean
    g = (1.0 + 0.1 * rng.standard_normal(768)).astype(np.float32)
    be = (0.1 * rng.standard_normal(768)).astype(np.float32)
    got, _ = engine_dec.test_stage_gemm(kind, A, Wt, b, resid=R, gamma=g, beta=be, gelu=gelu)
    y = torch.from_numpy(_bf16(A)).double() @ torch.from_numpy(_bf16(Wt)).double().T + torch.from_numpy(b).double()
    if gelu:
        y = torch.nn.functional.gelu(y)
    if R is not None:
        y = y + torch.from_numpy(R).double()
    ref = torch.nn.functional.layer_norm(y, (768,), torch.from_numpy(g).double(), torch.from_numpy(be).double(), eps=1e-12).numpy()
    assert np.abs(got - ref).max() <= 2e-3, float(np.abs(got - ref).max())
