"""GPU tier: every sm_100a kernel, through the C ABI, against a plain torch fp32 reference of the
same op on the same (fp16-rounded) inputs.  Tolerances are written per test."""
import pytest
import torch
import torch.nn.functional as F

from oadp_b200 import binding

pytestmark = pytest.mark.gpu

DEV = 'cuda'


def act_dtype():
    return {'f16': torch.float16, 'bf16': torch.bfloat16}[binding.act_dtype_name()]


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    return t.data_ptr() if t is not None else None


def run_gemm(lib, A, W, bias=None, colsum=None, ln_stats=None, act=0, residual=None, out_stats=None,
             out_f32=False, impl=0, inplace=False):
    M, K = A.shape
    N = W.shape[0]
    if inplace:
        out = residual
    else:
        out = torch.full((M, N), float('nan'), device=DEV, dtype=torch.float32 if out_f32 else act_dtype())
    binding.check(lib.oake_test_gemm(A.data_ptr(), W.data_ptr(), M, N, K, ptr(bias), ptr(colsum), ptr(ln_stats),
                                     act, ptr(residual), ptr(out_stats), out.data_ptr(), int(out_f32), impl,
                                     stream()))
    torch.cuda.synchronize()
    return out


def ref_gemm(A, W, bias=None, act=0, residual=None):
    y = A.float() @ W.float().T
    if bias is not None:
        y = y + bias
    if act == 1:
        y = y * torch.sigmoid(1.702 * y)
    if residual is not None:
        y = y + residual.float()
    return y


def act_tol(scale=1.0):
    """one rounding of a value of magnitude ~scale to the activation type (+ tanh.approx slack)"""
    return scale * (2e-3 if act_dtype() == torch.float16 else 1.6e-2)


@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (128, 256, 128), (300, 768, 768), (1000, 2304, 768),
                                   (77, 512, 768), (4097, 768, 3072), (19000, 3072, 768)])
def test_gemm_tcgen05_plain(lib, M, N, K):
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    A = (torch.randn(M, K, device=DEV, generator=g)).to(act_dtype())
    W = (torch.randn(N, K, device=DEV, generator=g) * K**-0.5).to(act_dtype())
    out = run_gemm(lib, A, W, out_f32=True)
    ref = ref_gemm(A, W)
    # fp32 accumulation of exactly representable products: only summation order differs
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max() < 2e-3, (out - ref).abs().max()


def test_gemm_matches_simt_reference(lib):
    g = torch.Generator(device=DEV).manual_seed(5)
    A = torch.randn(515, 768, device=DEV, generator=g).to(act_dtype())
    W = (torch.randn(2304, 768, device=DEV, generator=g) * 768**-0.5).to(act_dtype())
    bias = torch.randn(2304, device=DEV, generator=g)
    a = run_gemm(lib, A, W, bias=bias, out_f32=True, impl=0)
    b = run_gemm(lib, A, W, bias=bias, out_f32=True, impl=1)
    assert (a - b).abs().max() < 2e-3


def test_gemm_epilogues(lib):
    g = torch.Generator(device=DEV).manual_seed(6)
    M, K = 1030, 768
    A = torch.randn(M, K, device=DEV, generator=g).to(act_dtype())
    # c_fc + QuickGELU -> act output
    W1 = (torch.randn(3072, K, device=DEV, generator=g) * K**-0.5).to(act_dtype())
    b1 = torch.randn(3072, device=DEV, generator=g) * 0.1
    out = run_gemm(lib, A, W1, bias=b1, act=1)
    ref = ref_gemm(A, W1, b1, 1)
    assert out.dtype == act_dtype()
    assert (out.float() - ref).abs().max() < act_tol(8)
    # c_proj + bias + in-place act residual + row statistics of what was stored
    H = out
    W2 = (torch.randn(768, 3072, device=DEV, generator=g) * 3072**-0.5).to(act_dtype())
    b2 = torch.randn(768, device=DEV, generator=g) * 0.1
    x = (torch.randn(M, 768, device=DEV, generator=g) * 2).to(act_dtype())
    ref2 = ref_gemm(H, W2, b2, 0, x.clone())
    stats = torch.zeros(M, 8, 2, device=DEV)
    out2 = run_gemm(lib, H, W2, bias=b2, residual=x, out_stats=stats, inplace=True)
    assert out2.data_ptr() == x.data_ptr()
    assert (out2.float() - ref2).abs().max() < act_tol(8)
    assert (stats[:, 6:] == 0).all()  # 768 columns fill slots 0..5 only
    # statistics are taken on the fp32 values just before the rounding to the activation type
    assert torch.allclose(stats[..., 0].sum(1), ref2.sum(-1), rtol=1e-3, atol=5e-2)
    assert torch.allclose(stats[..., 1].sum(1), (ref2**2).sum(-1), rtol=1e-3, atol=5e-2)
    assert torch.allclose(stats[:, 3, 0], ref2[:, 384:512].sum(-1), rtol=1e-3, atol=5e-2)
    # the CUDA-core reference of the same contract agrees
    x2 = (torch.randn(M, 768, device=DEV, generator=g) * 2).to(act_dtype())
    a = run_gemm(lib, H, W2, bias=b2, residual=x2, impl=0)
    b = run_gemm(lib, H, W2, bias=b2, residual=x2, impl=1)
    assert (a.float() - b.float()).abs().max() < act_tol(8)


def test_gemm_layernorm_fold(lib):
    """y = LN(x) W^T + b evaluated as rstd * (x W'^T - mean * s) + c with statistics from a buffer."""
    from oadp_b200.model import fold_layernorm
    g = torch.Generator(device=DEV).manual_seed(8)
    M, K, N = 777, 768, 2304
    x = (torch.randn(M, K, device=DEV, generator=g) * 3 + 0.7).to(act_dtype())
    W = torch.randn(N, K, device=DEV, generator=g) * K**-0.5
    b = torch.randn(N, device=DEV, generator=g) * 0.1
    gamma = 1 + 0.1 * torch.randn(K, device=DEV, generator=g)
    beta = 0.1 * torch.randn(K, device=DEV, generator=g)
    Wf, s, c = fold_layernorm(W.cpu(), b.cpu(), gamma.cpu(), beta.cpu(), act_dtype())
    Wf, s, c = Wf.to(DEV), s.to(DEV), c.to(DEV)
    xf = x.float()
    stats = torch.zeros(M, 8, 2, device=DEV)
    for j in range(6):  # partial statistics, as the producing GEMM leaves them
        stats[:, j, 0] = xf[:, 128 * j:128 * j + 128].sum(-1)
        stats[:, j, 1] = (xf[:, 128 * j:128 * j + 128]**2).sum(-1)
    out = run_gemm(lib, x, Wf, bias=c, colsum=s, ln_stats=stats, out_f32=False)
    ref = F.layer_norm(xf, (K, ), gamma, beta, 1e-5) @ W.T + b
    assert (out.float() - ref).abs().max() < act_tol(6) + 6e-3, (out.float() - ref).abs().max()


def test_layernorm(lib):
    g = torch.Generator(device=DEV).manual_seed(7)
    x = (torch.randn(1001, 768, device=DEV, generator=g) * 3 + 0.5).to(act_dtype())
    w = 1 + 0.1 * torch.randn(768, device=DEV, generator=g)
    b = 0.1 * torch.randn(768, device=DEV, generator=g)
    out = torch.empty(1001, 768, device=DEV, dtype=act_dtype())
    binding.check(lib.oake_test_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), 1001, stream()))
    torch.cuda.synchronize()
    ref = F.layer_norm(x.float(), (768, ), w, b, 1e-5)
    assert (out.float() - ref).abs().max() < act_tol(5)


def ref_attention(qkv, B, P, side_mask=None):
    """rows [B*P | B | (B)] -> same-layout output; fp32 math on the rounded inputs."""
    W = 768
    q, k, v = qkv.float().split(W, dim=-1)

    def gather(t):  # -> (B, T, 12, 64) tokens ordered patches then class
        pat = t[:B * P].reshape(B, P, 12, 64)
        cls = t[B * P:B * P + B].reshape(B, 1, 12, 64)
        return torch.cat([pat, cls], 1)

    Q, K, V = gather(q).transpose(1, 2), gather(k).transpose(1, 2), gather(v).transpose(1, 2)
    o = torch.softmax(Q @ K.transpose(-1, -2) / 8, -1) @ V  # (B,12,T,64)
    o = o.transpose(1, 2).reshape(B, P + 1, W)
    out = torch.zeros(qkv.shape[0], W, device=qkv.device)
    out[:B * P] = o[:, :P].reshape(B * P, W)
    out[B * P:B * P + B] = o[:, P]
    if side_mask is not None:
        ys = slice(B * P + B, B * P + 2 * B)
        qy = q[ys].reshape(B, 1, 12, 64).transpose(1, 2)
        ky = torch.cat([k[:B * P].reshape(B, P, 12, 64), k[ys].reshape(B, 1, 12, 64)], 1).transpose(1, 2)
        vy = torch.cat([v[:B * P].reshape(B, P, 12, 64), v[ys].reshape(B, 1, 12, 64)], 1).transpose(1, 2)
        bias = torch.cat([side_mask * -100.0, side_mask.new_zeros(B, 1)], 1)[:, None, None, :]
        oy = torch.softmax(qy @ ky.transpose(-1, -2) / 8 + bias, -1) @ vy
        out[ys] = oy.transpose(1, 2).reshape(B, W)
    return out


# B = 40 x 12 heads = 480 items: more than 3 per SM, so the persistent 197-token kernel cycles its
# two-stage ring and both tile placements
@pytest.mark.parametrize('P,B', [(49, 5), (196, 3), (196, 40)])
def test_attention_main(lib, P, B):
    g = torch.Generator(device=DEV).manual_seed(P)
    R = B * (P + 1)
    qkv = (torch.randn(R, 2304, device=DEV, generator=g) * 1.5).to(act_dtype())
    out = torch.zeros(R, 768, device=DEV, dtype=act_dtype())
    binding.check(lib.oake_test_attention_main(qkv.data_ptr(), out.data_ptr(), B, P, stream()))
    torch.cuda.synchronize()
    ref = ref_attention(qkv, B, P)
    tol = 6e-3 if act_dtype() == torch.float16 else 4e-2
    assert (out.float() - ref).abs().max() < tol, (out.float() - ref).abs().max()


@pytest.mark.parametrize('side_only,B', [(1, 5), (0, 5), (0, 53)])
def test_attention_side(lib, side_only, B):
    P = 196
    g = torch.Generator(device=DEV).manual_seed(11)
    R = B * (P + 2)
    qkv = (torch.randn(R, 2304, device=DEV, generator=g) * 1.5).to(act_dtype())
    mask = (torch.rand(B, P, device=DEV, generator=g) > 0.5).float()
    mask[0] = 0
    mask[1] = 1
    out = torch.zeros(R, 768, device=DEV, dtype=act_dtype())
    binding.check(lib.oake_test_attention_side(qkv.data_ptr(), mask.data_ptr(), out.data_ptr(), B, P, side_only,
                                               stream()))
    torch.cuda.synchronize()
    ref = ref_attention(qkv, B, P, mask)
    ys = slice(B * P + B, R)
    tol = 6e-3 if act_dtype() == torch.float16 else 4e-2
    assert (out[ys].float() - ref[ys]).abs().max() < tol
    if side_only:
        assert (out[:B * P + B] == 0).all()  # only the side rows are written
    else:
        assert (out.float() - ref).abs().max() < tol  # main stream and side token from one kernel


@pytest.mark.parametrize('variant', [0, 1])
def test_im2col(lib, variant):
    """T50: conv1's im2col.  T197 (stride 16, pad 15; objects.py:299-301): the 15 x 15 matrix of 16 x 16 blocks of
    the zero-padded crop, which the patch GEMM reads at four row shifts."""
    B = 3
    g = torch.Generator(device=DEV).manual_seed(12)
    px = torch.randn(B, 3, 224, 224, device=DEV, generator=g)
    if variant == 0:
        out = torch.empty(B * 49, 3072, device=DEV, dtype=act_dtype())
        ref = F.unfold(px, 32, stride=32).transpose(1, 2).reshape(B * 49, 3072)
    else:
        out = torch.empty(B * 225, 768, device=DEV, dtype=act_dtype())
        ref = F.unfold(F.pad(px, (15, 1, 15, 1)), 16, stride=16).transpose(1, 2).reshape(B * 225, 768)
    binding.check(lib.oake_test_im2col(px.data_ptr(), out.data_ptr(), B, variant, stream()))
    torch.cuda.synchronize()
    assert torch.equal(out, ref.to(act_dtype()))  # pure data movement + one rounding: bit-exact


@pytest.mark.parametrize('variant,B', [(0, 3), (1, 1), (1, 5), (1, 37)])
def test_patch_embed_matches_conv2d(lib, variant, B):
    """Front-end matrix + conv1 GEMM against F.conv2d on the same rounded operands (objects.py:299-301 for T197:
    the shifted-A GEMM over the block matrix must equal the stride-16 / pad-15 convolution)."""
    g = torch.Generator(device=DEV).manual_seed(13 + B)
    px = torch.randn(B, 3, 224, 224, device=DEV, generator=g)
    w = (torch.randn(768, 3, 32, 32, device=DEV, generator=g) * 0.02).to(act_dtype())
    stride, pad, grid, pb = ((32, 0, 7, 49), (16, 15, 14, 225))[variant]
    out = torch.full((B * pb, 768), float('nan'), device=DEV)
    binding.check(lib.oake_test_patch_embed(px.data_ptr(), w.data_ptr(), out.data_ptr(), B, variant, stream()))
    torch.cuda.synchronize()
    ref = F.conv2d(px.to(act_dtype()).double(), w.double(), stride=stride, padding=pad)  # [B, 768, grid, grid]
    ref = ref.permute(0, 2, 3, 1)
    if variant == 0:
        got = out.view(B, 7, 7, 768)
    else:
        got = out.view(B, 15, 15, 768)[:, :14, :14]
    assert torch.isfinite(got).all()
    assert (got.double() - ref).abs().max() < 2e-3  # fp32 accumulation over K = 3072 of O(1) x O(0.02) products
