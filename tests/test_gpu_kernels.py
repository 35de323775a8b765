"""Kernel-level parity on a B200, every call through the C-ABI (ctypes -> libcre_b200.so).  Floating-point
kernels are compared with a plain fp32 PyTorch statement of the same op; tolerances are bf16-rounding level
and written next to each check.  Index work (top-k, merge) is bit-exact."""
import numpy as np
import pytest
import torch

from oracle import common, preprocess_ref, reid_ref
from vision_sam3_yolo_lameless_b200 import _lib

pytestmark = pytest.mark.gpu


def rel_err(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert torch.isfinite(got).all(), "non-finite output"
    return ((got - ref).abs().max() / (ref.abs().max() + 1e-12)).item()


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (1, 64, 64), (1000, 768, 768), (4021, 2304, 768), (515, 768, 3072),
                                   (300, 1024, 1024), (257, 192, 128)])
def test_gemm_epilogues(engine_small, cg, m, n, k):
    eng, dev = engine_small, engine_small.device
    g = torch.Generator(device=dev).manual_seed(m * 7 + n)
    a = (torch.randn(m, k, device=dev, generator=g) * 0.5).to(torch.bfloat16)
    b = (torch.randn(n, k, device=dev, generator=g) * 0.5).to(torch.bfloat16)
    bias = torch.randn(n, device=dev, generator=g)
    ref = a.float() @ b.float().t() + bias
    assert rel_err(eng.gemm(a, b, _lib.EPI_F32, bias=bias, cta_group=cg), ref) < 2e-5          # fp32 accumulate
    assert rel_err(eng.gemm(a, b, _lib.EPI_BF16, bias=bias, cta_group=cg), ref) < 5e-3         # bf16 output rounding
    assert rel_err(eng.gemm(a, b, _lib.EPI_GELU, bias=bias, cta_group=cg), torch.nn.functional.gelu(ref)) < 5e-3
    scale = torch.rand(n, device=dev, generator=g) + 0.5
    res = torch.randn(m, n, device=dev, generator=g)
    out = res.clone()
    eng.gemm(a, b, _lib.EPI_RESID, bias=bias, scale=scale, out=out, cta_group=cg)
    assert rel_err(out, res + scale * ref) < 2e-5
    out = eng.gemm(a, b, _lib.EPI_F32, bias=None, cta_group=cg)
    assert rel_err(out, ref - bias) < 2e-5


def test_gelu_epilogue_error(engine_small):
    """The up-projection epilogue's GELU against nn.functional.gelu (erf form, HF:activations.py) on a dense grid of
    pre-activations: acc[m, n] = x_m + bias_n covers [-12, 12] in steps of 2^-10.  Allowed: the bf16 rounding of the output
    (2^-8 relative, half-ulp plus slack) plus 1e-3 absolute for the approximation (gemm_tcgen05.cuh gelu2: fit 2.5e-5 +
    the MUFU tanh's 2^-11 relative error times |x| / 2)."""
    eng, dev = engine_small, engine_small.device
    xs = torch.arange(-12.0, 12.0, 1.0 / 16.0, device=dev)                       # bf16-exact
    a = torch.zeros(xs.numel(), 64, device=dev)
    a[:, 0] = xs
    b = torch.zeros(64, 64, device=dev)
    b[:, 0] = 1.0
    bias = torch.arange(64, device=dev, dtype=torch.float32) / 1024.0
    got = eng.gemm(a.to(torch.bfloat16), b.to(torch.bfloat16), _lib.EPI_GELU, bias=bias, cta_group=1).double().cpu()
    acc = (xs[:, None] + bias[None, :]).double().cpu()
    ref = torch.nn.functional.gelu(acc)
    err = (got - ref).abs()
    excess = (err - ref.abs() * 2.0 ** -8).clamp_min(0)
    i = int(excess.argmax())
    print(f"gelu epilogue: max |err| {err.max():.3e}, max excess over bf16 rounding {excess.max():.3e} at x = {acc.flatten()[i]:.4f}; "
          f"max |err| on x < 0: {err[acc < 0].max():.3e}")
    assert excess.max() < 1e-3
    assert (got[acc < -9.0].abs() < 1e-3).all() and torch.isfinite(got).all()


def test_gemm_rejects_bad_shapes(engine_small):
    eng, dev = engine_small, engine_small.device
    a = torch.zeros(8, 100, device=dev, dtype=torch.bfloat16)
    b = torch.zeros(32, 100, device=dev, dtype=torch.bfloat16)
    with pytest.raises(_lib.CreError, match="multiple of 64"):
        eng.gemm(a, b)
    with pytest.raises(_lib.CreError):
        eng.gemm(a[:, :64].contiguous(), b[:, :64].contiguous(), epilogue=9)


@pytest.mark.parametrize("rows,dim", [(1003, 768), (77, 1024), (1, 768)])
def test_layernorm(engine_small, rows, dim):
    dev = engine_small.device
    x = torch.randn(rows, dim, device=dev) * 3 + 1.5
    g, b = torch.randn(dim, device=dev), torch.randn(dim, device=dev)
    ref = torch.nn.functional.layer_norm(x, (dim,), g, b, 1e-5)
    assert rel_err(engine_small.layernorm(x, g, b), ref) < 5e-3                                 # bf16 output


def _stats_combine(stats, dim):
    """numpy restatement of the consumer-side combination of a statistics row -> (pivot, mean, biased variance)."""
    st = stats.double().cpu()
    s = dim // 128
    means, m2s = st[:, 4:4 + 2 * s:2], st[:, 5:5 + 2 * s:2]
    mean = means.mean(dim=1)
    m2 = m2s.sum(dim=1) + 128.0 * ((means - mean[:, None]) ** 2).sum(dim=1)
    return st[:, 0], mean, m2 / dim


@pytest.mark.parametrize("rows,dim", [(1003, 768), (77, 1024), (1, 768)])
def test_row_stats(engine_small, rows, dim):
    dev = engine_small.device
    x = torch.randn(rows, dim, device=dev) * 3 + torch.randn(rows, 1, device=dev) * 5
    xb, stats = engine_small.row_stats(x)
    pivot, mean, var = _stats_combine(stats, dim)
    xd = x.double().cpu()
    assert (pivot - xd.mean(dim=1)).abs().max() < 1e-5 and (mean - xd.mean(dim=1)).abs().max() < 1e-5
    assert ((var - xd.var(dim=1, unbiased=False)).abs() / xd.var(dim=1, unbiased=False)).max() < 1e-5
    assert rel_err(xb, x - x.mean(dim=1, keepdim=True)) < 5e-3                                  # bf16 of the centred row


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("m,n,dim,epi", [(1000, 2304, 768, "bf16"), (515, 3072, 768, "gelu"), (300, 4096, 1024, "gelu"), (1, 64, 768, "bf16")])
def test_gemm_layernorm_folded_consumer(engine_small, cg, m, n, dim, epi):
    """LN(x) W^T + b through row_stats + fold_ln_weights + cre_gemm_ln == fp32 layer_norm + matmul, at the accuracy of the
    unfolded bf16 path (bf16 LN output x bf16 weights), including rows with a large mean and outlier channels."""
    eng, dev = engine_small, engine_small.device
    g = torch.Generator(device=dev).manual_seed(m + n)
    x = torch.randn(m, dim, device=dev, generator=g) * 2 + torch.randn(m, 1, device=dev, generator=g) * 4
    x[:, 5] += 60.0                                                                             # a massive-activation channel
    w = (torch.randn(n, dim, device=dev, generator=g) * 0.05).to(torch.bfloat16)
    gamma, beta = torch.rand(dim, device=dev, generator=g) + 0.5, torch.randn(dim, device=dev, generator=g) * 0.3
    bias = torch.randn(n, device=dev, generator=g)
    ref = torch.nn.functional.layer_norm(x, (dim,), gamma, beta, 1e-5) @ w.float().t() + bias
    wf, c1, c2 = eng.fold_ln_weights(w, gamma, beta, bias)
    assert rel_err(wf, w.float() * gamma) < 5e-3 and rel_err(c2, bias + w.float() @ beta) < 1e-5
    assert rel_err(c1, wf.float().sum(dim=1)) < 1e-5
    ws, c1s, c2s = eng.fold_ln_weights(w, gamma, beta, bias, scaled_rows=min(n, 32), row_scale=0.125)   # q rows: exact 2^-3
    rs = torch.ones(n, device=dev)
    rs[:32] = 0.125
    assert torch.equal(ws.float(), wf.float() * rs[:, None]) and torch.equal(c1s, c1 * rs) and torch.equal(c2s, c2 * rs)
    xb, stats = eng.row_stats(x)
    e = _lib.EPI_BF16 if epi == "bf16" else _lib.EPI_GELU
    out = eng.gemm_ln(xb, wf, e, stats, dim, bias=c2, c1=c1, cta_group=cg)
    want = ref if epi == "bf16" else torch.nn.functional.gelu(ref)
    unfolded = eng.gemm(eng.layernorm(x, gamma, beta), w, e, bias=bias, cta_group=cg)
    assert rel_err(out, want) < max(8e-3, 2.0 * rel_err(unfolded, want))


@pytest.mark.parametrize("epi", [_lib.EPI_RESID_LN, _lib.EPI_RESID_LN3])
@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("m,n,k", [(1000, 768, 768), (515, 768, 3072), (300, 1024, 1024), (1, 768, 64), (40021, 768, 768)])
def test_gemm_resid_layernorm_producer(engine_small, cg, epi, m, n, k):
    """x += scale * (a W^T + b) with the LayerNorm statistics of the NEW x and bf16(x - pivot), pivot = the row mean recorded
    in the incoming statistics; then the chain into a folded consumer reproduces LN(x_new) W2^T."""
    eng, dev = engine_small, engine_small.device
    g = torch.Generator(device=dev).manual_seed(m * 3 + k)
    x0 = torch.randn(m, n, device=dev, generator=g) * 2 + torch.randn(m, 1, device=dev, generator=g) * 3
    a = (torch.randn(m, k, device=dev, generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(n, k, device=dev, generator=g) * 0.1).to(torch.bfloat16)
    bias, scale = torch.randn(n, device=dev, generator=g), torch.rand(n, device=dev, generator=g) + 0.5
    _, stats0 = eng.row_stats(x0)
    ref = x0 + scale * (a.float() @ w.float().t() + bias)
    x, xb, stats1 = eng.gemm_ln(a, w, epi, stats0, n, bias=bias, scale=scale, out=x0.clone(), cta_group=cg)
    assert rel_err(x, ref) < 2e-5                                                               # fp32 accumulate + fp32 update
    pivot, mean, var = _stats_combine(stats1, n)
    rd = ref.double().cpu()
    assert (pivot - x0.double().cpu().mean(dim=1)).abs().max() < 1e-4
    assert (mean - rd.mean(dim=1)).abs().max() < 1e-4
    assert ((var - rd.var(dim=1, unbiased=False)).abs() / rd.var(dim=1, unbiased=False)).max() < 1e-4
    assert rel_err(xb, ref - x0.mean(dim=1, keepdim=True)) < 5e-3                               # bf16 of x - pivot
    # chain: the folded consumer on (xb, stats1) == LN(x_new) W2^T + b2
    n2 = 256
    w2 = (torch.randn(n2, n, device=dev, generator=g) * 0.05).to(torch.bfloat16)
    gamma, beta = torch.rand(n, device=dev, generator=g) + 0.5, torch.randn(n, device=dev, generator=g) * 0.3
    b2 = torch.randn(n2, device=dev, generator=g)
    wf, c1, c2 = eng.fold_ln_weights(w2, gamma, beta, b2)
    out = eng.gemm_ln(xb, wf, _lib.EPI_BF16, stats1, n, bias=c2, c1=c1, cta_group=cg)
    want = torch.nn.functional.layer_norm(ref, (n,), gamma, beta, 1e-5) @ w2.float().t() + b2
    assert rel_err(out, want) < 8e-3


@pytest.mark.parametrize("epi", [_lib.EPI_RESID_SP, _lib.EPI_RESID_SP3])
@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("m,n,k", [(1000, 768, 768), (515, 768, 3072), (300, 1024, 1024), (1, 768, 64), (40021, 768, 768)])
def test_gemm_resid_split_producer(engine_small, cg, epi, m, n, k):
    """The residual stream as two bf16 halves around the row pivot (what cre_vit_forward runs): x = pivot + hi + lo.  Two chained
    updates x += scale * (a W^T + b): the reconstructed stream stays within 2^-16 of the centred magnitude of the fp32 result (bf16
    pair = 16 significant bits), the statistics describe the NEW x, the pivot moves to the previous row mean, hi alone is the
    bf16 A operand of the next folded GEMM, and the folded consumer on (hi, stats) reproduces LN(x_new) W2^T."""
    eng, dev = engine_small, engine_small.device
    g = torch.Generator(device=dev).manual_seed(m * 3 + k)
    x0 = torch.randn(m, n, device=dev, generator=g) * 2 + torch.randn(m, 1, device=dev, generator=g) * 3
    x0[:, 7] += 60.0                                                                            # a massive-activation channel
    hi, lo, stats = eng.row_stats_split(x0)
    mean0 = x0.mean(dim=1, keepdim=True)
    assert ((mean0 + hi.float() + lo.float()) - x0).abs().max() < 2.0 ** -15 * (x0 - mean0).abs().max()
    ref = x0
    for step in range(2):
        a = (torch.randn(m, k, device=dev, generator=g) * 0.5).to(torch.bfloat16)
        w = (torch.randn(n, k, device=dev, generator=g) * 0.1).to(torch.bfloat16)
        bias, scale = torch.randn(n, device=dev, generator=g), torch.rand(n, device=dev, generator=g) + 0.5
        prev_mean = ref.mean(dim=1, keepdim=True)
        ref = ref + scale * (a.float() @ w.float().t() + bias)
        hi, lo, stats = eng.gemm_ln(a, w, epi, stats, n, bias=bias, scale=scale, out=(hi, lo), cta_group=cg)
        pivot, mean, var = _stats_combine(stats, n)
        rd = ref.double().cpu()
        assert (pivot - prev_mean.double().cpu()[:, 0]).abs().max() < 1e-4                      # pivot = the mean at the previous norm
        assert (mean - rd.mean(dim=1)).abs().max() < 1e-4
        assert ((var - rd.var(dim=1, unbiased=False)).abs() / rd.var(dim=1, unbiased=False)).max() < 1e-4
        x = stats[:, :1] + (hi.float() + lo.float())
        cen = (ref - prev_mean).abs().amax(dim=1, keepdim=True)
        assert ((x - ref).abs() / cen).max() < (step + 1) * 2.0 ** -15 + 2e-6                     # split rounding + the fp32 update itself
        assert rel_err(hi, ref - prev_mean) < 5e-3                                              # bf16 of x - pivot
    n2 = 256
    w2 = (torch.randn(n2, n, device=dev, generator=g) * 0.05).to(torch.bfloat16)
    gamma, beta = torch.rand(n, device=dev, generator=g) + 0.5, torch.randn(n, device=dev, generator=g) * 0.3
    b2 = torch.randn(n2, device=dev, generator=g)
    wf, c1, c2 = eng.fold_ln_weights(w2, gamma, beta, b2)
    out = eng.gemm_ln(hi, wf, _lib.EPI_BF16, stats, n, bias=c2, c1=c1, cta_group=cg)
    want = torch.nn.functional.layer_norm(ref, (n,), gamma, beta, 1e-5) @ w2.float().t() + b2
    assert rel_err(out, want) < 8e-3


def _attention_case(engine, t, n, heads, q_scale=0.125, plant=None, seed=None):
    """qkv with N(0,1) q / k / v (q times q_scale) -> (kernel output, fp32 softmax reference of HF:modeling_dinov3_vit.py:210-235
    on the same bf16 inputs).  plant = (key, boost): that key's logit is lifted `boost` nats above every other key of every row."""
    dev = engine.device
    d = heads * 64
    gen = torch.Generator(device=dev).manual_seed(t if seed is None else seed)
    q, k, v = (torch.randn(n, t, heads, 64, device=dev, generator=gen) for _ in range(3))
    if plant is not None:
        # every q gets a large component along e0, the planted key is +e0 scaled, every other key has no e0 component
        key, boost = plant
        k[:, :, :, 0] = 0.0
        q[:, :, :, 0] = 16.0 / q_scale                      # stored q[..., 0] = 16 (bf16-exact)
        k[:, key, :, 0] = boost / 16.0
    qs, kb, vb = (q * q_scale).to(torch.bfloat16), k.to(torch.bfloat16), v.to(torch.bfloat16)
    qkv = torch.cat([qs.reshape(n * t, d), kb.reshape(n * t, d), vb.reshape(n * t, d)], dim=1).contiguous()
    out = engine.attention(qkv, n, t, heads)
    att = torch.softmax(qs.double().permute(0, 2, 1, 3) @ kb.double().permute(0, 2, 3, 1), dim=-1)
    ref = (att @ vb.double().permute(0, 2, 1, 3)).permute(0, 2, 1, 3).reshape(n * t, d)
    return out, ref.float()


@pytest.mark.parametrize("t,n,heads", [(201, 3, 12), (1029, 1, 12), (1374, 1, 12), (37, 2, 16), (256, 2, 12), (257, 1, 12),
                                       (9, 2, 12), (16, 1, 12), (31, 2, 12), (48, 1, 12), (185, 2, 12), (129, 5, 12), (240, 2, 16),
                                       (161, 2, 12), (176, 3, 12), (177, 2, 12), (192, 2, 16), (200, 1, 12), (208, 2, 12), (209, 2, 12),
                                       (201, 70, 12), (193, 40, 16),
                                       (300, 3, 12), (288, 2, 16), (289, 3, 12), (385, 1, 12), (1029, 2, 16), (640, 7, 12)])
def test_attention(engine_small, t, n, heads):
    """All three tensor-core kernels (long-sequence persistent: T > 256; single-S persistent: T <= 160 and 208 < T <= 256; split-S:
    160 < T <= 208) incl. launches with several units per CTA ((201, 70, 12) = 840 units: both halves-first orders of the split-S
    kernel) and the long kernel's edges: a full last key block (288 = 3 x 96), a last block with ONE valid key (289, 385), an odd
    number of units, query tiles without a live row in some warps (300: rows 256..299 of the third tile), more units than streams
    (640 x 7 x 12 = 420 units)."""
    out, ref = _attention_case(engine_small, t, n, heads)
    assert rel_err(out, ref) < 8e-3                                                             # bf16 P and bf16 output


@pytest.mark.parametrize("t,n", [(201, 14), (129, 3), (256, 2), (1029, 1)])
@pytest.mark.parametrize("sigma", [8.0, 30.0])
def test_attention_large_logits(engine_small, t, n, sigma):
    """Trained weights can give logits far from N(0, 1).  The kernels' fixed 32-key stabiliser is only a shift: rows whose later
    keys outrun it are caught by the row-sum flag and recomputed exactly (csrc/attention.cu "Exactness").  Logit sigma 8 and 30 nats
    (peaked, near one-hot rows); same tolerance as the small-logit cases."""
    out, ref = _attention_case(engine_small, t, n, 12, q_scale=sigma / 8.0, seed=1000 + t)     # q.k has sigma 8 for unit q, k
    assert rel_err(out, ref) < 8e-3


@pytest.mark.parametrize("t,n,key", [(201, 14, 150), (201, 2, 200), (201, 2, 40), (129, 3, 100), (256, 2, 255), (1029, 1, 900)])
def test_attention_planted_late_key(engine_small, t, n, key):
    """One key 100 nats above every other key of every row, outside the stabiliser's 32-key prefix: exp overflows the single-pass
    kernels, every unit must be flagged and come back from the exact kernel as (almost) that key's value row."""
    out, ref = _attention_case(engine_small, t, n, 12, plant=(key, 100.0), seed=2000 + t)
    assert rel_err(out, ref) < 8e-3


@pytest.mark.parametrize("h,w,kind,bgr", [(1080, 1920, "noise", True), (720, 1280, "smooth", True), (224, 224, "noise", False),
                                          (270, 482, "noise", True), (100, 60, "noise", True), (2160, 3840, "noise", True)])
def test_preprocess_vs_oracle(engine_small, h, w, kind, bgr):
    fr = common.noise_frames(2, h, w, seed=3) if kind == "noise" else common.smooth_frames(2, h, w, seed=3)
    ref = preprocess_ref.patchify(preprocess_ref.preprocess(fr, bgr=bgr))
    out = engine_small.preprocess(torch.from_numpy(fr).to(engine_small.device), bgr=bgr)
    assert out.shape == (2 * 196, 768) and out.dtype == torch.bfloat16
    err = (out.float().cpu() - torch.from_numpy(ref)).abs().max().item()
    assert err < 1.2e-2, err                    # |x| <= 2.64 -> bf16 half-ulp 7.8e-3, plus fp32 summation-order noise


@pytest.mark.parametrize("bgr", [True, False])
def test_preprocess_no_resize_kernel_bit_identical(engine_small, bgr):
    """Frames that are already 224 x 224 take the normalise + patchify kernel; it must reproduce the filtering kernel (identity
    taps) bit for bit, for dense and for row-pitched input, and match the oracle."""
    eng, dev = engine_small, engine_small.device
    fr = common.noise_frames(5, 224, 224, seed=21)
    dense = torch.from_numpy(fr).to(dev)
    pitched = torch.zeros((5, 230, 240, 3), dtype=torch.uint8, device=dev)[:, :224, :224, :]    # row pitch 720 B, frame pitch 165 600 B
    pitched.copy_(dense)
    fast = eng.preprocess(dense, bgr=bgr).clone()
    fast_p = eng.preprocess(pitched, bgr=bgr).clone()
    _lib.set_tuning("preprocess_identity", 0)
    try:
        slow = eng.preprocess(dense, bgr=bgr).clone()
    finally:
        _lib.set_tuning("preprocess_identity", 1)
    assert torch.equal(fast.view(torch.int16), slow.view(torch.int16))
    assert torch.equal(fast_p.view(torch.int16), slow.view(torch.int16))
    # rows that are not 16-byte aligned (pitch 690 B, odd base offsets): the shifted-load path of the same kernel, same bits; the last
    # frame ends exactly at the end of its allocation slice (no bytes to spare behind the last row)
    for off in (0, 2, 5, 13):
        flat = torch.zeros(5 * 224 * 230 * 3 + 16, dtype=torch.uint8, device=dev)
        odd = flat[off:off + 5 * 224 * 230 * 3].view(5, 224, 230, 3)[:, :, :224, :]
        odd.copy_(dense)
        assert torch.equal(eng.preprocess(odd, bgr=bgr).view(torch.int16), slow.view(torch.int16)), off
    ref = preprocess_ref.patchify(preprocess_ref.preprocess(fr, bgr=bgr))
    assert (fast.float().cpu() - torch.from_numpy(ref)).abs().max().item() < 1.2e-2


@pytest.mark.parametrize("h,w", [(1080, 1920), (720, 1280), (2160, 3840), (1000, 1500)])
def test_preprocess_variants_bit_identical(engine_small, h, w):
    """The three filtering kernels -- direct-load (0), TMA-staged one row per warp (1), TMA-staged row pairs (2, the default) -- evaluate
    the same integer vertical pass and the same horizontal FMA chain: identical bits, BGR and RGB, noise frames (every tap matters)."""
    dev = engine_small.device
    fr = torch.from_numpy(common.noise_frames(3, h, w, seed=h + w)).to(dev)
    outs = {}
    try:
        for mode in (0, 1, 2):
            _lib.set_tuning("preprocess_tma", mode)
            outs[mode] = (engine_small.preprocess(fr, bgr=True).clone(), engine_small.preprocess(fr, bgr=False).clone())
    finally:
        _lib.set_tuning("preprocess_tma", 2)
    for mode in (1, 2):
        for a, b in zip(outs[0], outs[mode]):
            assert torch.equal(a.view(torch.int16), b.view(torch.int16)), mode


def test_preprocess_vs_hf_processor_golden(engine_small, golden):
    from oracle.make_golden import PREPROCESS_CASES, frames_for
    want = np.load(golden / "preprocess.npz")
    for name, n, h, w, kind, seed in PREPROCESS_CASES:
        fr = frames_for(kind, n, h, w, seed)
        out = engine_small.preprocess(torch.from_numpy(fr).to(engine_small.device), bgr=True).float().cpu().numpy()
        ref = preprocess_ref.patchify(want[name][None])
        assert np.abs(out - ref).max() < 1.2e-2, name


def test_preprocess_plain_bilinear_would_fail(engine_small):
    """The antialias filter is mandatory: a non-antialiased resize of noise differs grossly (hard part 3)."""
    fr = common.noise_frames(1, 1080, 1920, seed=9)
    x = torch.from_numpy(fr[..., ::-1].copy()).permute(0, 3, 1, 2).float() / 255
    plain = torch.nn.functional.interpolate(x, (224, 224), mode="bilinear", antialias=False)
    plain = (plain - torch.tensor(preprocess_ref.MEAN).view(1, 3, 1, 1)) / torch.tensor(preprocess_ref.STD).view(1, 3, 1, 1)
    out = engine_small.preprocess(torch.from_numpy(fr).to(engine_small.device), bgr=True).float().cpu().numpy()
    assert np.abs(out - preprocess_ref.patchify(plain.numpy())).max() > 0.5


def test_preprocess_rois_vs_oracle(engine_small):
    """K1 in region-of-interest mode == the oracle on the numpy crop; a ROI covering the whole frame reproduces the full-frame
    call bit for bit (the device-built tables equal the cached host tables)."""
    eng, dev = engine_small, engine_small.device
    fr = common.smooth_frames(2, 720, 1280, seed=12)
    fr[1] = common.noise_frames(1, 720, 1280, seed=13)[0]
    frd = torch.from_numpy(fr).to(dev)
    rois = [(0, 100, 50, 500, 400), (1, 0, 0, 1280, 720), (1, 601, 333, 777, 700), (0, 1000, 500, 1280, 720), (0, 37, 411, 150, 500),
            (1, 3, 7, 4, 8), (0, 0, 0, 1280, 720)]
    out = eng.preprocess_rois(frd, rois).clone().view(len(rois), 196, 768)
    for i, (f, x0, y0, x1, y1) in enumerate(rois):
        ref = preprocess_ref.patchify(preprocess_ref.preprocess(np.ascontiguousarray(fr[f:f + 1, y0:y1, x0:x1]), bgr=True))
        err = (out[i].float().cpu() - torch.from_numpy(ref)).abs().max().item()
        assert err < 1.2e-2, (i, err)
    whole = eng.preprocess(frd).view(2, 196, 768)
    assert torch.equal(out[1], whole[1]) and torch.equal(out[6], whole[0])
    unaligned = torch.zeros(fr.size + 1, dtype=torch.uint8, device=dev)       # scalar-load path, same bits
    unaligned[1:] = frd.reshape(-1)
    assert torch.equal(eng.preprocess_rois(unaligned[1:].view(2, 720, 1280, 3), rois).view(len(rois), 196, 768), out)
    with pytest.raises(ValueError):
        eng.preprocess_rois(frd, [(0, 10, 10, 10, 50)])
    with pytest.raises(ValueError):
        eng.preprocess_rois(frd, [(2, 0, 0, 10, 10)])


def test_preprocess_pitched_and_unaligned_input(engine_small):
    """Row pitch > 3*w and a base pointer that is not 16-byte aligned (scalar-load path) give identical results."""
    dev = engine_small.device
    fr = common.noise_frames(2, 300, 500, seed=4)
    base = engine_small.preprocess(torch.from_numpy(fr).to(dev)).clone()
    wide = torch.zeros(2, 300, 512, 3, dtype=torch.uint8, device=dev)
    wide[:, :, :500] = torch.from_numpy(fr).to(dev)
    assert torch.equal(engine_small.preprocess(wide[:, :, :500]), base)
    flat = torch.zeros(2 * 300 * 500 * 3 + 1, dtype=torch.uint8, device=dev)
    flat[1:] = torch.from_numpy(fr).to(dev).reshape(-1)
    assert torch.equal(engine_small.preprocess(flat[1:].view(2, 300, 500, 3)), base)


def test_pool_clips(engine_small):
    dev = engine_small.device
    emb = torch.randn(23, 768, device=dev) * 2 + 0.3
    offs = np.array([0, 1, 1, 9, 23], dtype=np.int32)          # a one-frame clip, an EMPTY clip, ragged lengths
    mean, unit = engine_small.pool_clips(emb, torch.from_numpy(offs))
    ref_mean = reid_ref.clip_mean(emb.cpu().numpy(), np.array([0, 1, 1, 9, 23]))
    ref_mean[1] = 0.0                                          # empty clip -> zeros (np.mean would give NaN)
    np.testing.assert_allclose(mean.cpu().numpy(), ref_mean, atol=2e-6 * 8)
    np.testing.assert_allclose(unit.cpu().numpy(), reid_ref.l2_normalise(ref_mean), atol=1e-6)
    assert torch.isfinite(unit).all()


@pytest.mark.parametrize("q,n,dim", [(3, 1000, 768), (130, 5000, 768), (64, 100000, 768), (5, 3, 768), (9, 700, 1024),
                                     (1, 100000, 768), (2, 5000, 1024), (1, 3, 768), (2, 777, 768), (1, 20011, 1024)])
def test_gallery_topk_bit_exact(engine_small, q, n, dim):
    eng, dev = engine_small, engine_small.device
    gen = torch.Generator(device=dev).manual_seed(q + n)
    g = torch.nn.functional.normalize(torch.randn(n, dim, device=dev, generator=gen), dim=1)
    if n > 10:
        g[n // 2] = g[7]
        g[n - 1] = g[7]                                        # exact duplicate rows -> ties
    gb = g.to(torch.bfloat16).contiguous()
    qv = torch.nn.functional.normalize(torch.randn(q, dim, device=dev, generator=gen), dim=1)
    qv[0] = torch.nn.functional.normalize(g[min(7, n - 1)] + 0.05 * torch.randn(dim, device=dev, generator=gen), dim=0)
    s, i, dump = eng.gallery_topk(qv, gb, k=5, row_base=1000, dump_scores=True)
    ref_scores = reid_ref.cosine_scores(qv.cpu().numpy(), gb.float().cpu().numpy())
    np.testing.assert_allclose(dump.cpu().numpy(), ref_scores, atol=5e-6)      # fp32 query (hi+lo bf16) x bf16 gallery
    kk = min(5, n)
    ref_top, ref_idx = reid_ref.topk_rule(dump.cpu().numpy(), kk, row_base=1000)
    assert (i.cpu().numpy()[:, :kk] == ref_idx).all(), "top-k indices must be bit-exact under (score desc, index asc)"
    assert (s.cpu().numpy()[:, :kk] == ref_top).all()
    if kk < 5:
        assert (i.cpu().numpy()[:, kk:] == 0x7FFFFFFF).all() and np.isneginf(s.cpu().numpy()[:, kk:]).all()
    if n > 10:
        assert i[0].tolist()[:3] == [1007, 1000 + n // 2, 1000 + n - 1]


@pytest.mark.parametrize("q,dim", [(1, 768), (2, 768), (1, 1024), (2, 1024)])
def test_gallery_serving_scan_agrees_with_tile_gemm(engine_small, q, dim):
    """Q <= 2 takes the HBM-streaming scan (fp32 query x bf16 row on the CUDA cores); the tile GEMM (hi + lo bf16 query halves on the
    tensor cores) must pick the same rows, with scores equal to fp32 rounding, including planted ties and every k."""
    eng, dev = engine_small, engine_small.device
    gen = torch.Generator(device=dev).manual_seed(100 + q + dim)
    n = 30011
    g = torch.nn.functional.normalize(torch.randn(n, dim, device=dev, generator=gen), dim=1)
    g[17] = g[29000]
    g[12345] = g[29000]
    gb = g.to(torch.bfloat16).contiguous()
    qv = torch.nn.functional.normalize(g[29000:29000 + q] + 0.02 * torch.randn(q, dim, device=dev, generator=gen), dim=1)
    for k in (1, 5, 8):
        s1, i1, d1 = eng.gallery_topk(qv, gb, k=k, row_base=7, dump_scores=True)
        _lib.set_tuning("scan_small", 0)
        try:
            s0, i0, d0 = eng.gallery_topk(qv, gb, k=k, row_base=7, dump_scores=True)
        finally:
            _lib.set_tuning("scan_small", 1)
        assert (i1.cpu().numpy() == reid_ref.topk_rule(d1.cpu().numpy(), k, row_base=7)[1]).all()
        np.testing.assert_allclose(d1.cpu().numpy(), d0.cpu().numpy(), atol=5e-6)    # the oracle tolerance of both paths
        assert i1[0, 0].item() == 7 + 17                                   # three identical rows: the smallest index leads
        if k >= 5:
            assert i1[0].tolist()[:3] == [7 + 17, 7 + 12345, 7 + 29000] and i0[0].tolist()[:3] == i1[0].tolist()[:3]


def test_gallery_topk_empty_and_k_range(engine_small):
    eng, dev = engine_small, engine_small.device
    qv = torch.nn.functional.normalize(torch.randn(4, 768, device=dev), dim=1)
    s, i = eng.gallery_topk(qv, torch.zeros(0, 768, device=dev, dtype=torch.bfloat16), k=5)
    assert np.isneginf(s.cpu().numpy()).all() and (i.cpu().numpy() == 0x7FFFFFFF).all()
    gb = torch.nn.functional.normalize(torch.randn(300, 768, device=dev), dim=1).to(torch.bfloat16)
    for k in (1, 8):
        s, i, dump = eng.gallery_topk(qv, gb, k=k, dump_scores=True)
        assert (i.cpu().numpy() == reid_ref.topk_rule(dump.cpu().numpy(), k)[1]).all()
    with pytest.raises(_lib.CreError):
        eng.gallery_topk(qv, gb, k=257)                    # > CRE_TOPK_LIMIT


@pytest.mark.parametrize("q,n,k", [(1, 30011, 20), (2, 5000, 64), (33, 10007, 12), (130, 3000, 9), (64, 100000, 17), (3, 40, 256),
                                   (1, 300, 256), (70, 700, 100)])
def test_gallery_topk_beyond_one_pass(engine_small, q, n, k):
    """k > 8 = ceil(k / 8) passes over the shard, each admitting only candidates that rank after the previous pass's last entry
    (include/cre.h): both scan forms, planted exact ties that straddle a pass boundary, k > rows.  Bit-exact against the oracle's
    (score desc, index asc) rule applied to the GPU's own score matrix."""
    eng, dev = engine_small, engine_small.device
    gen = torch.Generator(device=dev).manual_seed(q * 1000 + k)
    g = torch.nn.functional.normalize(torch.randn(n, 768, device=dev, generator=gen), dim=1)
    qv = torch.nn.functional.normalize(torch.randn(q, 768, device=dev, generator=gen), dim=1)
    if n > 30:                                             # 12 identical rows: ties across the first pass boundary (entries 7 | 8)
        dup = torch.randperm(n, device=dev, generator=gen)[:12]
        g[dup] = g[dup[0]].clone()
        qv[0] = torch.nn.functional.normalize(g[dup[0]] + 0.02 * torch.randn(768, device=dev, generator=gen), dim=0)
    gb = g.to(torch.bfloat16).contiguous()
    s, i, dump = eng.gallery_topk(qv, gb, k=k, row_base=50, dump_scores=True)
    kk = min(k, n)
    ref_top, ref_idx = reid_ref.topk_rule(dump.cpu().numpy(), kk, row_base=50)
    assert (i.cpu().numpy()[:, :kk] == ref_idx).all() and (s.cpu().numpy()[:, :kk] == ref_top).all()
    if kk < k:
        assert (i.cpu().numpy()[:, kk:] == 0x7FFFFFFF).all() and np.isneginf(s.cpu().numpy()[:, kk:]).all()
    if n > 30:
        m = min(k, 12)                                     # the tied rows lead, in ascending index, whatever pass they fall in
        assert i[0, :m].tolist() == sorted((dup + 50).tolist())[:m]
    # sharded: per-shard top-k + merge == the whole scan
    from vision_sam3_yolo_lameless_b200.sharded import shard_range
    if 4 * k <= 4096 and n >= 8:
        parts = [eng.gallery_topk(qv, gb[lo:hi].contiguous(), k=k, row_base=50 + lo) for lo, hi in (shard_range(n, r, 4) for r in range(4))]
        ms, mi = eng.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
        assert torch.equal(mi, i) and torch.equal(ms, s)


def test_sharded_topk_merge_equals_whole(engine_small):
    """Row shards scored separately (global indices via row_base) + cre_merge_topk == one scan of the whole gallery."""
    eng, dev = engine_small, engine_small.device
    gen = torch.Generator(device=dev).manual_seed(11)
    n, q = 10007, 33
    g = torch.nn.functional.normalize(torch.randn(n, 768, device=dev, generator=gen), dim=1)
    g[9000] = g[5]
    g[123] = g[5]
    gb = g.to(torch.bfloat16).contiguous()
    qv = torch.nn.functional.normalize(torch.randn(q, 768, device=dev, generator=gen), dim=1)
    qv[0] = g[5]
    whole_s, whole_i = eng.gallery_topk(qv, gb, k=5)
    from vision_sam3_yolo_lameless_b200.sharded import shard_range
    parts = [eng.gallery_topk(qv, gb[lo:hi].contiguous(), k=5, row_base=lo) for lo, hi in (shard_range(n, r, 8) for r in range(8))]
    ms, mi = eng.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    assert torch.equal(mi, whole_i) and torch.equal(ms, whole_s)
    rs, ri = reid_ref.merge_rule(torch.stack([p[0] for p in parts]).cpu().numpy(), torch.stack([p[1] for p in parts]).cpu().numpy(), 5)
    assert (ri == mi.cpu().numpy()).all() and mi[0].tolist()[:3] == [5, 123, 9000]


def test_gallery_update_row(engine_small):
    eng, dev = engine_small, engine_small.device
    gal = torch.nn.functional.normalize(torch.randn(6, 768, device=dev), dim=1).to(torch.bfloat16)
    old = gal[2].float().cpu().numpy().astype(np.float64)
    new = np.random.default_rng(0).standard_normal(768)
    uq = torch.from_numpy(reid_ref.l2_normalise(new).astype(np.float32)).to(dev)
    eng.gallery_update_row(gal, 2, uq, 0.9)
    np.testing.assert_allclose(gal[2].float().cpu().numpy(), reid_ref.momentum_update(old, new, 0.9), atol=1e-3)   # bf16 store: half-ulp 4.9e-4 at |x| in [0.125, 0.25)
    eng.gallery_update_row(gal, 4, uq, 0.0)
    np.testing.assert_allclose(gal[4].float().cpu().numpy(), reid_ref.l2_normalise(new), atol=1e-3)
    # with an fp32 master copy the blend reads and writes IT (what the reference keeps in Qdrant, matcher.py:267-301): 50 updates
    # stay within fp32 rounding of the fp64 recurrence, and the bf16 row is always the rounding of the master row
    master = torch.nn.functional.normalize(torch.randn(6, 768, device=dev), dim=1)
    gal2 = master.to(torch.bfloat16)
    want = master[3].cpu().numpy().astype(np.float64)
    rng = np.random.default_rng(1)
    for _ in range(50):
        new = rng.standard_normal(768)
        eng.gallery_update_row(gal2, 3, torch.from_numpy(reid_ref.l2_normalise(new).astype(np.float32)).to(dev), 0.9, master=master)
        want = reid_ref.momentum_update(want, new, 0.9)
    np.testing.assert_allclose(master[3].cpu().numpy(), want, atol=2e-6)
    assert torch.equal(gal2[3], master[3].to(torch.bfloat16))
    assert torch.equal(gal2[[0, 1, 2, 4, 5]], master[[0, 1, 2, 4, 5]].to(torch.bfloat16)), "other rows untouched"
