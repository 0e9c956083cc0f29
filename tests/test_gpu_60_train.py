"""Fine-tune step on the GPU (BASELINE.json configs[3]): every backward kernel against torch
autograd in fp64, then the whole step through the reference-facing model API against the oracle's autograd --
which tests/test_oracle_golden.py pins to the REAL reference's gradients (tests/golden/grad_*.pt).

Tolerances (stated per north_star / VERDICT): gradient error = ||g - g_ref|| / ||g_ref|| per parameter tensor.
  fp32 mode (CUDA-core kernels):  <= 2e-5   (fp32 accumulation order differs from MKL; 1e-5 typical)
  bf16 mode (tcgen05 kernels):    <= 2e-2   with the oracle forced to the GPU run's kept tokens (a flipped near-tie
                                            token is a different function; selection parity is test_gpu_25)
"""
import math

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

import conftest  # noqa: F401
from conftest import load_golden
from gpu_util import dev, rel_err
from oracle import vit_oracle as vo, weights
from oracle.golden_configs import GRAD_CONFIGS

pytestmark = pytest.mark.gpu


def nerr(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


# ---------------------------------------------------------------- kernels

@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_f32_all_layouts(ta, tb):
    from tpat import ops
    torch.manual_seed(0)
    M, N, K = 200, 77, 131
    A = torch.randn((K, M) if ta else (M, K), device=dev())
    B = torch.randn((N, K) if tb else (K, N), device=dev())
    C0 = torch.randn(M, N, device=dev())
    want = (A.T if ta else A).double() @ (B.T if tb else B).double()
    got = ops.gemm_f32(A, B, ta, tb)
    assert rel_err(got, want) < 1e-5
    acc = ops.gemm_f32(A, B, ta, tb, out=C0.clone(), accumulate=True)
    assert rel_err(acc, want + C0.double()) < 1e-5


def test_transpose_with_padding():
    from tpat import ops
    x = torch.randn(130, 70, device=dev())
    t = ops.transpose(x, torch.float32, ld_dst=192)
    assert t.shape == (70, 192) and torch.equal(t[:, :130], x.T) and bool((t[:, 130:] == 0).all())
    xb = x.to(torch.bfloat16)
    tb = ops.transpose(xb, torch.bfloat16, ld_dst=192)
    assert torch.equal(tb[:, :130], xb.T)
    tc = ops.transpose(x, torch.bfloat16, ld_dst=130)
    assert torch.equal(tc, x.T.to(torch.bfloat16))


@pytest.mark.parametrize("scatter", [False, True])
@pytest.mark.parametrize("op_dtype", [torch.float32, torch.bfloat16])
def test_row_bwd_layernorm_backward_scatter_and_sums(scatter, op_dtype):
    from tpat import ops
    torch.manual_seed(1)
    B, n_in, k, extra, D = 3, 40, 25, 1, 768
    rows_src = extra + (k if scatter else n_in)
    x = torch.randn(B, rows_src, D, device=dev())
    dy = torch.randn(B, rows_src, D, device=dev())
    g_up = torch.randn(B, rows_src, D, device=dev())
    gamma = torch.rand(D, device=dev()) + 0.5
    scale = torch.tensor([0.0, 1.25, 1.25], device=dev())
    idx = torch.stack([torch.randperm(n_in)[:k] for _ in range(B)]).to(dev()) if scatter else None
    # reference by autograd (fp64)
    xd = x.double().requires_grad_(True)
    gd = gamma.double().requires_grad_(True)
    bd = torch.zeros(D, dtype=torch.float64, device=dev(), requires_grad=True)
    y = F.layer_norm(xd, (D,), gd, bd, 1e-6)
    y.backward(dy.double())
    g_src = g_up.double() + xd.grad
    if scatter:
        want = torch.zeros(B, extra + n_in, D, dtype=torch.float64, device=dev())
        want[:, :extra] = g_src[:, :extra]
        want.scatter_(1, (idx + extra).unsqueeze(-1).expand(-1, -1, D), g_src[:, extra:])
    else:
        want = g_src
    g_out, gb, dg, db, dbias = ops.row_bwd(dy, x, gamma, g_up, op_dtype, row_scale=scale, idx=idx, n_in=n_in, extra=extra, eps=1e-6)
    assert rel_err(g_out, want) < 1e-5
    want_b = want * scale.double().view(B, 1, 1)
    assert rel_err(gb.float(), want_b) < (1e-5 if op_dtype == torch.float32 else 6e-3)
    assert rel_err(dg, gd.grad) < 1e-5 and rel_err(db, bd.grad) < 1e-5
    assert rel_err(dbias, want_b.sum(dim=(0, 1))) < 1e-5


def test_colsum_and_batch_sum():
    from tpat import ops
    x = torch.randn(1000, 3072, device=dev())
    assert rel_err(ops.colsum(x), x.double().sum(0)) < 1e-5
    assert rel_err(ops.colsum(x.to(torch.bfloat16)), x.to(torch.bfloat16).double().sum(0)) < 1e-5
    for M, C in ((64 * 178 + 3, 2304), (17, 3072), (4000, 772)):         # 16-byte-load kernel (C % 8 == 0) and the narrow one
        xb = torch.randn(M, C, device=dev()).to(torch.bfloat16)
        assert rel_err(ops.colsum(xb), xb.double().sum(0)) < 1e-5
    y = torch.randn(7, 527, device=dev())
    assert rel_err(ops.batch_sum(y), y.double().sum(0)) < 1e-6


@pytest.mark.parametrize("variant", ["audiomae", "ast"])
def test_pool_norm_bwd(variant):
    from tpat import ops, _lib
    torch.manual_seed(2)
    B, N, D = 4, 30, 768
    x = torch.randn(B, N, D, device=dev())
    g1, b1 = torch.rand(D, device=dev()) + 0.5, torch.randn(D, device=dev()) * 0.1
    g2, b2 = torch.rand(D, device=dev()) + 0.5, torch.randn(D, device=dev()) * 0.1
    dp = torch.randn(B, D, device=dev())
    xd = x.double().requires_grad_(True)
    ps = [t.double().requires_grad_(True) for t in (g1, b1, g2, b2)]
    if variant == "audiomae":
        pooled = F.layer_norm(xd[:, 1:].mean(1), (D,), ps[0], ps[1], 1e-6)
    else:
        t = F.layer_norm(xd, (D,), ps[0], ps[1], 1e-6)
        pooled = F.layer_norm((t[:, 0] + t[:, 1]) / 2, (D,), ps[2], ps[3], 1e-5)
    pooled.backward(dp.double())
    var = _lib.VARIANT_AUDIOMAE if variant == "audiomae" else _lib.VARIANT_AST
    dx, dg1, db1, dg2, db2 = ops.pool_norm_bwd(x, dp, var, g1, b1, 1e-6, g2, 1e-5)
    assert rel_err(dx, xd.grad) < 1e-5
    assert rel_err(dg1, ps[0].grad) < 1e-5 and rel_err(db1, ps[1].grad) < 1e-5
    if variant == "ast":
        assert rel_err(dg2, ps[2].grad) < 1e-5 and rel_err(db2, ps[3].grad) < 1e-5


@pytest.mark.parametrize("impl_name", ["simt", "tc"])
@pytest.mark.parametrize("N", [66, 200, 513])
def test_attention_backward(impl_name, N):
    """dqkv of the fused attention against autograd over the materialised softmax (fp64); the forward's log-sum-exp
    feeds the backward.  bf16 path: operands rounded to bf16 on both sides."""
    from tpat import ops, _lib
    torch.manual_seed(N)
    B, H = 2, 12
    impl = _lib.IMPL_SIMT if impl_name == "simt" else _lib.IMPL_TC
    dt = torch.float32 if impl_name == "simt" else torch.bfloat16
    qkv = (torch.randn(B * N, 3 * H * 64, device=dev()) * 1.2).to(dt)
    d_out = torch.randn(B * N, H * 64, device=dev()).to(dt)
    out, lse = ops.attention_train(qkv, B, N, H, 1, _lib.SCORE_NONE, impl)
    dqkv = ops.attention_bwd(qkv, out, d_out, lse, B, N, H, impl)
    q = qkv.double().requires_grad_(True)
    x = q.reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    attn = ((x[0] @ x[1].transpose(-2, -1)) * 0.125).softmax(-1)
    o = (attn @ x[2]).transpose(1, 2).reshape(B * N, H * 64)
    o.backward(d_out.double())
    lse_ref = torch.logsumexp((x[0] @ x[1].transpose(-2, -1)) * 0.125, dim=-1)
    tol = 2e-5 if impl_name == "simt" else 2e-2
    e_lse, e_o, e_dq = rel_err(lse, lse_ref.detach()), rel_err(out.float(), o.detach()), nerr(dqkv.float(), q.grad)
    print(f"[attention bwd {impl_name}] N={N}: lse err {e_lse:.2e}, out err {e_o:.2e}, dqkv err {e_dq:.2e}")
    assert e_lse < (1e-5 if impl_name == "simt" else 2e-3)
    assert e_dq < tol
    for w in range(3):   # dq, dk, dv separately
        a, b = dqkv.float()[:, w * 768:(w + 1) * 768], q.grad[:, w * 768:(w + 1) * 768]
        assert nerr(a, b) < tol, w
    if impl_name == "tc":
        # fused qkv bias gradient: column sums taken where dqkv is produced, added to what dbias holds; dqkv itself unchanged
        base = torch.randn(3 * H * 64, device=dev())
        dbias = base.clone()
        dqkv2 = ops.attention_bwd(qkv, out, d_out, lse, B, N, H, impl, dbias=dbias)
        assert torch.equal(dqkv2[:, 768:], dqkv[:, 768:])           # dK, dV: bit-identical
        assert nerr(dqkv2[:, :768].float(), dqkv[:, :768].float()) < 3e-3   # dQ: TMA reduce-adds land in any order
        want = dqkv.double().sum(0)
        e_b = nerr(dbias - base, want)
        print(f"[attention bwd tc] N={N}: fused qkv bias gradient err {e_b:.2e} (vs the column sums of the bf16 dqkv)")
        assert e_b < 4e-3            # the q third sums the fp32 values before their rounding to bf16 (~1e-3 expected)
        assert nerr((dbias - base)[768:], want[768:]) < 1e-5
    else:
        with pytest.raises(RuntimeError):
            ops.attention_bwd(qkv, out, d_out, lse, B, N, H, impl, dbias=torch.zeros(3 * H * 64, device=dev()))


@pytest.mark.parametrize("impl_name", ["simt", "tc"])
def test_gemm_train_epilogues(impl_name):
    """fc1 keeping the pre-activation, the GELU-backward epilogue of the data gradient, DropPath's per-clip scale."""
    from tpat import ops, _lib
    torch.manual_seed(3)
    impl = _lib.IMPL_SIMT if impl_name == "simt" else _lib.IMPL_TC
    dt = torch.float32 if impl_name == "simt" else torch.bfloat16
    tol = 1e-5 if impl_name == "simt" else 1e-2
    Bc, n, D, Dh = 3, 50, 768, 3072
    M = Bc * n
    a = torch.randn(M, D, device=dev()).to(dt)
    w1 = (torch.randn(Dh, D, device=dev()) * 0.03).to(dt)
    b1 = torch.randn(Dh, device=dev()) * 0.1
    act, dact = ops.gemm_train(a, w1, b1, dt, _lib.EPI_BIAS_GELU, impl, want_dact=True)
    pre_ref = (a.double() @ w1.double().T + b1.double()).requires_grad_(True)
    act_ref = F.gelu(pre_ref)
    act_ref.sum().backward()                                              # pre_ref.grad = gelu'(pre)
    assert rel_err(act.float(), act_ref.detach()) < tol and rel_err(dact.float(), pre_ref.grad) < tol
    plain = ops.gemm_train(a, w1, b1, dt, _lib.EPI_BIAS_GELU, impl)
    assert torch.equal(plain, act)                                        # keeping the derivative does not change the forward
    # dgrad with the saved derivative: dh = (g W2) * gelu'(pre)
    g = torch.randn(M, D, device=dev()).to(dt)
    w2 = (torch.randn(D, Dh, device=dev()) * 0.03).to(dt)              # fc2.weight [out=D, in=Dh]; B operand = its [in, out] copy
    w2t = w2.T.contiguous()
    dh = ops.gemm_train(g, w2t, None, dt, _lib.EPI_DGELU, impl, aux=dact)
    want_dh = (g.double() @ w2.double()) * pre_ref.grad
    assert nerr(dh.float(), want_dh) < (2e-5 if impl_name == "simt" else 1.5e-2)
    # residual with per-clip scale
    res = torch.randn(M, D, device=dev())
    scale = torch.tensor([0.0, 1.0 / 0.9, 1.0 / 0.9], device=dev())
    hid = torch.randn(M, Dh, device=dev()).to(dt)
    b2 = torch.randn(D, device=dev()) * 0.1
    out = ops.gemm_train(hid, w2, b2, torch.float32, _lib.EPI_BIAS_RESIDUAL, impl, residual=res, row_scale=scale, rows_per_clip=n)
    want = res.double() + scale.double().repeat_interleave(n).view(M, 1) * (hid.double() @ w2.double().T + b2.double())
    assert rel_err(out, want) < tol
    assert torch.equal(out[:n], res[:n])                                # a dropped clip keeps its residual bit for bit


@pytest.mark.parametrize("M,Nout,Nin", [(64 * 178, 3072, 768), (5000, 768, 3072), (3000, 768, 3072), (3000, 3072, 768), (700, 2304, 768),
                                        (300, 768, 768), (130, 768, 320)])
def test_gemm_dgrad_on_forward_weight(M, Nout, Nin):
    """dX = dY W with W = the forward weight [out, in] read as an MN-major tcgen05 B operand (tpat_gemm_extra.w_kn):
    bit-identical to the GEMM on the transposed copy, for the CTA-pair kernel and the 1-CTA kernel's 256 / 128 / 64 tiles."""
    from tpat import ops, _lib
    torch.manual_seed(M)
    dy = (torch.randn(M, Nout, device=dev()) * 0.1).to(torch.bfloat16)
    w = (torch.randn(Nout, Nin, device=dev()) * 0.05).to(torch.bfloat16)
    for out_dtype in (torch.bfloat16, torch.float32):
        got = ops.gemm_train(dy, w, None, out_dtype, _lib.EPI_BIAS, _lib.IMPL_TC, w_kn=True)
        ref = ops.gemm_train(dy, w.T.contiguous(), None, out_dtype, _lib.EPI_BIAS, _lib.IMPL_TC)
        assert torch.equal(got, ref)
        assert nerr(got.float(), dy.double() @ w.double()) < (1e-2 if out_dtype == torch.bfloat16 else 2e-5)
    if Nout == 768 and Nin == 3072:      # GELU-backward epilogue on top
        aux = torch.rand(M, Nin, device=dev()).to(torch.bfloat16)
        got = ops.gemm_train(dy, w, None, torch.bfloat16, _lib.EPI_DGELU, _lib.IMPL_TC, aux=aux, w_kn=True)
        ref = ops.gemm_train(dy, w.T.contiguous(), None, torch.bfloat16, _lib.EPI_DGELU, _lib.IMPL_TC, aux=aux)
        assert torch.equal(got, ref)
    with pytest.raises(RuntimeError):
        ops.gemm_train(dy.float(), w.float(), None, torch.float32, _lib.EPI_BIAS, _lib.IMPL_SIMT, w_kn=True)


@pytest.mark.parametrize("impl_name,M", [("tc", 64 * 178 + 5), ("tc", 3000), ("tc", 100), ("simt", 300)])
def test_gemm_dgelu_fused_bias_gradient(impl_name, M):
    """The GELU-backward GEMM also leaves the column sums of its output (= fc1's bias gradient), accumulated into
    colsum_out; the output itself does not change.  CTA-pair kernel, 1-CTA kernel and the fp32 path (separate pass)."""
    from tpat import ops, _lib
    torch.manual_seed(M)
    impl = _lib.IMPL_SIMT if impl_name == "simt" else _lib.IMPL_TC
    dt = torch.float32 if impl_name == "simt" else torch.bfloat16
    D, Dh = 768, 3072
    g = (torch.randn(M, D, device=dev()) * 0.3).to(dt)
    w2 = (torch.randn(D, Dh, device=dev()) * 0.03).to(dt)
    aux = torch.rand(M, Dh, device=dev()).to(dt)
    base = torch.randn(Dh, device=dev())
    cs = base.clone()
    kw = dict(w_kn=True) if impl_name == "tc" else {}
    wop = w2 if impl_name == "tc" else w2.T.contiguous()
    dh = ops.gemm_train(g, wop, None, dt, _lib.EPI_DGELU, impl, aux=aux, colsum_out=cs, **kw)
    ref = ops.gemm_train(g, wop, None, dt, _lib.EPI_DGELU, impl, aux=aux, **kw)
    assert torch.equal(dh, ref)
    want = (g.double() @ w2.double()) * aux.double()
    e = nerr(cs - base, want.sum(0))
    print(f"[dgelu colsum {impl_name}] M={M}: err {e:.2e}")
    assert e < (1e-5 if impl_name == "simt" else 2e-3)


@pytest.mark.parametrize("K,Mo,No", [(513 * 3, 768, 768), (1000, 2304, 768), (4104, 768, 3072), (70, 256, 256), (32832, 3072, 768)])
def test_gemm_wgrad_tcgen05(K, Mo, No):
    """dW += dY^T X with both operands in their natural layout (MN-major UMMA operands), split-K + TMA reduce-add."""
    from tpat import ops
    torch.manual_seed(K)
    dy = (torch.randn(K, Mo, device=dev()) * 0.1).to(torch.bfloat16)
    x = torch.randn(K, No, device=dev()).to(torch.bfloat16)
    base = torch.randn(Mo, No, device=dev())
    got = ops.gemm_wgrad(dy, x, out=base.clone())
    want = base.double() + dy.double().T @ x.double()
    e = nerr(got, want)
    print(f"[wgrad tcgen05] K={K} {Mo}x{No}: rel err {e:.2e}")
    assert e < 2e-5          # products of bf16 values are exact in fp32; only the fp32 summation order differs


def test_fused_adamw_matches_torch():
    from tpat import ops
    torch.manual_seed(4)
    n = 70000
    p = torch.randn(n, device=dev()); g = torch.randn(n, device=dev())
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([{"params": [ref], "lr": 3e-3, "weight_decay": 0.05}], betas=(0.9, 0.95), eps=1e-8)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    pb = torch.empty(n, device=dev(), dtype=torch.bfloat16)
    chunks = torch.tensor([[0, 16384, 0, 0], [16384, 16384, 0, 0], [32768, n - 32768, 1, 0]], dtype=torch.int32, device=dev())
    groups = torch.tensor([[3e-3, 0.05], [3e-3, 0.05]], device=dev())
    for step in range(1, 4):
        ref.grad = g.clone() * step
        opt.step()
        ops.adamw(p, g * step * 4.0, m, v, pb, chunks, groups, 1.0, 0.9, 0.95, 1e-8, step, grad_scale=0.25)
    assert rel_err(p, ref.detach()) < 1e-6
    assert torch.equal(pb, p.to(torch.bfloat16))


# ---------------------------------------------------------------- the whole step through the model API

def build_train_model(cfg, sd, precision, drop_path_rate=0.1):
    from tpat import models_vit, ASTModel
    if cfg["variant"] == "audiomae":
        m = models_vit.vit_base_patch16(num_classes=cfg["num_classes"], drop_path_rate=drop_path_rate, mean_pooling=True,
                                        mask_2d=True, target_length=cfg["T"], drop_loc=tuple(cfg["drop_loc"]),
                                        base_keep_rate=cfg["base_keep_rate"], precision=precision)
        m.patch_embed = models_vit.PatchEmbed((cfg["T"], 128), 16, 1, 768)
        m.pos_embed = nn.Parameter(torch.zeros(1, m.patch_embed.num_patches + 1, 768), requires_grad=False)
        m.load_state_dict(sd, strict=True)
    else:
        m = ASTModel(label_dim=cfg["num_classes"], input_tdim=cfg["T"], imagenet_pretrain=False, audioset_pretrain=False,
                     verbose=False, drop_loc=tuple(cfg["drop_loc"]), base_keep_rate=cfg["base_keep_rate"], precision=precision)
        m.load_state_dict(sd, strict=False)
    return m.to(dev()).train()


def case_inputs(cfg):
    mk = weights.make_audiomae_state_dict if cfg["variant"] == "audiomae" else weights.make_ast_state_dict
    sd = mk(cfg["num_classes"], cfg["T"], cfg["wseed"], cfg["flavour"])
    x = weights.make_spectrogram(cfg["variant"], cfg["B"], cfg["T"], cfg["xseed"])
    y = (torch.rand(cfg["B"], cfg["num_classes"], generator=torch.Generator().manual_seed(cfg["tseed"])) < 0.1).float()
    return sd, x, y


def oracle_draws(cfg, rates):
    torch.manual_seed(cfg["dseed"])
    noise = None
    if cfg["mask_t_prob"] > 0 or cfg["mask_f_prob"] > 0:
        noise = (torch.rand(cfg["B"], cfg["T"] // 16), torch.rand(cfg["B"], 8))
    return noise, vo.drop_path_scales(rates, cfg["B"])


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", list(GRAD_CONFIGS))
def test_training_step_gradients_match_reference(name, precision):
    g = load_golden("grad_" + name)
    cfg = g["meta"]
    sd, x, y = case_inputs(cfg)
    model = build_train_model(cfg, sd, precision)
    rates = cfg["drop_rates"]
    noise, scales = oracle_draws(cfg, rates)
    model._drop_scales_override = [tuple(None if t is None else t.to(dev()) for t in pair) for pair in scales]
    keep_idx = None
    if noise is not None:
        model._mask_noise_override = noise
        keep_idx = vo.masking_2d_keep_indices(noise[0], noise[1], cfg["mask_t_prob"], cfg["mask_f_prob"])
    kw = dict(mask_t_prob=cfg["mask_t_prob"], mask_f_prob=cfg["mask_f_prob"]) if cfg["variant"] == "audiomae" else {}
    xin = x.to(dev())
    logits = model(xin, keep_rate_list=cfg["keep_rate_list"], **kw)
    assert logits.requires_grad
    loss = F.binary_cross_entropy_with_logits(logits, y.to(dev()))
    loss.backward()
    torch.cuda.synchronize()
    forced = None
    if precision == "bf16":
        forced = {i: t.cpu() for i, t in enumerate(model.last_topk_idx) if t is not None}
    frozen = ("pos_embed",) if cfg["variant"] == "audiomae" else ()
    o_loss, o_logits, o_grads = vo.loss_and_grads(cfg["variant"], sd, x, y, cfg["keep_rate_list"], cfg["drop_loc"],
                                                  cfg["base_keep_rate"], dtype=torch.float64, drop_scales=scales,
                                                  mask_keep_idx=keep_idx, frozen=frozen, forced_idx=forced)
    if precision == "fp32":     # the fp32 run must agree with the REAL reference's numbers as well (golden)
        assert abs(loss.item() - g["loss"]) < 1e-5 * max(1.0, abs(g["loss"]))
        assert rel_err(logits.detach().cpu(), g["logits"]) < 2e-5
    tol = 2e-5 if precision == "fp32" else 2e-2
    named = dict(model.named_parameters())
    worst, worst_name, missing = 0.0, None, []
    for k, ref in o_grads.items():
        p = named[k]
        if p.grad is None:
            missing.append(k)
            continue
        e = nerr(p.grad, ref)
        if e > worst:
            worst, worst_name = e, k
    print(f"[train grads {precision}] {name}: loss {loss.item():.6f} (oracle {o_loss.item():.6f}), {len(o_grads)} tensors, "
          f"worst rel err {worst:.2e} ({worst_name})")
    assert not missing, missing
    unused = [k for k, p in named.items() if p.grad is not None and k not in o_grads]
    assert not unused, unused
    assert worst < tol, (worst_name, worst)
    if precision == "fp32":     # spot-check against the reference's own gradient fingerprints
        for k, summ in g["grads"].items():
            f = named[k].grad.reshape(-1).double().cpu()
            stride = max(1, f.numel() // 64)
            assert abs(f.norm().item() - summ["norm"]) <= 3e-5 * summ["norm"] + 1e-12, k
            assert ((f[::stride][:64] - summ["samples"]).norm() / summ["samples"].norm().clamp_min(1e-30)).item() < 1e-4, k


def test_training_step_at_the_benchmark_shape_bf16():
    """AudioMAE 1024x128, keep 0.7, B = 2 (golden audiomae_1024_b2_kr07's weights and inputs), no DropPath: all 151
    trainable tensors against the fp64 oracle's autograd with the GPU run's kept tokens."""
    cfg = dict(load_golden("audiomae_1024_b2_kr07")["meta"])
    sd, x = conftest.make_case(cfg)
    y = (torch.rand(cfg["B"], cfg["num_classes"], generator=torch.Generator().manual_seed(3)) < 0.05).float()
    worst = {}
    for precision, tol in (("fp32", 2e-5), ("bf16", 2e-2)):
        model = build_train_model(cfg, sd, precision, drop_path_rate=0.0)
        logits = model(x.to(dev()))
        F.binary_cross_entropy_with_logits(logits, y.to(dev())).backward()
        forced = {i: t.cpu() for i, t in enumerate(model.last_topk_idx) if t is not None}
        _, _, o_grads = vo.loss_and_grads("audiomae", sd, x, y, None, cfg["drop_loc"], cfg["base_keep_rate"], dtype=torch.float64,
                                          forced_idx=forced)
        named = dict(model.named_parameters())
        assert len(o_grads) == 151 and all(named[k].grad is not None for k in o_grads)
        errs = {k: nerr(named[k].grad, v) for k, v in o_grads.items()}
        k = max(errs, key=errs.get)
        worst[precision] = (k, errs[k])
        print(f"[train grads {precision}] audiomae_1024_b2_kr07: 151 tensors, worst rel err {errs[k]:.2e} ({k}), median "
              f"{sorted(errs.values())[75]:.2e}")
        assert errs[k] < tol, (k, errs[k])


def test_fused_adamw_training_loop_decreases_loss_and_matches_torch_adamw():
    """Three steps of the fine-tune loop (engine_finetune.py:91-105 + misc.py:259-273 minus the GradScaler) with the
    layer-wise lr decay groups: FusedAdamW on the flat buffers against torch.optim.AdamW on the same groups (fp32 mode)."""
    from tpat.lr_decay import param_groups_lrd
    from tpat.optim import FusedAdamW
    cfg = dict(GRAD_CONFIGS["audiomae_256_b2_train"])
    sd, x, y = case_inputs(cfg)
    runs = {}
    for kind in ("fused", "torch"):
        model = build_train_model(cfg, sd, "fp32", drop_path_rate=0.0)
        groups = param_groups_lrd(model, 0.05, no_weight_decay_list=model.no_weight_decay(), layer_decay=0.75)
        assert len(groups) == 28                    # layer ids 0..13 x {decay, no_decay} (the frozen pos_embed is skipped)
        opt = (FusedAdamW(groups, lr=1e-3, betas=(0.9, 0.95), model=model) if kind == "fused"
               else torch.optim.AdamW(groups, lr=1e-3, betas=(0.9, 0.95)))
        losses = []
        for step in range(3):
            for gq in opt.param_groups:
                gq["lr"] = 1e-3 * gq["lr_scale"]                   # util/lr_sched.py:17-19
            loss = F.binary_cross_entropy_with_logits(model(x.to(dev())), y.to(dev()))
            opt.zero_grad()
            loss.backward()
            opt.step()
            losses.append(loss.item())
        runs[kind] = (losses, {k: v.detach().clone() for k, v in model.named_parameters()})
    print(f"[train loop] losses fused {runs['fused'][0]} torch {runs['torch'][0]}")
    assert runs["fused"][0][-1] < runs["fused"][0][0]
    for a, b in zip(runs["fused"][0], runs["torch"][0]):
        assert abs(a - b) < 2e-5
    for k, v in runs["torch"][1].items():
        assert nerr(runs["fused"][1][k], v) < 1e-4, k


def test_eval_after_fused_adamw_steps_sees_the_new_weights_and_grads_accumulate():
    """(1) FusedAdamW writes the flat parameter buffer from a kernel (no tensor version bump): the inference engine must
    still repack -- an eval forward after training equals a fresh model loaded with the trained state-dict.
    (2) autograd's contract: a second backward without zero_grad ACCUMULATES into p.grad."""
    from tpat.optim import FusedAdamW
    cfg = dict(GRAD_CONFIGS["audiomae_256_b2_train"])
    sd, x, y = case_inputs(cfg)
    model = build_train_model(cfg, sd, "fp32", drop_path_rate=0.0)
    xd, yd = x.to(dev()), y.to(dev())
    with torch.no_grad():
        model.eval()
        before = model(xd).clone()
        model.train()
    opt = FusedAdamW([{"params": [p for p in model.parameters() if p.requires_grad]}], lr=1e-3, betas=(0.9, 0.95), weight_decay=0.0,
                     model=model)
    F.binary_cross_entropy_with_logits(model(xd), yd).backward()
    g1 = model.head.weight.grad.clone()
    F.binary_cross_entropy_with_logits(model(xd), yd).backward()          # no zero_grad: accumulate
    assert nerr(model.head.weight.grad, 2.0 * g1) < 1e-6
    opt.zero_grad()
    assert float(model.head.weight.grad.abs().max()) == 0.0
    for _ in range(2):
        loss = F.binary_cross_entropy_with_logits(model(xd), yd)
        opt.zero_grad(); loss.backward(); opt.step()
    model.eval()
    with torch.no_grad():
        after = model(xd)
        fresh = build_train_model(cfg, {k: v.detach().cpu() for k, v in model.state_dict().items()}, "fp32").eval()
        want = fresh(xd)
    assert not torch.equal(after, before)
    assert torch.equal(after, want)


def test_training_other_sizes_and_fallback_paths():
    """ViT-S (D = 384: the tcgen05 weight-gradient kernel does not cover 384-wide outputs -> transpose route) and a
    batch of one clip, bf16 and fp32, against the oracle's autograd."""
    from tpat import models_vit
    T, C, dim, depth, heads = 128, 12, 384, 12, 6
    sd = weights.make_audiomae_state_dict(C, T, seed=21, flavour="perturbed", depth=depth, dim=dim)
    for B in (1, 3):
        x = weights.make_spectrogram("audiomae", B, T, seed=22 + B)
        y = (torch.rand(B, C, generator=torch.Generator().manual_seed(5)) < 0.2).float()
        for precision, tol in (("fp32", 2e-5), ("bf16", 2.5e-2)):
            m = models_vit.vit_small_patch16(num_classes=C, drop_path_rate=0.0, mean_pooling=True, mask_2d=True, target_length=T,
                                             drop_loc=(2, 5), base_keep_rate=0.6, precision=precision)
            m.patch_embed = models_vit.PatchEmbed((T, 128), 16, 1, dim)
            m.pos_embed = nn.Parameter(torch.zeros(1, m.patch_embed.num_patches + 1, dim), requires_grad=False)
            m.load_state_dict(sd, strict=True)
            m = m.to(dev()).train()
            F.binary_cross_entropy_with_logits(m(x.to(dev())), y.to(dev())).backward()
            forced = {i: t.cpu() for i, t in enumerate(m.last_topk_idx) if t is not None}
            _, _, o_grads = vo.loss_and_grads("audiomae", sd, x, y, None, (2, 5), 0.6, num_heads=heads, dtype=torch.float64,
                                              forced_idx=forced if precision == "bf16" else None)
            named = dict(m.named_parameters())
            errs = {k: nerr(named[k].grad, v) for k, v in o_grads.items()}
            k = max(errs, key=errs.get)
            print(f"[train grads {precision}] vit_small B={B}: worst rel err {errs[k]:.2e} ({k})")
            assert errs[k] < tol, (k, errs[k])


def test_graphed_train_step_matches_eager_steps():
    """GraphedTrainStep: the captured step (forward + loss + backward + FusedAdamW, device-side step counter) replays to the
    same losses and weights as the eager loop (fp32 mode, no DropPath -> deterministic)."""
    from tpat.lr_decay import param_groups_lrd
    from tpat.optim import FusedAdamW
    from tpat.train import GraphedTrainStep
    cfg = dict(GRAD_CONFIGS["audiomae_256_b2_train"])
    sd, x, y = case_inputs(cfg)
    xs = [x.to(dev()), (x * 0.9 + 0.05).to(dev()), (x * 1.1).to(dev())]
    yd = y.to(dev())
    crit = lambda lg, t: F.binary_cross_entropy_with_logits(lg, t)
    runs = {}
    for kind in ("eager", "graph"):
        model = build_train_model(cfg, sd, "fp32", drop_path_rate=0.0)
        groups = param_groups_lrd(model, 0.05, no_weight_decay_list=model.no_weight_decay(), layer_decay=0.75)
        opt = FusedAdamW(groups, lr=1e-3, betas=(0.9, 0.95), model=model)
        for gq in opt.param_groups:
            gq["lr"] = 1e-3 * gq["lr_scale"]
        losses = []
        if kind == "eager":
            for i in range(3 + 3):                       # the graph run spends 3 eager warm-up steps on xs[0]; the capture
                xi = xs[0] if i < 3 else xs[(i - 3) % 3]  # pass itself executes nothing
                loss = crit(model(xi), yd)
                opt.zero_grad(); loss.backward(); opt.step()
                losses.append(loss.item())
            losses = losses[3:]
        else:
            step = GraphedTrainStep(model, opt, crit, xs[0], yd, warmup=3)
            for i in range(3):
                losses.append(step(xs[i % 3], yd).item())
            assert opt.state_dict()["step"] == 6     # device-side counter: 3 warm-up steps + 3 replays
        runs[kind] = (losses, {k: v.detach().clone() for k, v in model.named_parameters()})
    print(f"[graphed step] losses eager {runs['eager'][0]} graph {runs['graph'][0]}")
    for a, b in zip(runs["eager"][0], runs["graph"][0]):
        assert abs(a - b) < 1e-5
    for k, v in runs["eager"][1].items():
        assert nerr(runs["graph"][1][k], v) < 1e-5, k
