"""Per-operation references for the kernel-level parity tests (tests/test_ops_gpu.py).

Every function restates ONE module of the reference network (rightLaneNetwork/models/FCDenseNet/layers.py) or its
autograd derivative with plain torch ops in float64, on inputs that are exactly representable in bf16, and mirrors the
storage roundings of the CUDA pipeline (activations / activation gradients are stored as bf16 between kernels, the
tensor-core operands are bf16) so that what is left between reference and kernel is fp32 summation order: the tests can
then hold every kernel to ~1e-3 instead of the network-level bf16 noise envelope.

Channel order is the PHYSICAL arena order throughout ([skip | up 80 | new 80], csrc/rl_plan.h); `ref_channel` maps a
physical input channel to the reference's index for parameters of up-path layers.
"""
import torch
import torch.nn.functional as F

EPS = 1e-5


def bf16r(t):
    """round to bf16 and back (keeps dtype)"""
    return t.to(torch.bfloat16).to(t.dtype)


def rand_bf16(*shape, scale=1.0, gen=None, device="cuda"):
    return bf16r(torch.randn(*shape, generator=gen, device=device, dtype=torch.float32) * scale)


def ref_channel(c, cs_perm):
    if cs_perm <= 0:
        return c
    if c < cs_perm:
        return c + 80
    if c < cs_perm + 80:
        return c - cs_perm
    return c


def perm_index(cin, cs_perm, device="cuda"):
    """idx[c_phys] = reference channel"""
    return torch.tensor([ref_channel(c, cs_perm) for c in range(cin)], device=device, dtype=torch.long)


def channel_sums(x_nhwc):
    """Σx, Σx² per channel in float64 (what the producers accumulate into the level statistics)"""
    x = x_nhwc.double().reshape(-1, x_nhwc.shape[-1])
    return torch.stack([x.sum(0), (x * x).sum(0)])


def bn_coef(sums, n, gamma, beta):
    """csrc/rl_dev.cuh bn_coef: fp64 mean / variance, fp32 rsqrt and coefficients.  Returns float32 (s, t, mean, rstd)"""
    m = sums[0] / n
    v = (sums[1] / n - m * m).clamp(min=0)
    mean = m.float()
    rstd = torch.rsqrt(v.float() + EPS)
    s = gamma.float() * rstd
    t = beta.float() - mean * s
    return s, t, mean, rstd


def pre_act(x_nhwc, s, t):
    """fp32 fma(x, s, t) exactly as the kernels evaluate the BatchNorm + ReLU decision"""
    return torch.addcmul(t.float(), x_nhwc.float(), s.float())


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


# ---------------------------------------------------------------------------------------------- forward modules
def dense_layer_fwd(x, sums, n, gamma, beta, w, bias, keep_scale):
    """DenseLayer (layers.py:5-15): BN(batch stats) -> ReLU -> Conv3x3(cin->16) + bias -> Dropout2d scale.
    x [B,H,W,cin] bf16-exact, w [16,cin,3,3] bf16-exact, keep_scale [B,16].  Returns the fp64 result [B,H,W,16]."""
    s, t, _, _ = bn_coef(sums, n, gamma, beta)
    a = bf16r(torch.relu(pre_act(x, s, t))).double()
    y = F.conv2d(nchw(a), w.double(), bias.double(), padding=1)
    return nhwc(y * keep_scale.double()[:, :, None, None])


def transition_down_fwd(x, sums, n, gamma, beta, w, bias, keep_scale):
    """TransitionDown (layers.py:43-55): BN -> ReLU -> Conv1x1 + bias -> Dropout2d scale -> MaxPool2d(2) (floor).
    Returns (pooled fp64 [B,H/2,W/2,C], argmax position 0..3 inside the window, gap between the two largest candidates,
    magnitude of the largest)."""
    s, t, _, _ = bn_coef(sums, n, gamma, beta)
    a = bf16r(torch.relu(pre_act(x, s, t))).double()
    z = torch.einsum("bhwc,oc->bhwo", a, w.double()[:, :, 0, 0])
    B, H, W, C = z.shape
    Hp, Wp = H // 2, W // 2
    win = z[:, :2 * Hp, :2 * Wp].reshape(B, Hp, 2, Wp, 2, C).permute(0, 1, 3, 5, 2, 4).reshape(B, Hp, Wp, C, 4)
    top2 = win.topk(2, dim=-1).values
    mx, am = win.max(-1)
    pooled = (mx + bias.double()) * keep_scale.double()[:, None, None, :]
    return pooled, am, (top2[..., 0] - top2[..., 1]), top2[..., 0].abs()


def transition_up_fwd(x, w, bias, Hout, Wout):
    """TransitionUp (layers.py:58-69, 82-86): ConvTranspose2d(80->80, k3, s2, p0) + bias, center_crop (offset 0).
    The parity kernel accumulates every tap in fp32 and rounds once.  x [B,h,w,80], w [80(in),80(out),3,3].
    Returns fp64 [B,Hout,Wout,80]."""
    out = F.conv_transpose2d(nchw(x.double()), w.double(), bias.double(), stride=2)
    return nhwc(out[:, :, :Hout, :Wout])


# ---------------------------------------------------------------------------------------------- dense block backward
def conv_w_grad(a_nhwc, dz_nhwc):
    """dW[co, c, ky, kx] = Σ_p a[p + (ky-1, kx-1), c] · dz[p, co]  (3x3, padding 1)"""
    return torch.nn.grad.conv2d_weight(nchw(a_nhwc.double()), (dz_nhwc.shape[-1], a_nhwc.shape[-1], 3, 3), nchw(dz_nhwc.double()),
                                       padding=1)


def dense_block_bwd(x, dX, sums, n, acc, layers, C0, rmw_from_last, fused):
    """Backward of one DenseBlock (five DenseLayers, layers.py:18-40) as the CUDA schedule runs it.

    x    [B,H,W,Ct] activations (bf16-exact); dX [B,H,W,Ct] gradient arena before the block (bf16-exact)
    sums [2,Ct] float64 channel statistics; acc [2,Ct] float64 deferred BatchNorm-backward terms (A, B) before the block
    layers[j]: dict(gamma, beta [cin_j] physical order, w [16,cin_j,3,3] bf16-exact physical order, keep [B,16])
    fused: block-fused schedule (chain over the new channels + one accumulation into the block input, dγ/dβ/dW from the
           m·x̂ / m operand pair) or the per-layer schedule.
    Returns dict(dX, acc, per-layer dgamma, dbeta, dw, dbias, dz)."""
    dX = dX.double().clone()
    acc = acc.clone()
    xd = x.double()
    out = {"layers": [None] * 5}
    delta = torch.zeros_like(dX[..., :C0])         # fused: total added to the block-input gradient
    for j in range(4, -1, -1):
        L = layers[j]
        cin = C0 + 16 * j
        o = cin
        s, t, mean, rstd = bn_coef(sums[:, :cin], n, L["gamma"], L["beta"])
        _, _, mean_o, rstd_o = bn_coef(sums[:, o:o + 16], n, torch.ones(16, device=x.device), torch.zeros(16, device=x.device))
        # ---- dz of the layer's own output channels (dz_dense16_kernel): fp32 constants as the kernel forms them
        A = (acc[0, o:o + 16] / n).float()
        Bc = (acc[1, o:o + 16] / n).float()
        b2 = rstd_o * Bc
        a2 = A - mean_o * b2
        dzf = L["keep"].float()[:, None, None, :] * (dX[..., o:o + 16].float() - torch.addcmul(a2, xd[..., o:o + 16].float(), b2))
        dz = bf16r(dzf).double()
        dbias = dzf.double().sum((0, 1, 2))
        # ---- input gradient through the convolution and the ReLU
        G = nhwc(F.conv_transpose2d(nchw(dz), L["w"].double(), padding=1))      # [B,H,W,cin]
        z = pre_act(x[..., :cin], s, t)
        m = (z > 0)
        xh32 = torch.addcmul(-mean * rstd, x[..., :cin].float(), rstd)                   # fma(x, rstd, -mean·rstd)
        g = torch.where(m, G, torch.zeros_like(G))
        k = (L["gamma"].float() * rstd).double()                                       # γ·rstd
        contrib = g * k
        if fused:
            # operand pair of wgrad_dense2: m·(x - c) with c = bf16(mean) (rounded to bf16), and m;  x̂ = r·(x - c) - r·(mean - c)
            cc = bf16r(mean)
            d_op = bf16r(x[..., :cin].float() - cc).double()
            m_op = m.double()
            a_op = m_op * (rstd.double() * d_op - (rstd * (mean - cc)).double())
            dgamma = (G * a_op).sum((0, 1, 2))
            dbeta = (G * m_op).sum((0, 1, 2))
            dw = conv_w_grad(a_op, dz) * L["gamma"].double()[None, :, None, None] + \
                conv_w_grad(m_op, dz) * L["beta"].double()[None, :, None, None]
            # chain channels [C0, cin): per-layer read-modify-write with the kernel's own pixel sums
            if j > 0:
                xm = xd[..., C0:cin] - mean[C0:].double()
                S1 = g[..., C0:].sum((0, 1, 2))
                S2 = (g[..., C0:] * xm).sum((0, 1, 2)) * rstd[C0:].double()
                acc[0, C0:cin] += k[C0:] * S1
                acc[1, C0:cin] += k[C0:] * S2
                dX[..., C0:cin] = bf16r((dX[..., C0:cin] + contrib[..., C0:]).float()).double()
            delta += contrib[..., :C0]
        else:
            a_op = bf16r(torch.relu(z)).double()
            dw = conv_w_grad(a_op, dz)
            xm = xd[..., :cin] - mean.double()
            S1 = g.sum((0, 1, 2))
            S2 = (g * xm).sum((0, 1, 2)) * rstd.double()
            dbeta, dgamma = S1, S2
            acc[0, :cin] += k * S1
            acc[1, :cin] += k * S2
            rf = rmw_from_last if j == 4 else 0
            base = dX[..., :cin].clone()
            base[..., :rf] = 0
            dX[..., :cin] = bf16r((base + contrib).float()).double()
        out["layers"][j] = dict(dgamma=dgamma, dbeta=dbeta, dw=dw, dbias=dbias, dz=dz)
    if fused:
        for j in range(5):   # Σ_p Δ = Σ_j γ_j r·dβ_j and Σ_p Δ·x̂ = Σ_j γ_j r·dγ_j: added by the weight-gradient kernel
            kj = (layers[j]["gamma"].float()[:C0] * rstd[:C0]).double()
            acc[0, :C0] += kj * out["layers"][j]["dbeta"][:C0]
            acc[1, :C0] += kj * out["layers"][j]["dgamma"][:C0]
        base = dX[..., :C0] if rmw_from_last == 0 else torch.zeros_like(delta)
        dX[..., :C0] = bf16r((base + delta).float()).double()
    out["dX"], out["acc"] = dX, acc
    return out


# ---------------------------------------------------------------------------------------------- comparison helpers
def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-300))


def bf16_ulps(ours_bf16, ref_f64):
    """distance of a bf16 result from the (unrounded) reference in units of the bf16 spacing at the reference's magnitude"""
    ref = ref_f64.double()
    spacing = torch.pow(2.0, torch.floor(torch.log2(ref.abs().clamp(min=1e-30))) - 7)
    return ((ours_bf16.double() - ref).abs() / spacing)
