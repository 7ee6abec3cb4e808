"""Kernel-level parity through the C ABI (rl_forward_range / rl_backward_range / rl_head / optimizers): ONE module of the
reference network at a time, on bf16-representable random inputs written straight into the plan's arenas, against the
float64 restatements in tests/opref.py (reference rightLaneNetwork/models/FCDenseNet/layers.py:5-69 and its autograd).

Stated tolerances (north_star: 1e-3 relative in fp32, a stated looser bound for bf16 storage):
  * fp32 outputs (weight / BatchNorm / bias gradients, pixel sums, loss, optimizer state): <= 1e-3 norm-wise
    (optimizers: 1e-6 after 10 steps);
  * bf16-stored outputs (activations, activation gradients): the reference is evaluated in float64 and the kernel's
    stored value must lie within 1 bf16 spacing of it for all but a vanishing fraction of elements (an exactly rounded
    result is within 0.5; fp32 accumulation order moves values across a rounding boundary now and then), and the
    norm-wise error against the bf16-ROUNDED reference must be <= 2e-3.
"""
import pytest
import torch

from sim2real_lane_segment_b200 import engine as E
from sim2real_lane_segment_b200.engine import Engine
from tests import opref as R

pytestmark = pytest.mark.gpu
DEV = "cuda"


# ------------------------------------------------------------------------------------------------ harness
class Net:
    """a bound training plan plus a random, bf16-exact parameter arena"""

    def __init__(self, B, H, W, C=4, seed=0, options=None):
        self.eng = Engine(B, H, W, C, training=True).bind()
        for k, v in (options or {}).items():
            self.eng.set_option(k, v)
        self.B, self.H, self.W, self.C = B, H, W, C
        self.gen = torch.Generator(device=DEV).manual_seed(seed)
        self.params = torch.zeros(self.eng.n_params, device=DEV)
        self.bn = torch.zeros(self.eng.n_bnbuf, device=DEV)
        self.nlayers = self.eng.query(E.Q_NUM_LAYERS)
        self.layers = [self._layer(i) for i in range(self.nlayers)]

    def q(self, what, i=0):
        return self.eng.query(what, i)

    def _layer(self, i):
        d = dict(type=self.q(E.Q_LAYER_TYPE, i), lin=self.q(E.Q_LAYER_LEVEL_IN, i), lout=self.q(E.Q_LAYER_LEVEL_OUT, i),
                 cin=self.q(E.Q_LAYER_CIN, i), cout=self.q(E.Q_LAYER_COUT, i), in_off=self.q(E.Q_LAYER_IN_OFF, i),
                 out_off=self.q(E.Q_LAYER_OUT_OFF, i), cs_perm=self.q(E.Q_LAYER_CS_PERM, i), p_w=self.q(E.Q_LAYER_PARAM_W, i),
                 p_gamma=self.q(E.Q_LAYER_PARAM_GAMMA, i), p_b=self.q(E.Q_LAYER_PARAM_B, i), rmw_from=self.q(E.Q_LAYER_RMW_FROM, i))
        return d

    def level(self, l):
        return dict(H=self.q(E.Q_H, l), W=self.q(E.Q_W, l), Ct=self.q(E.Q_CT, l))

    def rand(self, *shape, scale=1.0):
        return R.rand_bf16(*shape, scale=scale, gen=self.gen)

    def set_bn_conv(self, i, k):
        """random BatchNorm weight/bias + bf16-exact conv weight/bias for layer i (k = kernel size); returns them in
        PHYSICAL input-channel order"""
        L = self.layers[i]
        cin, cout = L["cin"], L["cout"]
        gamma = 1.0 + 0.3 * torch.randn(cin, generator=self.gen, device=DEV)
        gamma[::7] *= -1            # negative BatchNorm weights flip the ReLU decision's direction
        beta = 0.3 * torch.randn(cin, generator=self.gen, device=DEV)
        w = self.rand(cout, cin, k, k, scale=(2.0 / (cin * k * k)) ** 0.5)
        bias = 0.1 * torch.randn(cout, generator=self.gen, device=DEV)
        idx = R.perm_index(cin, L["cs_perm"])
        pg = L["p_gamma"]
        self.params[pg + idx] = gamma
        self.params[pg + cin + idx] = beta
        wr = torch.empty_like(w)
        wr[:, idx] = w
        self.params[L["p_w"]:L["p_w"] + w.numel()] = wr.reshape(-1)
        self.params[L["p_b"]:L["p_b"] + cout] = bias
        return dict(gamma=gamma, beta=beta, w=w, bias=bias, idx=idx)

    def set_input_level(self, l, scale=1.0, offset=0.3):
        """random activations for a whole level arena + consistent channel statistics"""
        g = self.level(l)
        x = R.bf16r(torch.randn(self.B, g["H"], g["W"], g["Ct"], generator=self.gen, device=DEV) * scale + offset)
        self.eng.set_level(l, x, "act")
        sums = R.channel_sums(x)
        self.eng.level_stats(l).copy_(sums)
        return x, sums

    def keep(self, i, p=0.2):
        L = self.layers[i]
        m = (torch.rand(self.B, L["cout"], generator=self.gen, device=DEV) >= p).float() / (1 - p)
        self.eng.layer_mask_view(i).copy_(m)
        return m


def check_bf16(ours, ref64, what, frac_ok=0.999, norm_tol=2e-3):
    u = R.bf16_ulps(ours, ref64)
    # elements far below the tensor's scale (saturated softmax gradients ~1e-20 ...) are held to an absolute bound instead
    floor = 1e-6 * float(ref64.abs().max())
    u = torch.where((ours.double() - ref64.double()).abs() <= floor, torch.zeros_like(u), u)
    within = float((u <= 1.0).double().mean())
    e = R.rel(ours, R.bf16r(ref64.float()))
    print(f"{what}: {100 * within:.4f}% within 1 bf16 spacing, max {float(u.max()):.2f}, norm rel err vs rounded ref {e:.2e}")
    assert within >= frac_ok, (what, within)
    assert e <= norm_tol, (what, e)


def check_f32(ours, ref64, what, tol=1e-3):
    e = R.rel(ours, ref64)
    print(f"{what}: rel err {e:.2e}")
    assert e <= tol, (what, e)


def dense_layer_indices(net, level, up):
    out = [i for i, L in enumerate(net.layers) if L["type"] == 1 and L["lin"] == level and (L["cs_perm"] > 0) == up]
    return out


# ------------------------------------------------------------------------------------------------ forward modules
@pytest.mark.parametrize("B,H,W,level,up,j", [
    (2, 120, 160, 0, False, 0), (2, 120, 160, 0, True, 4), (2, 120, 160, 1, True, 2), (3, 120, 160, 3, False, 3),
    (4, 120, 160, 4, True, 1), (4, 120, 160, 5, False, 4), (2, 64, 352, 0, True, 0), (2, 40, 56, 2, False, 2)],
    ids=["l0-down0", "l0-up4", "l1-up2", "l3-down3", "l4-up1-7x10", "bottleneck4-3x5", "strips352", "odd-40x56-l2"])
def test_dense_layer_forward(B, H, W, level, up, j):
    net = Net(B, H, W, seed=level * 10 + j)
    li = dense_layer_indices(net, level, up)[j]
    L = net.layers[li]
    x, sums = net.set_input_level(level)
    P = net.set_bn_conv(li, 3)
    keep = net.keep(li)
    net.eng.pack_weights(net.params)
    cin, o = L["cin"], L["out_off"]
    assert o == cin
    stats = net.eng.level_stats(level)
    stats[:, o:o + 16] = 0
    n = B * x.shape[1] * x.shape[2]
    net.eng.forward_range(net.params, net.bn, None, True, li, li)
    torch.cuda.synchronize()
    ref = R.dense_layer_fwd(x[..., :cin], sums[:, :cin], n, P["gamma"], P["beta"], P["w"], P["bias"], keep)
    got = net.eng.level_view(level)[..., o:o + 16]
    check_bf16(got, ref, f"DenseLayer cin={cin} level {level}")
    # the planar mirror holds the same values, the untouched channels are untouched
    xp = net.eng.level_planar(level, "xp")[o // 16].reshape(B, x.shape[1], x.shape[2], 16)
    assert torch.equal(xp, got)
    assert torch.equal(net.eng.level_view(level)[..., :cin].float(), x[..., :cin])
    check_f32(stats[:, o:o + 16], R.channel_sums(got), "Σx, Σx² of the new channels", tol=1e-5)
    # eval mode: running statistics, no dropout
    rm = torch.randn(cin, device=DEV) * 0.2 + 0.3
    rv = torch.rand(cin, device=DEV) + 0.5
    b_rm = net.q(E.Q_LAYER_BN_RM, li)
    net.bn[b_rm + P["idx"]] = rm
    net.bn[b_rm + cin + P["idx"]] = rv
    net.eng.forward_range(net.params, net.bn, None, False, li, li)
    torch.cuda.synchronize()
    fake = torch.stack([rm.double() * n, (rv.double() + rm.double() ** 2) * n])   # sums whose mean / variance are rm / rv
    ref = R.dense_layer_fwd(x[..., :cin], fake, n, P["gamma"], P["beta"], P["w"], P["bias"], torch.ones_like(keep))
    check_bf16(net.eng.level_view(level)[..., o:o + 16], ref, "eval mode", frac_ok=0.995)


@pytest.mark.parametrize("B,H,W,level", [(2, 120, 160, 0), (2, 120, 160, 2), (3, 120, 160, 3), (4, 120, 160, 4), (2, 40, 56, 1)],
                         ids=["l0-128ch", "l2-288ch", "l3-368ch-15x20", "l4-448ch-7x10", "odd-40x56-l1"])
def test_transition_down_forward(B, H, W, level):
    net = Net(B, H, W, seed=100 + level)
    li = [i for i, L in enumerate(net.layers) if L["type"] == 2 and L["lin"] == level][0]
    L = net.layers[li]
    x, sums = net.set_input_level(level)
    P = net.set_bn_conv(li, 1)
    keep = net.keep(li)
    net.eng.pack_weights(net.params)
    C = L["cin"]
    n = B * x.shape[1] * x.shape[2]
    stats_out = net.eng.level_stats(level + 1)
    stats_out.zero_()
    net.eng.forward_range(net.params, net.bn, None, True, li, li)
    torch.cuda.synchronize()
    pooled, am, gap, top = R.transition_down_fwd(x[..., :C], sums[:, :C], n, P["gamma"], P["beta"], P["w"], P["bias"], keep)
    got = net.eng.level_view(level + 1)[..., :C]
    check_bf16(got, pooled, f"TransitionDown C={C}")
    check_f32(stats_out[:, :C], R.channel_sums(got), "Σx, Σx² of the pooled channels", tol=1e-5)
    Hp, Wp = pooled.shape[1], pooled.shape[2]
    off = net.q(E.Q_LAYER_AMAX_OFF, li)
    amax = net.eng._ws_view[off:off + B * Hp * Wp * C].view(B, Hp, Wp, C).long()
    # the arg-max may differ only where the two best candidates of a window are (nearly) tied
    bad = (amax != am) & (gap > 1e-5 * top.clamp(min=1e-3))
    assert int(bad.sum()) == 0, int(bad.sum())


@pytest.mark.parametrize("B,H,W,lin", [(2, 120, 160, 1), (3, 120, 160, 3), (4, 120, 160, 5), (2, 40, 56, 2)],
                         ids=["60x80->120x160", "15x20->30x40", "3x5->7x10", "odd-10x14->20x28"])
def test_transition_up_forward(B, H, W, lin):
    net = Net(B, H, W, seed=200 + lin)
    li = [i for i, L in enumerate(net.layers) if L["type"] == 3 and L["lin"] == lin][0]
    L = net.layers[li]
    x, _ = net.set_input_level(lin)
    w = net.rand(80, 80, 3, 3, scale=0.05)
    bias = 0.1 * torch.randn(80, generator=net.gen, device=DEV)
    net.params[L["p_w"]:L["p_w"] + w.numel()] = w.reshape(-1)
    net.params[L["p_b"]:L["p_b"] + 80] = bias
    net.eng.pack_weights(net.params)
    go = net.level(L["lout"])
    stats_out = net.eng.level_stats(L["lout"])
    stats_out.zero_()
    net.eng.forward_range(net.params, net.bn, None, True, li, li)
    torch.cuda.synchronize()
    ref = R.transition_up_fwd(x[..., L["in_off"]:L["in_off"] + 80], w, bias, go["H"], go["W"])
    o = L["out_off"]
    got = net.eng.level_view(L["lout"])[..., o:o + 80]
    check_bf16(got, ref, "TransitionUp", frac_ok=0.995)
    check_f32(stats_out[:, o:o + 80], R.channel_sums(got), "Σx, Σx² of the up-sampled channels", tol=1e-5)


# ------------------------------------------------------------------------------------------------ dense block backward
def _block_case(net, level, up, fused):
    idxs = dense_layer_indices(net, level, up)
    assert len(idxs) == 5
    g = net.level(level)
    B, H, W, Ct = net.B, g["H"], g["W"], g["Ct"]
    n = B * H * W
    x, sums = net.set_input_level(level)
    layers = []
    for i in idxs:
        P = net.set_bn_conv(i, 3)
        P["keep"] = net.keep(i)
        layers.append(P)
    net.eng.pack_weights(net.params)
    dX = net.rand(B, H, W, Ct, scale=1e-2)
    net.eng.set_level(level, dX, "dact")
    acc = torch.randn(2, Ct, generator=net.gen, device=DEV, dtype=torch.float64) * 1e-2 * n ** 0.5
    net.eng.level_acc(level).copy_(acc)
    net.eng.gstage().zero_()
    C0 = net.layers[idxs[0]]["cin"]
    rf = net.layers[idxs[4]]["rmw_from"]
    net.eng.backward_range(net.params, idxs[4], idxs[0])
    torch.cuda.synchronize()
    ref = R.dense_block_bwd(x, dX, sums, n, acc, layers, C0, rf, fused)
    Cend = C0 + 80
    got_dX = net.eng.level_view(level, grad=True)
    # channels beyond the block (Cend..Ct) must be untouched
    assert torch.equal(got_dX[..., Cend:].float(), dX[..., Cend:])
    check_bf16(got_dX[..., :Cend], ref["dX"][..., :Cend], f"dX level {level} C0={C0} fused={fused}", frac_ok=0.995)
    got_acc = net.eng.level_acc(level)
    scale = float(ref["acc"][:, :Cend].abs().mean())
    err = float((got_acc[:, :Cend] - ref["acc"][:, :Cend]).abs().max()) / scale
    print(f"acc A/B: max abs err / mean magnitude {err:.2e}")
    assert err < 2e-3
    gs = net.eng.gstage()
    for j, i in enumerate(idxs):
        L, P, Rj = net.layers[i], layers[j], ref["layers"][j]
        cin = L["cin"]
        idx = P["idx"]
        dg = gs[L["p_gamma"]:L["p_gamma"] + cin][idx]
        db = gs[L["p_gamma"] + cin:L["p_gamma"] + 2 * cin][idx]
        dw = gs[L["p_w"]:L["p_w"] + 16 * cin * 9].view(16, cin, 3, 3)[:, idx]
        dbias = gs[L["p_b"]:L["p_b"] + 16]
        check_f32(dw, Rj["dw"], f"layer {j} dW")
        check_f32(dg, Rj["dgamma"], f"layer {j} dgamma", tol=2e-3)
        check_f32(db, Rj["dbeta"], f"layer {j} dbeta", tol=2e-3)
        check_f32(dbias, Rj["dbias"], f"layer {j} dbias")


BLOCKS = [(2, 120, 160, 0, False), (2, 120, 160, 0, True), (2, 120, 160, 1, True), (3, 120, 160, 2, False),
          (4, 120, 160, 3, True), (4, 120, 160, 4, False), (4, 120, 160, 5, False), (2, 40, 56, 1, True)]
BLOCK_IDS = ["l0-down-C48", "l0-up-C208", "l1-up-C288", "l2-down-C208", "l3-up-15x20", "l4-down-7x10", "bottleneck-3x5",
             "odd-20x28-up"]


@pytest.mark.parametrize("B,H,W,level,up", BLOCKS, ids=BLOCK_IDS)
def test_dense_block_backward_fused(B, H, W, level, up):
    net = Net(B, H, W, seed=300 + level + (7 if up else 0), options={"bwd_fused": 1})
    _block_case(net, level, up, True)


@pytest.mark.parametrize("B,H,W,level,up", BLOCKS, ids=BLOCK_IDS)
def test_dense_block_backward_per_layer(B, H, W, level, up):
    net = Net(B, H, W, seed=300 + level + (7 if up else 0), options={"bwd_fused": 0})
    _block_case(net, level, up, False)


def test_dense_block_backward_is_deterministic_enough():
    """two runs of the same block: fp32 atomics make the weight gradients order-dependent, nothing else"""
    outs = []
    for _ in range(2):
        net = Net(2, 120, 160, seed=5)
        idxs = dense_layer_indices(net, 1, True)
        x, sums = net.set_input_level(1)
        for i in idxs:
            net.set_bn_conv(i, 3)
            net.keep(i)
        net.eng.pack_weights(net.params)
        g = net.level(1)
        net.eng.set_level(1, net.rand(2, g["H"], g["W"], g["Ct"], scale=1e-2), "dact")
        net.eng.level_acc(1).zero_()
        net.eng.gstage().zero_()
        net.eng.backward_range(net.params, idxs[4], idxs[0])
        torch.cuda.synchronize()
        outs.append((net.eng.level_view(1, grad=True).clone(), net.eng.gstage().clone()))
    frac = float((outs[0][0] != outs[1][0]).float().mean())
    print(f"dX elements that differ between two runs: {100 * frac:.4f}%; gstage rel diff {R.rel(outs[0][1], outs[1][1]):.2e}")
    assert frac < 1e-3 and R.rel(outs[0][1], outs[1][1]) < 1e-5


# ------------------------------------------------------------------------------------------------ TransitionDown backward
@pytest.mark.parametrize("B,H,W,level", [(2, 120, 160, 0), (3, 120, 160, 3), (2, 40, 56, 1)], ids=["l0", "l3-15x20", "odd-20x28"])
def test_transition_down_backward(B, H, W, level):
    net = Net(B, H, W, seed=400 + level)
    li = [i for i, L in enumerate(net.layers) if L["type"] == 2 and L["lin"] == level][0]
    L = net.layers[li]
    C = L["cin"]
    gi, go = net.level(level), net.level(level + 1)
    x, sums = net.set_input_level(level)
    P = net.set_bn_conv(li, 1)
    keep = net.keep(li)
    net.eng.pack_weights(net.params)
    n_in, n_out = B * gi["H"] * gi["W"], B * go["H"] * go["W"]
    net.eng.level_stats(level + 1).zero_()
    net.eng.forward_range(net.params, net.bn, None, True, li, li)      # produces the pooled output, its stats, the arg-max
    torch.cuda.synchronize()
    y = net.eng.level_view(level + 1)[..., :C].float()
    sums_o = net.eng.level_stats(level + 1)[:, :C].clone()
    off = net.q(E.Q_LAYER_AMAX_OFF, li)
    amax = net.eng._ws_view[off:off + n_out * C].view(B, go["H"], go["W"], C).long()
    dY = net.rand(B, go["H"], go["W"], go["Ct"], scale=1e-2)
    net.eng.set_level(level + 1, dY, "dact")
    acc_o = torch.randn(2, go["Ct"], generator=net.gen, device=DEV, dtype=torch.float64) * 1e-2 * n_out ** 0.5
    net.eng.level_acc(level + 1).copy_(acc_o)
    dX0 = net.rand(B, gi["H"], gi["W"], gi["Ct"], scale=1e-2)
    net.eng.set_level(level, dX0, "dact")
    net.eng.level_acc(level).zero_()
    net.eng.gstage().zero_()
    net.eng.backward_range(net.params, li, li)
    torch.cuda.synchronize()
    # ---- reference
    _, _, mean_o, rstd_o = R.bn_coef(sums_o, n_out, torch.ones(C, device=DEV), torch.zeros(C, device=DEV))
    A, Bc = (acc_o[0, :C] / n_out).float(), (acc_o[1, :C] / n_out).float()
    xh = (y - mean_o) * rstd_o
    dzp = keep[:, None, None, :] * (dY[..., :C] - A - xh * Bc)                          # pooled resolution, fp32
    dz = torch.zeros(B, gi["H"], gi["W"], C, device=DEV)
    Hp, Wp = go["H"], go["W"]
    sel = torch.nn.functional.one_hot(amax, 4).view(B, Hp, Wp, C, 2, 2).float()
    dz[:, :2 * Hp, :2 * Wp] = (sel * dzp[..., None, None]).permute(0, 1, 4, 2, 5, 3).reshape(B, 2 * Hp, 2 * Wp, C)
    dz = R.bf16r(dz).double()
    s, t, mean, rstd = R.bn_coef(sums[:, :C], n_in, P["gamma"], P["beta"])
    z = R.pre_act(x[..., :C], s, t)
    m = z > 0
    G = torch.einsum("bhwo,oc->bhwc", dz, P["w"].double()[:, :, 0, 0])
    g = torch.where(m, G, torch.zeros_like(G))
    k = (P["gamma"].float() * rstd).double()
    xm = x[..., :C].double() - mean.double()
    S1, S2 = g.sum((0, 1, 2)), (g * xm).sum((0, 1, 2)) * rstd.double()
    dX = R.bf16r((dX0[..., :C].double() + g * k).float())
    a = R.bf16r(torch.relu(z)).double()
    dW = torch.einsum("bhwo,bhwc->oc", dz, a)
    got = net.eng.level_view(level, grad=True)
    check_bf16(got[..., :C], (dX0[..., :C].double() + g * k), "TransitionDown dX", frac_ok=0.995)
    assert torch.equal(got[..., C:].float(), dX0[..., C:])
    gs = net.eng.gstage()
    pg = L["p_gamma"]
    check_f32(gs[L["p_w"]:L["p_w"] + C * C].view(C, C), dW, "TransitionDown dW")
    check_f32(gs[pg:pg + C], S2, "TransitionDown dgamma", tol=2e-3)
    check_f32(gs[pg + C:pg + 2 * C], S1, "TransitionDown dbeta", tol=2e-3)
    check_f32(gs[L["p_b"]:L["p_b"] + C], dzp.double().sum((0, 1, 2)), "TransitionDown dbias")
    acc = net.eng.level_acc(level)
    check_f32(acc[0, :C], k * S1, "A terms", tol=2e-3)
    check_f32(acc[1, :C], k * S2, "B terms", tol=2e-3)


# ------------------------------------------------------------------------------------------------ TransitionUp backward
@pytest.mark.parametrize("B,H,W,lin", [(2, 120, 160, 1), (4, 120, 160, 5), (2, 40, 56, 2)], ids=["60x80", "3x5", "odd-10x14"])
def test_transition_up_backward(B, H, W, lin):
    net = Net(B, H, W, seed=500 + lin)
    li = [i for i, L in enumerate(net.layers) if L["type"] == 3 and L["lin"] == lin][0]
    L = net.layers[li]
    lout = L["lout"]
    gi, go = net.level(lin), net.level(lout)
    x, _ = net.set_input_level(lin)
    w = net.rand(80, 80, 3, 3, scale=0.05)
    net.params[L["p_w"]:L["p_w"] + w.numel()] = w.reshape(-1)
    net.eng.pack_weights(net.params)
    y, sums_o = net.set_input_level(lout)                  # the up-sampled channels as the forward would have left them
    o, io = L["out_off"], L["in_off"]
    n_out = B * go["H"] * go["W"]
    dY = net.rand(B, go["H"], go["W"], go["Ct"], scale=1e-2)
    net.eng.set_level(lout, dY, "dact")
    acc_o = torch.randn(2, go["Ct"], generator=net.gen, device=DEV, dtype=torch.float64) * 1e-2 * n_out ** 0.5
    net.eng.level_acc(lout).copy_(acc_o)
    dX0 = net.rand(B, gi["H"], gi["W"], gi["Ct"], scale=1e-2)
    net.eng.set_level(lin, dX0, "dact")
    net.eng.gstage().zero_()
    net.eng.backward_range(net.params, li, li)
    torch.cuda.synchronize()
    _, _, mean_o, rstd_o = R.bn_coef(sums_o[:, o:o + 80], n_out, torch.ones(80, device=DEV), torch.zeros(80, device=DEV))
    A, Bc = (acc_o[0, o:o + 80] / n_out).float(), (acc_o[1, o:o + 80] / n_out).float()
    dzf = dY[..., o:o + 80] - A - (y[..., o:o + 80] - mean_o) * rstd_o * Bc
    dz = R.bf16r(dzf).double()
    # conv_transpose2d(x, w, stride 2) cropped to (H, W): its input gradient is conv2d(pad(dz), w, stride 2)
    Hf, Wf = 2 * gi["H"] + 1, 2 * gi["W"] + 1
    dzp = torch.zeros(B, Hf, Wf, 80, device=DEV, dtype=torch.float64)
    dzp[:, :go["H"], :go["W"]] = dz
    dX = R.nhwc(torch.nn.functional.conv2d(R.nchw(dzp), w.double(), stride=2))
    dW = torch.nn.grad.conv2d_weight(R.nchw(dzp), (80, 80, 3, 3), R.nchw(x[..., io:io + 80].double()), stride=2)
    # (conv2d_weight of the adjoint: weight[in, out] of the transposed convolution = weight[out', in'] of this conv2d)
    got = net.eng.level_view(lin, grad=True)
    check_bf16(got[..., io:io + 80], dX, "TransitionUp dX (plain store)", frac_ok=0.995)
    mask = torch.ones(gi["Ct"], dtype=torch.bool, device=DEV)
    mask[io:io + 80] = False
    assert torch.equal(got[..., mask].float(), dX0[..., mask])
    gs = net.eng.gstage()
    check_f32(gs[L["p_w"]:L["p_w"] + 80 * 80 * 9].view(80, 80, 3, 3), dW, "TransitionUp dW")
    check_f32(gs[L["p_b"]:L["p_b"] + 80], dzf.double().sum((0, 1, 2)), "TransitionUp dbias")


# ------------------------------------------------------------------------------------------------ first convolution
def test_first_conv_forward_and_weight_gradient():
    B, H, W = 2, 120, 160
    net = Net(B, H, W, seed=600)
    L = net.layers[0]
    xin = net.rand(B, 3, H, W)
    w = net.rand(48, 3, 3, 3, scale=0.2)
    bias = 0.1 * torch.randn(48, generator=net.gen, device=DEV)
    net.params[L["p_w"]:L["p_w"] + w.numel()] = w.reshape(-1)
    net.params[L["p_b"]:L["p_b"] + 48] = bias
    net.eng.pack_weights(net.params)
    net.eng.level_stats(0).zero_()
    net.eng.forward_range(net.params, net.bn, xin, True, 0, 0)
    torch.cuda.synchronize()
    ref = R.nhwc(torch.nn.functional.conv2d(xin.double(), w.double(), bias.double(), padding=1))
    got = net.eng.level_view(0)[..., :48]
    check_bf16(got, ref, "first conv (fp32 operands on CUDA cores)")
    sums = net.eng.level_stats(0)[:, :48].clone()
    check_f32(sums, R.channel_sums(got), "Σx, Σx² of the 48 channels", tol=1e-5)
    n = B * H * W
    dY = net.rand(B, H, W, 288, scale=1e-2)
    net.eng.set_level(0, dY, "dact")
    acc = torch.randn(2, 288, generator=net.gen, device=DEV, dtype=torch.float64) * 1e-2 * n ** 0.5
    net.eng.level_acc(0).copy_(acc)
    net.eng.gstage().zero_()
    net.eng.backward_range(net.params, 0, 0)
    torch.cuda.synchronize()
    _, _, mean, rstd = R.bn_coef(sums, n, torch.ones(48, device=DEV), torch.zeros(48, device=DEV))
    A, Bc = (acc[0, :48] / n).float(), (acc[1, :48] / n).float()
    dzf = dY[..., :48] - A - (got.float() - mean) * rstd * Bc
    dz = R.bf16r(dzf).double()
    dW = torch.nn.grad.conv2d_weight(xin.double(), (48, 3, 3, 3), R.nchw(dz), padding=1)
    gs = net.eng.gstage()
    check_f32(gs[L["p_w"]:L["p_w"] + 48 * 27].view(48, 3, 3, 3), dW, "first conv dW")
    check_f32(gs[L["p_b"]:L["p_b"] + 48], dzf.double().sum((0, 1, 2)), "first conv dbias")


# ------------------------------------------------------------------------------------------------ head
@pytest.mark.parametrize("C,weighted,absent", [(4, True, False), (4, True, True), (2, False, False), (3, True, False)],
                         ids=["4cls-weighted", "4cls-absent-classes-inf-weight", "2cls-plain", "3cls-weighted"])
def test_head_forward_backward(C, weighted, absent):
    """F.normalize -> Conv1x1 -> /T -> softmax -> cross_entropy(probabilities) -> argmax / confusion counts and its
    backward (tiramisu.py:105,120-125, SimpleTrain.py:15-20, TrainingBase.py:12-23) incl. a batch in which classes 2
    and 3 never occur: their weight is 1/0 = inf in the reference and must never be touched."""
    B, H, W = 2, 120, 160
    net = Net(B, H, W, C=C, seed=700 + C)
    f = net.rand(B, H, W, 288)
    net.eng.set_level(0, f, "act")
    n = net.eng.n_params
    pw = n - C * 288 - C
    wc = 0.5 * torch.randn(C, 288, generator=net.gen, device=DEV)
    bc = 0.1 * torch.randn(C, generator=net.gen, device=DEV)
    net.params[pw:pw + C * 288] = wc.reshape(-1)
    net.params[pw + C * 288:] = bc
    hi = 2 if absent else C
    y = torch.randint(0, hi, (B, H, W), generator=net.gen, device=DEV)
    probs = torch.empty(B, C, H, W, device=DEV)
    mask = torch.empty(B, H, W, dtype=torch.uint8, device=DEV)
    net.eng.head(net.params, y=y, weighted=weighted, probs=probs, mask=mask, backward=1)
    net.eng.backward_range(net.params, net.nlayers, net.nlayers)   # classifier weight gradient only
    torch.cuda.synchronize()
    # ---- reference in float64 (channels: physical [skip128 | up80 | new80] -> reference [up80 | skip128 | new80])
    fr = torch.cat([f[..., 128:208], f[..., :128], f[..., 208:]], -1).double().requires_grad_(True)
    wd, bd = wc.double().requires_grad_(True), bc.double().requires_grad_(True)
    fn = torch.nn.functional.normalize(R.nchw(fr), dim=1, eps=1e-12)
    logits = torch.einsum("bchw,kc->bkhw", fn, wd) + bd[None, :, None, None]
    p = torch.softmax(logits / 0.05, 1)
    cw = None
    if weighted:
        cw = torch.reciprocal(torch.bincount(y.reshape(-1), minlength=C).double())
        assert not absent or torch.isinf(cw[2:]).all()
    loss = torch.nn.functional.cross_entropy(p, y, weight=cw)
    loss.backward()
    hist, cm, ls, bad = net.eng.scalars()
    assert int(bad) == 0 and hist[:C].tolist() == torch.bincount(y.reshape(-1), minlength=C).tolist()
    got_loss = float(ls[0] / ls[1])
    assert abs(got_loss - float(loss)) < 1e-5 * max(1.0, abs(float(loss))), (got_loss, float(loss))
    assert float((probs.double() - p.detach()).abs().max()) < 5e-5
    pred = p.detach().float().argmax(1)      # fp32 probabilities, first maximum: what torch.max(outputs, 1) returns
    pred64 = p.detach().argmax(1)
    close_call = (p.detach().topk(2, 1).values.diff(dim=1).abs().squeeze(1) < 1e-6)
    mism = (mask.long() != pred64) & ~close_call
    assert int(mism.sum()) == 0, int(mism.sum())
    assert torch.equal(mask.long(), probs.argmax(1)), "mask must be the first-maximum argmax of the returned probabilities"
    ref_cm = torch.bincount((y * C + mask.long()).reshape(-1), minlength=C * C)
    assert cm[:C * C].tolist() == ref_cm.tolist()
    df = fr.grad                                                                 # reference order
    df_phys = torch.cat([df[..., 80:208], df[..., :80], df[..., 208:]], -1)
    check_bf16(net.eng.level_view(0, grad=True), df_phys, "d(features)", frac_ok=0.995)
    gs = net.eng.gstage()
    check_f32(gs[pw + C * 288:], bd.grad, "classifier bias gradient")
    check_f32(gs[pw:pw + C * 288].view(C, 288), wd.grad, "classifier weight gradient (dl stored as bf16)", tol=3e-3)
    assert torch.isfinite(gs).all()


def test_head_excludes_and_counts_bad_labels():
    B, H, W, C = 1, 64, 64, 4
    net = Net(B, H, W, C=C, seed=9)
    net.eng.set_level(0, net.rand(B, H, W, 288), "act")
    net.params[-C * 288 - C:] = 0.1 * torch.randn(C * 288 + C, generator=net.gen, device=DEV)
    y = torch.randint(0, C, (B, H, W), generator=net.gen, device=DEV)
    y[0, :2] = 255
    net.eng.head(net.params, y=y, weighted=True)
    torch.cuda.synchronize()
    assert int(net.eng.scalars()[3]) == 2 * W
    from sim2real_lane_segment_b200._lib import RightLaneError
    from sim2real_lane_segment_b200.runtime import Runtime
    with pytest.raises(RightLaneError):
        Runtime.check_labels(type("R", (), {"num_cls": C})(), net.eng)


# ------------------------------------------------------------------------------------------------ optimizers, dropout RNG
def test_adamw_matches_torch_over_ten_steps():
    """torch.optim.AdamW(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4) (SimpleTrain.py:28)"""
    n = 1_000_003
    g = torch.Generator(device=DEV).manual_seed(1)
    p0 = torch.randn(n, device=DEV, generator=g)
    mine, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    ref = torch.nn.Parameter(p0.clone().double())
    opt = torch.optim.AdamW([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    for step in range(1, 11):
        grad = torch.randn(n, device=DEV, generator=g) * (10.0 ** torch.randint(-4, 1, (n,), device=DEV, generator=g).float())
        E.adamw_step(mine, grad, m, v, 1e-3, 0.9, 0.999, 1e-8, 1e-4, step)
        ref.grad = grad.double()
        opt.step()
    torch.cuda.synchronize()
    st = opt.state[ref]
    assert R.rel(mine, ref.detach()) < 1e-6 and R.rel(m, st["exp_avg"]) < 1e-6 and R.rel(v, st["exp_avg_sq"]) < 1e-6
    assert float((mine.double() - ref.detach()).abs().max()) < 2e-6


def test_sgd_nesterov_matches_torch_over_ten_steps():
    """torch.optim.SGD(momentum=0.9, nesterov=True, weight_decay) (MMETrainingModule.py:16-19)"""
    n = 500_009
    g = torch.Generator(device=DEV).manual_seed(2)
    p0 = torch.randn(n, device=DEV, generator=g)
    mine, buf = p0.clone(), torch.zeros(n, device=DEV)
    ref = torch.nn.Parameter(p0.clone().double())
    opt = torch.optim.SGD([ref], lr=1e-3 / 3, momentum=0.9, nesterov=True, weight_decay=1e-4)
    for step in range(10):
        grad = torch.randn(n, device=DEV, generator=g)
        E.sgd_nesterov_step(mine, grad, buf, 1e-3 / 3, 0.9, 1e-4, step == 0)
        ref.grad = grad.double()
        opt.step()
    torch.cuda.synchronize()
    assert R.rel(mine, ref.detach()) < 1e-6 and R.rel(buf, opt.state[ref]["momentum_buffer"]) < 1e-6


def test_dropout_rng_statistics():
    """fill_masks_kernel (the in-kernel Dropout2d RNG the bench uses): keep-rate 0.8 within 4 sigma per layer, values in
    {0, 1.25}, independent across samples / channels / layers / steps / seeds"""
    B = 64
    net = Net(B, 64, 64, seed=3)
    masks = {}
    for step, seed in ((0, 1), (1, 1), (0, 2)):
        net.eng.set_dropout(0.2, seed=seed, step=step)
        torch.cuda.synchronize()
        masks[(step, seed)] = [net.eng.layer_mask_view(i).clone() for i, L in enumerate(net.layers) if L["type"] in (1, 2)]
    m0 = masks[(0, 1)]
    allv = torch.cat([m.reshape(-1) for m in m0])
    assert set(allv.unique().tolist()) == {0.0, 1.25}
    nall = allv.numel()
    keep = float((allv > 0).float().mean())
    assert abs(keep - 0.8) < 4 * (0.16 / nall) ** 0.5, keep
    for m in m0:
        k, nn = float((m > 0).float().mean()), m.numel()
        assert abs(k - 0.8) < 4.5 * (0.16 / nn) ** 0.5, (k, nn)
    big = [m for m in m0 if m.shape[1] >= 128][0]
    per_sample = (big > 0).float().mean(1)                       # every sample of a layer draws its own channels
    assert float(per_sample.std()) > 0 and float((per_sample - 0.8).abs().max()) < 0.25
    rows = (big > 0).float()
    assert float((rows[0] == rows[1]).float().mean()) < 0.8      # ~0.68 expected for independent rows
    for other in (masks[(1, 1)], masks[(0, 2)]):                 # next step / other seed: decorrelated
        a = torch.cat([m.reshape(-1) for m in m0]) > 0
        b = torch.cat([m.reshape(-1) for m in other]) > 0
        agree = float((a == b).float().mean())
        assert abs(agree - 0.68) < 0.01, agree
    net.eng.set_dropout(0.0)
    torch.cuda.synchronize()
    assert all(bool((net.eng.layer_mask_view(i) == 1).all()) for i, L in enumerate(net.layers) if L["type"] in (1, 2))
