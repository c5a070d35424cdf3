"""ORACLE (test infrastructure): closed-form weight gradient of the MC Fisher-trace regulariser (dual density pass + two masked reverse passes
per transform + reverse of the sampling chain), checked against autograd's double backward on the oracle (fp64).  This is the derivation behind
rabi_b200/fisher.py::fisher_trace_weight_grad; run as a script or through tests/test_oracle.py."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import rbi_oracle as O

def check():
    torch.manual_seed(0)
    dx, D, hidden, K, B, S = 7, 3, [12, 10], 3, 5, 4
    mean, std = 0.3 * torch.randn(dx), 0.5 + torch.rand(dx)
    m = O.OracleMAF(dx, D, hidden_dims=hidden, num_transforms=K, shuffle=False,
                    embedding_net=torch.nn.Sequential(O.ZScoreLayer(mean, std), torch.nn.Identity()),
                    log_scale_min_clip=-7, log_scale_max_clip=5, stable=False).double()
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.3 * torch.randn_like(p))
    x = torch.randn(B, dx, dtype=torch.float64)
    z = torch.randn(S, B, D, dtype=torch.float64)
    beta = 0.05
    # ---------------- reference: autograd double backward
    with O.injected_base_noise([z]):
        q = m(x)
        reg = O.mc_fisher_trace(m, x, q, S, create_graph=True)
    loss = beta * reg.mean()
    params = list(m.parameters())
    ref = torch.autograd.grad(loss, params)

    # ---------------- closed form
    lo, hi = -7.0, 5.0
    P = S * B
    ctx = ((x - mean.double()) / std.double())          # [B, C]
    ctxp = ctx.repeat(S, 1)                              # [P, C] pair p = s * B + b
    nets = [getattr(m, f"t{k}").nn for k in range(K)]
    def masked(k):
        return [(l.weight * l.mask, l.bias) for l in nets[k].layers]
    # sampling chain with graph (its reverse is what rabi_maf_rsample_bwd computes); here autograd stands in for it
    with O.injected_base_noise([z]):
        theta_g = m(x).rsample((S,))                     # [S, B, D] with graph to the weights
    theta = theta_g.detach().reshape(P, D)
    # primal density pass, storing everything
    C = dx
    st = []
    y = theta
    for k in range(K - 1, -1, -1):
        (W0, b0), (W1, b1), (Wo, bo) = [(w.detach(), b.detach()) for w, b in masked(k)]
        a0 = ctxp @ W0[:, :C].T + b0 + y @ W0[:, C:].T
        h0 = a0.clamp_min(0); m0 = (a0 > 0).double()
        a1 = h0 @ W1.T + b1
        h1 = a1.clamp_min(0); m1 = (a1 > 0).double()
        o = h1 @ Wo.T + bo
        mu, s = o[:, :D], o[:, D:]
        u = torch.exp(s) * y + mu
        st.append(dict(k=k, y=y, h0=h0, m0=m0, h1=h1, m1=m1, s=s, u=u, W0=W0, W1=W1, Wo=Wo))
        y = u
    u0 = y
    st = {d["k"]: d for d in st}
    # first reverse: score g = d ell / d x   (ell = log q = sum s + log N(u0))
    ubar = -u0
    gA = {}
    for k in range(K):
        d = st[k]
        es = torch.exp(d["s"])
        obar = torch.cat([ubar, ubar * es * d["y"] + 1.0], 1)
        ybar = ubar * es
        a1bar = (obar @ d["Wo"]) * d["m1"]
        a0bar = (a1bar @ d["W1"]) * d["m0"]
        ybar = ybar + a0bar @ d["W0"][:, C:]
        gA[k] = a0bar
        ubar = ybar
    g = sum(gA[k] @ st[k]["W0"][:, :C] for k in range(K)) / std.double()     # [P, dx] score per pair
    w = beta / (B * S)
    xdot = 2 * w * g                                                        # direction, held constant
    cdot = xdot / std.double()                                              # tangent of the z-scored context
    # tangent (forward-mode) density pass with the stored masks
    ydot = torch.zeros_like(theta)
    for k in range(K - 1, -1, -1):
        d = st[k]
        a0d = cdot @ d["W0"][:, :C].T + ydot @ d["W0"][:, C:].T
        h0d = a0d * d["m0"]
        a1d = h0d @ d["W1"].T
        h1d = a1d * d["m1"]
        od = h1d @ d["Wo"].T
        mud, sd = od[:, :D], od[:, D:]
        es = torch.exp(d["s"])
        ud = es * (sd * d["y"] + ydot) + mud
        d.update(h0d=h0d, h1d=h1d, sd=sd, ydot=ydot, ud=ud)
        ydot = ud
    u0d = ydot
    # reverse over the dual program, seed 1 on ell_dot = sum sd - u0 . u0d
    ut = -u0            # adjoint of u0_dot
    ub = -u0d           # adjoint of u0
    grads = {}
    for k in range(K):
        d = st[k]
        es = torch.exp(d["s"])
        ot = torch.cat([ut, 1.0 + ut * es * d["y"]], 1)                                   # adjoints of (mu_dot, s_dot)
        yt = ut * es
        ob = torch.cat([ub, ub * es * d["y"] + ut * es * (d["sd"] * d["y"] + d["ydot"])], 1)   # adjoints of (mu, s)
        yb = ub * es + ut * es * d["sd"]
        a1t = (ot @ d["Wo"]) * d["m1"]; a0t = (a1t @ d["W1"]) * d["m0"]; yt = yt + a0t @ d["W0"][:, C:]
        a1b = (ob @ d["Wo"]) * d["m1"]; a0b = (a1b @ d["W1"]) * d["m0"]; yb = yb + a0b @ d["W0"][:, C:]
        gWo = ot.T @ d["h1d"] + ob.T @ d["h1"]
        gW1 = a1t.T @ d["h0d"] + a1b.T @ d["h0"]
        gW0 = torch.cat([a0t.T @ cdot + a0b.T @ ctxp, a0t.T @ d["ydot"] + a0b.T @ d["y"]], 1)
        grads[k] = [(gW0, a0b.sum(0)), (gW1, a1b.sum(0)), (gWo, ob.sum(0))]
        ut, ub = yt, yb
    theta_bar = ub                                                                         # d phi / d theta
    # path through the reparameterised samples: reverse of the sampling chain seeded with theta_bar
    samp = torch.autograd.grad(theta_g.reshape(P, D), params, theta_bar, allow_unused=True)
    # assemble in parameter order (weight, bias per layer per transform), applying the masks
    out = []
    i = 0
    for k in range(K):
        for l, layer in enumerate(nets[k].layers):
            gw, gb = grads[k][l]
            out.append(gw * layer.mask.double() + (samp[i] if samp[i] is not None else 0)); i += 1
            out.append(gb + (samp[i] if samp[i] is not None else 0)); i += 1
    # params order check
    names = [n for n, _ in m.named_parameters()]
    worst = 0.0
    for n, a, b in zip(names, out, ref):
        e = float((a - b).abs().max()) / max(float(b.abs().max()), 1e-12)
        worst = max(worst, e)
    print("parameter order:", names[:6], "...")
    print("closed form vs autograd double backward: worst relative deviation", worst)
    assert worst < 1e-9
    return worst


if __name__ == "__main__":
    check()
