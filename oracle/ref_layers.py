"""TEST INFRASTRUCTURE ONLY -- layer/model level CPU restatement built on ref_spmm.py,
replaying the fixtures of tests/golden (made by oracle/make_golden.py from the
reference's own code).  Follows stag/layers.py:84-145, stag/models.py:39-84,
stag/zoo/gcn.py:58-116, stag/zoo/graph_sage.py:44-119, dgl.nn.GINConv.forward.
"""
import numpy as np
import torch

from . import ref_spmm as R


def t(a):
    return torch.from_numpy(np.asarray(a))


def case_kind(name):
    if name.startswith("sage_"):
        return "sage_" + name.split("_")[1]
    if name.startswith("gin"):
        return "gin"
    if name.startswith("gat"):
        return "gat"
    return "gcn"


def gcn_norm_of(name):
    for n in ("right", "left", "none"):
        if name == "gcn_" + n:
            return n
    return "both"


def layer_forward(name, d, feat, w):
    """Reference StagLayer(base).forward with external noise w (pre-relu / pre-in-norm).
    feat and w may require grad.  Parameters are returned so their grads can be read."""
    src, dst, N = t(d["src"]), t(d["dst"]), int(d["num_nodes"])
    relu, inn = bool(d["relu"]), bool(d["in_norm"])
    kind = case_kind(name)
    w = R.apply_relu(w, relu)                                   # stag/layers.py:98-99
    if inn:
        w, _ = R.in_norm(dst, N, w)                             # stag/layers.py:102-105
    params = {}
    if kind == "gcn":
        W = t(d["p_weight"]).clone().requires_grad_(True) if "p_weight" in d else None
        b = t(d["p_bias"]).clone().requires_grad_(True) if "p_bias" in d else None
        params = {"weight": W, "bias": b}
        out = R.gcn_forward(src, dst, N, feat, w, W, b, norm=gcn_norm_of(name))
    elif kind.startswith("sage"):
        agg = kind.split("_")[1]
        Wn = t(d["p_fc_neigh__weight"]).clone().requires_grad_(True)
        Ws = t(d["p_fc_self__weight"]).clone().requires_grad_(True) if agg == "mean" else None
        b = t(d["p_bias"]).clone().requires_grad_(True)
        params = {"fc_neigh.weight": Wn, "fc_self.weight": Ws, "bias": b}
        out = R.sage_forward(src, dst, N, feat, w, Ws, Wn, b, aggregator_type=agg)
    elif kind == "gat":  # stag/zoo/gat.py:89-145: noise [E,H] scales the leaky-relu logits before the softmax
        Wfc = t(d["p_fc__weight"]).clone().requires_grad_(True)
        al = t(d["p_attn_l"]).clone().requires_grad_(True)
        ar = t(d["p_attn_r"]).clone().requires_grad_(True)
        b = t(d["p_bias"]).clone().requires_grad_(True)
        Wres = t(d["p_res_fc__weight"]).clone().requires_grad_(True) if "p_res_fc__weight" in d else None
        params = {"fc.weight": Wfc, "attn_l": al, "attn_r": ar, "bias": b, "res_fc.weight": Wres}
        H, F = al.shape[1], al.shape[2]
        ft = (feat @ Wfc.t()).view(N, H, F)
        el, er = (ft * al).sum(-1), (ft * ar).sum(-1)
        e = torch.nn.functional.leaky_relu(el[src] + er[dst], 0.2) * w                # [E,H]
        idx = dst.unsqueeze(-1).expand_as(e)
        mx = torch.full((N, H), float("-inf"), dtype=e.dtype).scatter_reduce(0, idx, e, reduce="amax", include_self=True)
        ex = torch.exp(e - mx[dst])
        a = ex / torch.zeros((N, H), dtype=e.dtype).index_add(0, dst, ex)[dst]         # dgl edge_softmax over in-edges
        out = torch.zeros((N, H, F), dtype=feat.dtype).index_add(0, dst, ft[src] * a.unsqueeze(-1))
        if Wres is not None:
            out = out + (feat @ Wres.t()).view(N, -1, F)
        out = out + b.view(1, H, F)
        if name.startswith("gat_last"):
            out = torch.nn.functional.elu(out.mean(-2))
        else:
            out = out.flatten(-2, -1)
    else:  # gin: (1+eps) h_v + sum_e w h_u -> Linear
        Wl = t(d["p_apply_func__weight"]).clone().requires_grad_(True)
        bl = t(d["p_apply_func__bias"]).clone().requires_grad_(True)
        eps = t(d["p_eps"])
        params = {"apply_func.weight": Wl, "apply_func.bias": bl}
        neigh = R.u_mul_e_sum(src, dst, N, feat, w)
        out = ((1 + eps) * feat + neigh) @ Wl.t() + bl
    return out, params, w


def replay_layer_case(name, d):
    """-> dict(out, dfeat, dw, grads{...}) recomputed by the restatement."""
    feat = t(d["feat"]).clone().requires_grad_(True)
    w = t(d["w"]).clone().requires_grad_("dw" in d)
    out, params, w_used = layer_forward(name, d, feat, w)
    out.backward(t(d["gout"]))
    res = {"out": out.detach(), "dfeat": feat.grad, "dw": w.grad, "w_used": w_used.detach()}
    res["grads"] = {k: v.grad for k, v in params.items() if v is not None and v.grad is not None}
    return res


def replay_model_rc_vi(d):
    """StagModel.loss_terms (stag/models.py:63-84) for the 2-layer GCN / per-channel Normal /
    vi=True fixture.  Returns (nll, reg, grads by state_dict key)."""
    src, dst, N = t(d["src"]), t(d["dst"]), int(d["num_nodes"])
    feat, y, mask = t(d["feat"]), t(d["y"]), t(d["mask"])
    P = {k[2:].replace("__", "."): t(d[k]).clone().requires_grad_(True) for k in d.files if k.startswith("p_")}
    eps = [t(d["eps0"]), t(d["eps1"])]
    S = eps[0].shape[0]
    acts = [torch.relu, lambda x: torch.softmax(x, dim=-1)]
    total_nll, total_reg = 0.0, 0.0
    for s in range(S):
        h = feat
        for i in range(2):
            pre = "%d." % i
            loc, scale = P[pre + "q_a.loc"], P[pre + "q_a.log_scale"].exp()   # stag/distributions.py:137-144
            w = R.reparam_normal(loc, scale, eps[i][s])
            h = R.gcn_forward(src, dst, N, h, w, P[pre + "base_layer.weight"], P[pre + "base_layer.bias"],
                              "both", acts[i])
        nll = -torch.distributions.Categorical(probs=h).log_prob(y)[mask].mean()   # stag/likelihoods.py:13-16
        reg = 0.0
        for i in range(2):
            pre = "%d." % i
            reg = reg + R.kl_normal(P[pre + "q_a.loc"], P[pre + "q_a.log_scale"].exp(),
                                    P[pre + "p_a.loc"], P[pre + "p_a.log_scale"].exp())
        total_nll, total_reg = total_nll + nll, total_reg + reg
    total_nll = total_nll / S
    total_reg = total_reg / S * float(d["kl_scaling"])
    (total_nll + total_reg).backward()
    return total_nll.detach(), total_reg.detach(), {k: v.grad for k, v in P.items() if v.grad is not None}
