"""Test infrastructure: the reference MLP training step (collaborative/mlp.py:88-115 forward x2, hinge,
backward) restated in torch fp32 on the CPU with a bf16 rounding at exactly the points where csrc/mlp.cu
stores bf16 (input rows, layer weights, pre-BN activations Z, post-ReLU activations, dZ, and the
activation gradients handed from layer to layer).  Against this, the CUDA path may differ only by fp32
summation order, so the comparison is tight (it checks the kernels' logic); against the plain fp32 oracle
the difference is the bf16 arithmetic itself (it checks the tolerance claim)."""
import torch

BN_EPS, BN_MOMENTUM = 1e-5, 0.1


def bf(x):
    return x.bfloat16().float()


def n_layers(p):
    n = 0
    while f"fcs.{n}.weight" in p:
        n += 1
    return n


def mlp_grads_bf16(p, batch, use_bn=True):
    """p: {name: fp32 tensor} (running statistics are updated in place); batch: user/pos/neg[/pos_meta/neg_meta]
    LongTensors.  Returns (loss, dense grads {name: tensor}, per-lookup embedding grads
    {table: (ids, rows)} with the user rows of the two passes pre-summed as the kernels stage them)."""
    L = n_layers(p)
    user, B = batch["user"], batch["user"].shape[0]
    Wb = [bf(p[f"fcs.{l}.weight"]) for l in range(L)]
    metas = []
    while f"metadata_embeddings.{len(metas)}.weight" in p:
        metas.append(p[f"metadata_embeddings.{len(metas)}.weight"])
    passes = []
    for item, meta in ((batch["pos"], batch.get("pos_meta")), (batch["neg"], batch.get("neg_meta"))):
        cols = [p["user.weight"][user], p["item.weight"][item]]
        cols += [t[meta[:, f]] for f, t in enumerate(metas)]
        x = bf(torch.cat(cols, 1))
        cache = []
        for l in range(L):
            z = bf(x @ Wb[l].t() + p[f"fcs.{l}.bias"])
            if use_bn:
                mu = z.double().mean(0)
                var = ((z.double() ** 2).mean(0) - mu * mu).clamp_min(0)
                rstd = 1.0 / torch.sqrt(var.float() + BN_EPS)
                mu = mu.float()
                p[f"bns.{l}.running_mean"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mu)
                p[f"bns.{l}.running_var"].mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var.float() * (B / max(B - 1, 1)))
                xhat = (z - mu) * rstd
                y = xhat * p[f"bns.{l}.weight"] + p[f"bns.{l}.bias"]
            else:
                rstd, xhat, y = None, None, z
            a = bf(torch.relu(y))
            cache.append((x, xhat, rstd, a))
            x = a
        s = x @ p["output_layer.weight"].t() + p["output_layer.bias"]
        passes.append((cache, s[:, 0]))
    sp, sn = passes[0][1], passes[1][1]
    h = sn - sp + 1.0
    loss = torch.clamp(h, min=0).mean()
    g = (h >= 0).float() / B
    dense = {k: torch.zeros_like(v) for k, v in p.items() if k.startswith(("fcs", "bns", "output")) and "running" not in k
             and "num_batches" not in k}
    dxs = []
    for (cache, _), ds in zip(passes, (-g, g)):
        a_last = cache[-1][3]
        dense["output_layer.weight"] += (ds[None, :] @ a_last)
        da = ds[:, None] * p["output_layer.weight"]
        for l in reversed(range(L)):
            x, xhat, rstd, a = cache[l]
            dy = torch.where(a > 0, da, torch.zeros_like(da))
            if use_bn:
                S1, S2 = dy.sum(0), (dy * xhat).sum(0)
                dense[f"bns.{l}.bias"] += S1
                dense[f"bns.{l}.weight"] += S2
                dz = p[f"bns.{l}.weight"] * rstd * (dy - S1 / B - xhat * (S2 / B))
            else:
                dz = dy
            dense[f"fcs.{l}.bias"] += dz.sum(0)
            dzb = bf(dz)
            dense[f"fcs.{l}.weight"] += dzb.t() @ x
            da = dzb @ Wb[l]
            if l > 0:
                da = bf(da)
        dxs.append(da)
    D = p["user.weight"].shape[1]
    sparse = {"user.weight": (user, dxs[0][:, :D] + dxs[1][:, :D]),
              "item.weight": (torch.cat([batch["pos"], batch["neg"]]), torch.cat([dxs[0][:, D:2 * D], dxs[1][:, D:2 * D]]))}
    for f in range(len(metas)):
        sl = slice((2 + f) * D, (3 + f) * D)
        sparse[f"metadata_embeddings.{f}.weight"] = (
            torch.cat([batch["pos_meta"][:, f], batch["neg_meta"][:, f]]), torch.cat([dxs[0][:, sl], dxs[1][:, sl]]))
    return float(loss), dense, sparse


def sgd_step_bf16(p, batch, lr, use_bn=True):
    """One SGD step of the emulated path, in place.  Returns the loss."""
    loss, dense, sparse = mlp_grads_bf16(p, batch, use_bn)
    for k, g in dense.items():
        p[k] -= lr * g
    for k, (ids, rows) in sparse.items():
        acc = torch.zeros_like(p[k])
        acc.index_add_(0, ids, rows)
        p[k] -= lr * acc
    return loss
