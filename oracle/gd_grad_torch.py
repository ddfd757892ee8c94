"""Independent autodiff of the unrolled momentum-GD training graph (torch float64, CPU,
``create_graph=True``): stands in for TensorFlow's ``opt.compute_gradients(self.mse_, self.theta_)``
(multi-label-cls/icnn-back.py:120-139), which cannot run here.  TEST INFRASTRUCTURE ONLY: pins
oracle/gd_grad_np.py and generates tests/golden/gd_grad.npz (oracle/gen_golden_gd_grad.py)."""
import numpy as np
import torch

PNAMES = ("Wy", "Wz", "Wu", "bu", "Wzu", "bzu", "Wyu", "byu", "Wzx", "bzx")


def torch_unrolled_grads(p, x, y0, trueY, nIter, lr, mom):
    """Independent unrolled graph written from multi-label-cls/icnn-back.py:109-133,255-305."""
    T = {k: [None if a is None else torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=True)
             for a in getattr(p, k)] for k in PNAMES}
    x = torch.tensor(x, dtype=torch.float64)
    L = p.L

    def energy(y):
        us, prev = [], x
        for i in range(L):
            u = prev @ T["Wu"][i] + T["bu"][i]
            if i < L - 1:
                u = torch.relu(u)
            us.append(u)
            prev = u
        prevU, prevZ = x, None
        for i in range(L + 1):
            z = (y * (prevU @ T["Wyu"][i] + T["byu"][i])) @ T["Wy"][i] + prevU @ T["Wzx"][i] + T["bzx"][i]
            if i > 0:
                z = z + (prevZ * torch.relu(prevU @ T["Wzu"][i] + T["bzu"][i])) @ T["Wz"][i]
            if i < L:
                z = torch.nn.functional.leaky_relu(z, p.alpha) if p.alpha else torch.relu(z)
            prevU = us[i] if i < L else None
            prevZ = z
        return z.reshape(-1)

    yi = torch.tensor(y0, dtype=torch.float64, requires_grad=True)
    vi = torch.zeros_like(yi)
    for _ in range(nIter):
        (gi,) = torch.autograd.grad(energy(yi).sum(), yi, create_graph=True)
        vn = mom * vi - lr * gi
        yi = yi - mom * vi + (1.0 + mom) * vn
        vi = vn
    loss = ((yi - torch.tensor(trueY)) ** 2).mean()
    flat = [(k, i, t) for k in PNAMES for i, t in enumerate(T[k]) if t is not None]
    gs = torch.autograd.grad(loss, [t for _, _, t in flat], allow_unused=True)
    out = {k: [None] * len(T[k]) for k in PNAMES}
    for (k, i, _), g in zip(flat, gs):
        out[k][i] = None if g is None else g.numpy()
    return yi.detach().numpy(), float(loss.detach()), out
