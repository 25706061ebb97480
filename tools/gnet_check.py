"""Development check of the global network LSM (lsm_gnet.cu): gradients against torch autograd, then a small pricing."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import options_model_b200  # noqa: E402,F401
from options_model_b200 import engine as E  # noqa: E402

HP = dict(v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)


def torch_net():
    layers = [torch.nn.Linear(7, 128), torch.nn.ReLU()]
    for _ in range(2):
        layers += [torch.nn.Linear(128, 128), torch.nn.ReLU()]
    layers += [torch.nn.Linear(128, 1)]
    return torch.nn.Sequential(*layers)


def flat(net, grads=False):
    out = []
    for p in net.parameters():
        out.append((p.grad if grads else p.data).detach().reshape(-1).cpu().numpy())
    return np.concatenate(out)


def main():
    eng = E.Engine(0)
    torch.manual_seed(0)
    for n, shift in ((100, 0.0), (1000, 0.0), (100, 5.0), (128, 5.0), (1000, 5.0), (5000, 5.0)):
        net = torch_net()
        X = torch.randn(n, 7)
        y = torch.randn(n) - shift  # shift > 0: errors of one sign -> gradients are coherent sums, not cancellation noise
        loss = torch.nn.functional.mse_loss(net(X).squeeze(1), y)
        loss.backward()
        g_ref = flat(net, True)
        g, l = eng.gnet_grad_debug(X.numpy(), y.numpy(), flat(net))
        seg = {"W1": (0, 896), "b1": (896, 1024), "W2": (1024, 17408), "b2": (17408, 17536), "W3": (17536, 33920),
               "b3": (33920, 34048), "w4": (34048, 34176), "b4": (34176, 34177)}
        errs = {k: float(np.linalg.norm(g[a:b] - g_ref[a:b]) / (np.linalg.norm(g_ref[a:b]) + 1e-30)) for k, (a, b) in seg.items()}
        print(f"n={n} shift={shift} loss {l:.6f} vs {float(loss.detach()):.6f} relL2", {k: f"{v:.1e}" for k, v in errs.items()})
    # price against the torch restatement of the same algorithm (different RNG streams: statistical agreement)
    from oracle import lsm_oracle as orc
    model = E.heston(100.0, 0.05, 1.0, **HP)
    S = eng.paths(model, 40_000, 25, "f64", E.RngSpec(seed=8))
    Sn = S.cpu().numpy()
    for sem in ("reference", "textbook"):
        for variant, ep in (("gpu", 30), ("cpu", 3)):
            got = [eng.lsm_gnet(S, 100.0, 0.05, 1.0, "put", sem, variant=variant, epochs=ep, seed=sd)["price"] for sd in (1, 2, 3)]
            ref = []
            for sd in (1, 2):
                t0 = time.perf_counter()
                cf_mean, st = orc.lsm_global(Sn, 100.0, 0.05, 1.0, "put", orc.single_lsm_net_fit(variant, epochs=ep, seed=sd,
                                             inference_dropout=(sem == "reference")), target_ddof=0 if variant == "cpu" else 1) \
                    if sem == "reference" else (None, None)
                ref.append(cf_mean)
            print(sem, variant, "engine", [f"{p:.4f}" for p in got], "torch", ref, f"({time.perf_counter() - t0:.1f}s per torch run)")
    model = E.heston(100.0, 0.05, 1.0, **HP)
    S = eng.paths(model, 100_000, 50, "f32", E.RngSpec(seed=3))
    poly = eng.lsm(S, 100.0, 0.05, 1.0, "put")
    lin = eng.lsm_global(S, 100.0, 0.05, 1.0, "put")
    print("poly2 per-date", poly.price, poly.stderr, "| linear global", lin["price"])
    for variant, ep in (("gpu", 5), ("cpu", 2)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = eng.lsm_gnet(S, 100.0, 0.05, 1.0, "put", variant=variant, epochs=ep, seed=1)
        torch.cuda.synchronize()
        print(variant, {k: v for k, v in r.items() if k not in ("boundary", "ex_count")}, f"{time.perf_counter() - t0:.3f}s",
              eng.kernel_times())


if __name__ == "__main__":
    main()
