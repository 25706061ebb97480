"""Golden vectors for the local-volatility path, produced by the REAL reference in this container:
`om3.IVModel.get_volatility_batch` + `om3.simulate_local_vol_paths_antithetic` (om3:263-333) driving a freshly
initialised `ImprovedIVNetwork` (nniv:109-155).  Run:  python oracle/gen_golden_localvol.py
Writes tests/golden/ref_localvol.npz.  Test infrastructure only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import gen_golden as G  # noqa: E402


def main():
    import torch

    om3, _, _, _ = G.load_reference()
    import NN_training_stock_iv as nniv  # the reference's IV-surface module (same directory as om3)

    out = {}
    for tag, H, L, wmul in (("h64", 64, 4, 0.6), ("h32", 32, 2, 0.08)):
        torch.manual_seed(5 + H)
        cfg = nniv.TrainingConfig(hidden_dim=H, num_hidden_layers=L)
        net = nniv.ImprovedIVNetwork(cfg)
        with torch.no_grad():  # a fresh network outputs ~0; move it into the range of a volatility surface
            net.output.bias.fill_(0.22)
            net.output.weight.mul_(wmul)
            for layer in net.layers:  # non-trivial LayerNorm affine parameters
                layer[1].weight.add_(0.1 * torch.randn(H))
                layer[1].bias.add_(0.05 * torch.randn(H))
        sc = nniv.DataScaler()
        sc.m_mean, sc.m_scale, sc.tau_mean, sc.tau_scale, sc.S0 = 0.03, 0.17, 0.4, 0.35, 100.0
        net.scaler = sc
        ivm = om3.IVModel(net)
        weights = np.concatenate([v.detach().cpu().numpy().reshape(-1) for v in net.state_dict().values()]).astype(np.float32)
        S0, r, T, K, M, N, seed = 100.0, 0.05, 1.0, 105.0, 256, 24, 11
        S = om3.simulate_local_vol_paths_antithetic(S0, r, T, M, N, ivm, K, np.random.default_rng(seed))
        Zh = np.random.default_rng(seed).standard_normal((N, M // 2))
        spots = np.linspace(40.0, 220.0, 97)
        sig = np.stack([ivm.get_volatility_batch(K, spots, tau) for tau in (1.0, 0.3, 1e-9)])
        out.update({f"{tag}_weights": weights, f"{tag}_meta": np.array([H, L, sc.m_scale, sc.tau_scale, cfg.epsilon]),
                    f"{tag}_args": np.array([S0, r, T, K, M, N]), f"{tag}_Zh": Zh, f"{tag}_S": S, f"{tag}_spots": spots,
                    f"{tag}_sigma": sig})
    np.savez_compressed(os.path.join(G.OUT, "ref_localvol.npz"), **out)
    print("wrote ref_localvol.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
