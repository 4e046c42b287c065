"""Freezes the Philox stream layout (DESIGN.md "RNG"): a small run of the oracle in Philox mode, stored as
a fixture.  The oracle (CPU test) and the CUDA path (GPU test) must both keep reproducing it, so a change
of the counter layout, of the uniform construction or of the update order cannot slip through by changing
both sides at once.    python tests/golden/make_philox_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import Case, beta_ladder, run_oracle  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def case():
    x = np.round(np.random.Generator(np.random.Philox(12)).normal(3.0, 2.0, 25), 6)
    init = np.array([[1.0, 1.0], [-2.5, 3.0], [4.0, 0.5]])
    return Case("example_exact", [-10, 0], [10, np.inf], data=[x], beta=beta_ladder(5, 2.0), chains=3, burnin=40, samples=60,
                init=init, save_hot_draws=True)


if __name__ == "__main__":
    c = case()
    r = run_oracle(po, c, "philox", seed=0x1234567890ABCDEF, chain_offset=4_000_000_123)
    u = (po.C.c_double * 4)()
    po.lib().ora_philox_uniforms(0x1234567890ABCDEF, 4_000_000_123, 7, 3, 1, 0, u)
    out = {
        "what": "oracle, Philox mode, model example_exact, 3 chains x 5 rungs (alpha 2), 40 + 60 iterations, seed 0x1234567890ABCDEF, chain_offset 4000000123",
        "data_x": c.data[0].tolist(),
        "uniforms_chain4000000123_iter7_rung3_slot1_tag0": list(u),
        "theta_last": r.theta_sampling[:, :, -1, :].tolist(),
        "loglike_last": r.loglike_sampling[:, :, -1].tolist(),
        "theta_mid_cold": r.theta_sampling[:, -1, 29, :].tolist(),
        "accept_burnin": r.accept_burnin.tolist(), "accept_sampling": r.accept_sampling.tolist(),
        "mc_accept_burnin": r.mc_accept_burnin.tolist(), "mc_accept_sampling": r.mc_accept_sampling.tolist(),
    }
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "philox_stream.json"), "w"), indent=1)
    print("written")
