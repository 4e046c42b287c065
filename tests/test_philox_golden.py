"""The frozen Philox stream fixture (tests/golden/philox_stream.json): CPU oracle here, CUDA path on the GPU."""
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_philox_golden import case  # noqa: E402
from helpers import run_gpu, run_oracle  # noqa: E402

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "philox_stream.json")))
SEED, OFF = 0x1234567890ABCDEF, 4_000_000_123


def check(theta_s, ll_s, acc_b, acc_s, mc_b, mc_s, rtol):
    np.testing.assert_allclose(theta_s[:, :, -1, :], G["theta_last"], rtol=rtol)
    np.testing.assert_allclose(ll_s[:, :, -1], G["loglike_last"], rtol=rtol)
    np.testing.assert_allclose(theta_s[:, -1, 29, :], G["theta_mid_cold"], rtol=rtol)
    assert acc_b.tolist() == G["accept_burnin"] and acc_s.tolist() == G["accept_sampling"]
    assert mc_b.tolist() == G["mc_accept_burnin"] and mc_s.tolist() == G["mc_accept_sampling"]


def test_oracle_reproduces_the_frozen_stream(oracle):
    import ctypes as C
    u = (C.c_double * 4)()
    oracle.lib().ora_philox_uniforms(SEED, OFF, 7, 3, 1, 0, u)
    assert list(u) == G["uniforms_chain4000000123_iter7_rung3_slot1_tag0"]
    c = case()
    assert c.data[0].tolist() == G["data_x"]
    r = run_oracle(oracle, c, "philox", seed=SEED, chain_offset=OFF)
    check(r.theta_sampling, r.loglike_sampling, r.accept_burnin, r.accept_sampling, r.mc_accept_burnin, r.mc_accept_sampling, 1e-13)


@pytest.mark.gpu
def test_gpu_reproduces_the_frozen_stream():
    g = run_gpu(case(), "philox", seed=SEED, chain_offset=OFF)
    check(g["theta_sampling"], g["loglike_sampling"], g["accept_burnin"], g["accept_sampling"], g["mc_accept_burnin"],
          g["mc_accept_sampling"], 1e-10)
