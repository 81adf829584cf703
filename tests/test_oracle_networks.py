"""Oracle networks vs the reference's own TorchScript modules (fixtures: tools/extract_assets.py)."""
import os

import numpy as np

import oracle


def test_policy_known_answer(weights, golden_dir):
    wc, _ = weights
    z = np.load(os.path.join(golden_dir, "mlp_kat.npz"))
    lat, act = oracle.policy_forward(wc, z["x"])
    # SURVEY.md 8(c) literal known answer
    assert np.allclose(act[0, :4], [165.2033, -166.0023, 104.1395, 131.9488], atol=2e-3)
    assert np.allclose(act, z["action"], rtol=2e-5, atol=2e-4)
    assert np.allclose(lat, z["latent"], rtol=2e-5, atol=2e-4)
    lat, act = oracle.policy_forward(wc, z["xs"])
    assert np.allclose(act, z["actions"], rtol=2e-5, atol=2e-5)
    assert np.allclose(lat, z["latents"], rtol=2e-5, atol=2e-5)


def test_policy_fp32_build(weights, golden_dir):
    wc, _ = weights
    z = np.load(os.path.join(golden_dir, "mlp_kat.npz"))
    _, act = oracle.policy_forward(wc, z["xs"], precision="f32")
    assert np.allclose(act, z["actions"], rtol=1e-4, atol=1e-4)


def test_actuator_known_answer(weights, golden_dir):
    wc, _ = weights
    z = np.load(os.path.join(golden_dir, "mlp_kat.npz"))
    t = oracle.actuator_forward(wc, z["xa"])
    assert np.allclose(t, [19.5816, 22.6831, -19.3611, -4.5704, -15.3245], atol=2e-4)
    assert np.allclose(t, z["torque"].ravel(), rtol=1e-5, atol=1e-5)
    t = oracle.actuator_forward(wc, z["xas"])
    assert np.allclose(t, z["torques"].ravel(), rtol=1e-5, atol=2e-5)
