import numpy as np

from wbc_quadruped_dob_b200 import scenarios as S


def test_instances_do_not_depend_on_batch_or_shard():
    whole = S.make(9000, mode_mix=(0.25, 0.375, 0.375), pushes=True, terrain=True, seed=3)
    part = S.make(1000, mode_mix=(0.25, 0.375, 0.375), pushes=True, terrain=True, seed=3, start=4000)
    for k, v in part.items():
        assert np.array_equal(v, whole[k][..., 4000:5000]), k


def test_never_generates_the_spin_input_and_stays_in_limits():
    sc = S.make(5000, seed=9)
    assert (np.abs(sc["base_pos"]).sum(axis=0) > 0).all()          # main.cpp:584-588 would spin forever
    assert (sc["q"] >= S.QMIN[:, None]).all() and (sc["q"] <= S.QMAX[:, None]).all()
    R = sc["base_rot"].T.reshape(-1, 3, 3)
    assert np.allclose(R @ R.transpose(0, 2, 1), np.eye(3), atol=1e-12)


def test_empty_batch():
    sc = S.make(0)
    assert sc["q"].shape == (12, 0) and sc["mode"].shape == (0,)


def test_config_mode_mix():
    sc = S.make_config("trot_65536", n=20000)
    frac = np.bincount(sc["mode"], minlength=3) / 20000.0
    assert np.allclose(frac, [0.25, 0.375, 0.375], atol=0.02)
