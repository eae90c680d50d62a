import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def build_all():
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def prior_arrays():
    return np.load(os.path.join(GOLDEN, "prior_synth.npz"))


@pytest.fixture(scope="session")
def oracle_mod(build_all):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    return orc


@pytest.fixture(scope="session")
def omodel(oracle_mod, prior_arrays):
    return oracle_mod.OracleModel(os.path.join(GOLDEN, "model_synth.npz"), prior_arrays)


@pytest.fixture(scope="session")
def oopt(oracle_mod, omodel, prior_arrays):
    return oracle_mod.OracleOptimizer(omodel, int(prior_arrays["num_parts"]), prior_arrays["part_map"])


@pytest.fixture(scope="session")
def model(build_all, prior_arrays):
    from avatar_b200 import AvatarModel, GaussianMixture
    g = GaussianMixture.from_arrays(prior_arrays["weights"], prior_arrays["means"], prior_arrays["covs"])
    return AvatarModel(npz_path=os.path.join(GOLDEN, "model_synth.npz"), pose_prior=g)


@pytest.fixture(scope="session")
def frames(model, oracle_mod, omodel, prior_arrays):
    """a few deterministic synthetic frames: (x_gt, x_init, cloud Nx3, labels N)"""
    from harness import synth
    out = []
    for seed in range(3):
        rng = np.random.default_rng(1000 + seed)
        x_gt = synth.random_params(model, rng)
        x0 = synth.perturbed_start(model, x_gt, rng)
        cloud_gt, _, _ = omodel.update_x(x_gt)
        pts, lab, _, _ = synth.render_cloud(model, cloud_gt, prior_arrays["part_map"])
        out.append((x_gt, x0, pts, lab))
    return out
