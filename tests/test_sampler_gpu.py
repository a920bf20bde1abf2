"""Bit-exact parity of the device gather / frame-stack / crop against the sampler oracle (utils/datasets.py:68-112)."""
import numpy as np
import pytest

from oracle import sampler_oracle as SO

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('F,A', [(28, 5), (29, 8), (69, 21)])
def test_state_sample_bit_exact(F, A):
    from fql_b200.datasets import Dataset
    raw = SO.make_synthetic_dataset(5000, F, A, seed=1, episode_len=100)
    ref = SO.OracleDataset.create_from_initial_dataset(raw, size=6000)
    ds = Dataset.create_from_initial_dataset(raw, size=6000)
    for B in (1, 256, 1000):
        np.random.seed(B)
        r = ref.sample(B)
        np.random.seed(B)
        g = ds.sample(B)
        assert set(g) == set(r)
        for k in r:
            gk = g[k].cpu().numpy()
            assert gk.dtype == r[k].dtype and np.array_equal(gk, r[k]), k
    assert ds.sample(0)['observations'].shape == (0, F)


@pytest.mark.parametrize('frame_stack,p_aug', [(None, None), (3, None), (3, 0.5), (1, 1.0), (4, 1.0)])
def test_pixel_sample_bit_exact(frame_stack, p_aug):
    from fql_b200.datasets import Dataset
    raw = SO.make_synthetic_dataset(400, 0, 5, seed=2, episode_len=37, pixels=True, hw=64)
    ref = SO.OracleDataset.create_from_initial_dataset(raw, size=401)
    ds = Dataset.create_from_initial_dataset(raw, size=401)
    ref.frame_stack = ds.frame_stack = frame_stack
    ref.p_aug = ds.p_aug = p_aug
    np.random.seed(5)
    st = np.random.get_state()
    for it in range(6):  # several draws so both the augmented and the plain branch occur at p_aug=0.5
        np.random.set_state(st)
        r = ref.sample(64)
        np.random.set_state(st)
        g = ds.sample(64)
        st = np.random.get_state()
        for k in r:
            gk = g[k].cpu().numpy()
            assert gk.shape == r[k].shape and gk.dtype == r[k].dtype and np.array_equal(gk, r[k]), (it, k)
    # indices at episode starts: stacked frames must clamp at the initial state, never cross into the previous episode
    idxs = np.array([0, 1, 37, 38, 39, 73, 399])
    r = ref.sample(len(idxs), idxs=idxs) if p_aug is None else None
    if r is not None:
        g = ds.sample(len(idxs), idxs=idxs)
        for k in r:
            assert np.array_equal(g[k].cpu().numpy(), r[k]), k
