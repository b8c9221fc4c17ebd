"""The CPU oracle (oracle/ref_port.py) against the golden vectors produced by the
unmodified reference (oracle/make_golden.py).  Bit-exact where the op sequence is the
same; runs on CPU, no GPU needed."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden, rel_l2
from oracle import ref_port as P

torch.set_num_threads(1)


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def eq(a, b):
    assert torch.equal(torch.as_tensor(a), t(b))


@pytest.mark.parametrize("name", golden_names("disk_"))
def test_disk(name):
    m, z = load_golden(name)
    r = float("inf") if m["r"] == "inf" else m["r"]
    eq(P.fourier_disk_mask(t(z["x"]), r, m["inside_off"]), z["y"])
    eq(P.disk_binary_mask(z["x"].shape, r, 3, m["inside_off"])[0].to(torch.uint8), z["mask"])


@pytest.mark.parametrize("name", golden_names("diskmask2d_"))
def test_diskmask2d(name):
    m, z = load_golden(name)
    eq(P.disk_binary_mask(m["shape"], m["r"], 2, m["inside_off"]).to(torch.uint8), z["mask"])


@pytest.mark.parametrize("name", golden_names("planes_"))
def test_planes(name):
    m, z = load_golden(name)
    x = t(z["x"])
    coords = P.ellipsoid_shell_coords(x.shape[1:], m["a"], m["b"], m["c"])
    eq(coords.to(torch.int32), z["shell"])
    idx = P.sample_ellipsoid(x.shape[1:], m["a"], m["b"], m["c"], np.random.RandomState(m["ell_seed"]))
    assert list(idx) == m["idx"]
    eq(P.plane_wave_spike(x, idx, m["intensity"]), z["y"])


@pytest.mark.parametrize("name", golden_names("wrap_"))
def test_wrap(name):
    m, z = load_golden(name)
    eq(P.wrap_artifact(t(z["x"]), m["alpha"]), z["y"])


@pytest.mark.parametrize("name", golden_names("sap_"))
def test_sap(name):
    m, z = load_golden(name)
    eq(P.salt_and_pepper(t(z["x"]), m["p"], t(z["u"])), z["y"])


@pytest.mark.parametrize("name", golden_names("gibbs_"))
def test_gibbs(name):
    m, z = load_golden(name)
    eq(P.gibbs_noise(t(z["x"]), m["alpha"]), z["y"])
    assert np.array_equal(P.gibbs_mask(z["x"].shape[1:], m["alpha"]).astype(np.uint8), z["mask"])


@pytest.mark.parametrize("name", golden_names("kspike_"))
def test_kspike(name):
    m, z = load_golden(name)
    loc = m["loc"]
    loc = tuple(tuple(l) for l in loc) if isinstance(loc[0], list) else tuple(loc)
    inten = m["intensity"]
    inten = tuple(inten) if isinstance(inten, list) else inten
    eq(P.kspace_spike(t(z["x"]), loc, inten), z["y"])
    eq(P.logabs_mean(t(z["x"])), z["logabs_mean25"])


def test_kspike_spatial_loc_default_intensity_raises():
    with pytest.raises(TypeError):
        P.kspace_spike(P.synthetic_volume(6, (2, 16, 12, 8)), (3, 5, 2), None)


@pytest.mark.parametrize("name", golden_names("layer_"))
def test_layer(name):
    m, z = load_golden(name)
    with torch.no_grad():
        eq(P.gibbs_layer(t(z["x"]), min(max(m["alpha"], 0.), 1.)), z["y"])


@pytest.mark.parametrize("name", golden_names("chain127_"))
def test_chain127(name):
    m, z = load_golden(name)
    y = P.chain_127(t(z["x"]), m["r"], tuple(m["idx"]), m["intensity"], m["alpha"], m["p"], t(z["u"]))
    eq(y, z["y4"])


def test_identities_from_reference_docstrings():
    """S&P p=0 is the identity (F:437-438); layer alpha=1 is the identity to fp32 noise (S:60-61)."""
    x = P.synthetic_volume(3, (1, 16, 12, 8))
    assert torch.equal(P.salt_and_pepper(x, 0.0, P.synthetic_uniform(3, x.shape)), x)
    with torch.no_grad():
        assert rel_l2(P.gibbs_layer(x, 1.0).numpy(), x.numpy()) < 1e-6


def test_reference_notebook_fft_kats():
    """fourier_images_disk_masks.ipynb cells 7, 8, 12."""
    k = P.kspace(torch.ones(3, 3), 2)
    assert abs(k[1, 1] - 9) < 1e-6 and abs(k.abs().sum() - 9) < 1e-5
    tile = torch.tensor([1., 1, 0, 0, 1, 1, 0, 0]).repeat(8, 1)
    k = torch.fft.fftn(tile)
    assert abs(k[0, 0] - 32) < 1e-4 and abs(k[0, 2] - (16 - 16j)) < 1e-4 and abs(k[0, 6] - (16 + 16j)) < 1e-4


@pytest.mark.parametrize("name", golden_names("chain127_"))
def test_exact_phase_oracle_is_the_reference_modulo_the_ill_conditioned_plane_wave(name):
    """oracle.ref_port.chain_127_exact_phase (float64, angle(0) = 0 where the disk zeroed the spike's bin) against
    the UNMODIFIED reference's stage-3 output: equal to fp32 rounding when the spike is inside the disk; otherwise
    equal after removing the +-f_s plane wave from both (the reference's phase there is rounding noise, SURVEY
    section 0), with the same injected amplitude."""
    from test_gpu_parity import _remove_plane_wave
    m, z = load_golden(name)
    y = P.chain_127_exact_phase(torch.from_numpy(z["x"]), m["r"], m["idx"], m["intensity"], m["alpha"]).numpy()
    if m["spike_in_ball"]:
        assert rel_l2(y, z["y3"]) <= 1e-6
    else:
        ra, aa = _remove_plane_wave(y, m["idx"])
        rb, ab = _remove_plane_wave(z["y3"], m["idx"])
        assert rel_l2(ra, rb) <= 1e-6
        assert np.allclose(aa, ab, rtol=1e-5)
