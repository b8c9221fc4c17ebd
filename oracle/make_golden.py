"""Generate tests/golden/*.npz by running the UNMODIFIED reference (this container only).

TEST INFRASTRUCTURE ONLY.  Usage (from the repo root, in the build container where
/root/reference is mounted):

    python oracle/make_golden.py

For each case it (1) runs the reference class imported from
/root/reference/source_code through oracle/monai_shim, (2) runs oracle/ref_port.py on
the same input and asserts torch.equal (bit-identical), (3) stores input, parameters
and the REFERENCE output.  /root/reference does not exist on the GPU box; the tests
read only the .npz files written here.
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, "monai_shim"))
sys.path.insert(0, "/root/reference/source_code")
sys.path.insert(0, ROOT)

warnings.filterwarnings("ignore", message="torch.meshgrid")

import filters_and_operators as RF  # noqa: E402  (the unmodified reference)
import stylization_layers as RS  # noqa: E402
import monai.transforms as MT  # noqa: E402
from oracle import ref_port as P  # noqa: E402

torch.set_num_threads(1)          # make pocketfft/MKL summation order reproducible
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
CASES = {}


def vol(seed, shape):
    return P.synthetic_volume(seed, shape)


def save(name, meta, **arrays):
    arrs = {k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(meta), **arrs)
    CASES[name] = meta


def same(a, b, what):
    assert torch.equal(torch.as_tensor(a), torch.as_tensor(b)), f"port != reference: {what}"


SHAPES4 = {
    "s16x12x8c2": (2, 16, 12, 8),
    "s6x10x12": (1, 6, 10, 12),
    "s32x32x16": (1, 32, 32, 16),
    "s12x10x31": (1, 12, 10, 31),
    "s9x15x25": (1, 9, 15, 25),
    "s8x6x155": (1, 8, 6, 155),
    "s24x16x20c4": (4, 24, 16, 20),
}

# ------------------------------------------------------------------ a1/a2 disk masks
for sname, shape in SHAPES4.items():
    x = vol(1, shape)
    full = sname in ("s16x12x8c2", "s9x15x25")
    radii = [(2.5, False), (4.0, False), (3.0, True), (12.5, False), (np.sqrt(20.0000001), False), (float("inf"), False), (0.0, False)]
    for r, off in (radii if full else radii[1:3]):
        t = RF.RandFourierDiskMaskd("image", r=r, inside_off=off, prob=1.)
        y = t({"image": x})["image"]
        m = RF.disk_mask(torch.zeros(shape), r=r, dim=3, inside_off=off).binary_mask
        same(P.fourier_disk_mask(x, r, off), y, f"disk {sname} r={r}")
        same(P.disk_binary_mask(shape, r, 3, off), m, "disk mask")
        rs = "inf" if np.isinf(r) else f"{r:.7g}"
        save(f"disk_{sname}_r{rs}_off{int(off)}", dict(kind="disk", r=("inf" if np.isinf(r) else float(r)), inside_off=off),
             x=x, y=y, mask=m[0].to(torch.uint8))
for shape in [(3, 20, 15), (2, 16, 16)]:
    for r in (3.0, 5.5):
        m = RF.disk_mask(torch.zeros(shape), r=r, dim=2, inside_off=True).binary_mask
        same(P.disk_binary_mask(shape, r, 2, True), m, "disk mask 2d")
        save(f"diskmask2d_{shape[1]}x{shape[2]}_r{r}", dict(kind="diskmask2d", r=r, inside_off=True, shape=shape),
             mask=m.to(torch.uint8))

# ------------------------------------------------------------------ a3/a4 ellipsoid + plane waves
ELL = {"s16x12x8c2": (5., 4., 3.), "s32x32x16": (10., 10., 5.), "s12x10x31": (4., 4., 10.), "s8x6x155": (3., 2., 50.),
       "s24x16x20c4": (8., 6., 7.)}
for sname, (a, b, c) in ELL.items():
    shape = SHAPES4[sname]
    x = vol(2, shape)
    for inten in (3.0, 9.5):
        t = RF.RandPlaneWaves_ellipsoid("image", a, b, c, intensity_value=inten, prob=1.)
        t.set_random_state(seed=11)
        t.ellipsoid.set_random_state(seed=5)
        y = t({"image": x})["image"]
        coords = P.ellipsoid_shell_coords(shape[1:], a, b, c)
        idx = P.sample_ellipsoid(shape[1:], a, b, c, np.random.RandomState(5))
        assert tuple(int(v) for v in t.idx) == idx
        same(P.plane_wave_spike(x, idx, inten), y, f"planes {sname}")
        save(f"planes_{sname}_I{inten}", dict(kind="planes", a=a, b=b, c=c, intensity=inten, idx=list(idx), ell_seed=5,
                                              n_shell=int(len(coords))), x=x, y=y, shell=coords.to(torch.int32))

# ------------------------------------------------------------------ a6 wrap
for sname, shape in SHAPES4.items():
    x = vol(3, shape)
    for alpha in ((0.25,) if int(np.prod(shape)) > 4000 else (0.0, 0.25, 0.5, 1.0)):
        y = RF.WrapArtifact(alpha)(x)
        yd = RF.WrapArtifactd("image", alpha)({"image": x})["image"]
        same(y, yd, "wrapd")
        same(P.wrap_artifact(x, alpha), y, f"wrap {sname}")
        save(f"wrap_{sname}_a{alpha}", dict(kind="wrap", alpha=alpha), x=x, y=y)

# ------------------------------------------------------------------ a5 salt and pepper
for sname in ("s16x12x8c2", "s9x15x25", "s24x16x20c4"):
    shape = SHAPES4[sname]
    x = vol(4, shape)
    for p in (0.0, 0.05, 0.35, 1.0):
        torch.manual_seed(99)
        u = torch.rand(x.size())
        torch.manual_seed(99)
        t = RF.SaltAndPepper(p)
        y = t.salt_and_pepper(x)
        same(P.salt_and_pepper(x, p, u), y, f"sap {sname}")
        save(f"sap_{sname}_p{p}", dict(kind="sap", p=p), x=x, u=u, y=y)

# ------------------------------------------------------------------ a8 GibbsNoise (2-D and 3-D)
G_SHAPES = dict(SHAPES4)
G_SHAPES.update({"p20x15c3": (3, 20, 15), "p16x16c2": (2, 16, 16)})
for sname, shape in G_SHAPES.items():
    x = vol(5, shape)
    big = int(np.prod(shape)) > 4000
    for alpha in ((0.3, 0.7) if big else (0.0, 0.3, 0.5, 0.7, 0.95, 1.0)):
        y = RF.GibbsNoise(alpha)(x)
        same(P.gibbs_noise(x, alpha), y, f"gibbs {sname}")
        save(f"gibbs_{sname}_a{alpha}", dict(kind="gibbs", alpha=alpha), x=x, y=y,
             mask=P.gibbs_mask(shape[1:], alpha).astype(np.uint8))

# ------------------------------------------------------------------ a10 KSpaceSpikeNoise
K_CASES = [
    ("s16x12x8c2", (3, 5, 2), 6.0),
    ("s16x12x8c2", (1, 9, 7, 4), 5.0),
    ("s16x12x8c2", ((0, 3, 5, 2), (1, 8, 6, 4), (5, 5, 5)), (5.0, 6.0, 4.5)),
    ("s16x12x8c2", (8, 6, 4), 6.0),          # DC bin: self-conjugate
    ("s16x12x8c2", (0, 0, 0), 6.0),          # Nyquist corner: self-conjugate
    ("s16x12x8c2", (1, 3, 5, 2), None),      # default intensity, one channel
    ("s9x15x25", (2, 11, 20), 7.0),
    ("s9x15x25", (4, 7, 12), 7.0),           # DC on all-odd shape
    ("s8x6x155", (1, 2, 140), 8.0),
    ("s24x16x20c4", ((2, 3, 5, 2), (0, 20, 9, 13)), (6.5, 7.5)),
    ("p20x15c3", (4, 9), 5.0),
    ("p20x15c3", (2, 13, 3), 5.5),
    ("p16x16c2", ((0, 3, 4), (1, 8, 8)), (5.0, 4.0)),
]
for n, (sname, loc, inten) in enumerate(K_CASES):
    shape = G_SHAPES[sname]
    x = vol(6, shape)
    y = RF.KSpaceSpikeNoise(loc, inten)(x)
    same(P.kspace_spike(x, loc, inten), y, f"kspike {n}")
    lm = P.logabs_mean(x)
    save(f"kspike_{n:02d}_{sname}", dict(kind="kspike", loc=loc, intensity=inten), x=x, y=y, logabs_mean25=lm)

# spatial loc + default intensity: the reference itself raises (F:940 hands a tuple of tensors to F:981)
for sname, loc in (("s16x12x8c2", (3, 5, 2)), ("p16x16c2", (5, 11))):
    try:
        RF.KSpaceSpikeNoise(loc, None)(vol(6, G_SHAPES[sname]))
        raise SystemExit("expected the reference to raise TypeError")
    except TypeError:
        pass

# ------------------------------------------------------------------ a13 GibbsNoiseLayer (4-D, 5-D)
L_SHAPES = {"s16x12x8c2": (2, 16, 12, 8), "s9x15x25": (1, 9, 15, 25), "b2c1_16x12x8": (2, 1, 16, 12, 8),
            "b2c2_8x6x10": (2, 2, 8, 6, 10), "b1c4_8x6x10": (1, 4, 8, 6, 10), "b3c3_6x4x9": (3, 3, 6, 4, 9)}
for sname, shape in L_SHAPES.items():
    x = vol(7, shape)
    for alpha in (0.2167, 0.5, 0.71, 0.9, 1.0):
        layer = RS.GibbsNoiseLayer(alpha)
        with torch.no_grad():
            y = layer(x)
        same(P.gibbs_layer(x, alpha), y, f"layer {sname}")
        with torch.no_grad():
            m = P.gibbs_layer_mask(x.shape[1:], torch.tensor([min(max(alpha, 0.), 1.)]))
        save(f"layer_{sname}_a{alpha}", dict(kind="layer", alpha=alpha), x=x, y=y, mask=m.to(torch.uint8))

# ------------------------------------------------------------------ a14 spike_layer (class-level R)
for sname, shape, inten in [("b2c1_16x12x8", (2, 1, 16, 12, 8), 6.0), ("b2c2_8x6x10", (2, 2, 8, 6, 10), 5.0)]:
    x = vol(8, shape)
    MT.Randomizable.R = np.random.RandomState(2024)
    y = RS.spike_layer(inten)(x)
    R = np.random.RandomState(2024)
    gate = R.rand()
    spatial = tuple(int(R.randint(0, k)) for k in shape[1:])
    val = R.uniform(inten, inten)
    locs = [(i,) + spatial for i in range(shape[0])]
    same(P.kspace_spike(x, locs, [val] * shape[0]), y, "spike_layer")
    save(f"spikelayer_{sname}", dict(kind="spikelayer", intensity=inten, seed=2024, spatial=list(spatial), gate=gate), x=x, y=y)
MT.Randomizable.R = np.random.RandomState()

# ------------------------------------------------------------------ host RNG order (A.7)
x = vol(9, (2, 16, 12, 8))
t = RF.RandGibbsNoise(prob=0.9, alpha=(0.2, 0.8)); t.set_random_state(seed=7)
ys = [t(x) for _ in range(3)]
save("rng_randgibbs", dict(kind="rng_randgibbs", prob=0.9, alpha=[0.2, 0.8], seed=7, sampled_alpha=float(t.sampled_alpha)),
     x=x, y0=ys[0], y1=ys[1], y2=ys[2])
t = RF.RandGibbsNoised(["image", "other"], prob=1.0, alpha=(0.1, 0.6)); t.set_random_state(seed=8)
d = t({"image": x, "other": x * 2})
save("rng_randgibbsd", dict(kind="rng_randgibbsd", alpha=[0.1, 0.6], seed=8, sampled_alpha=float(t.sampled_alpha)),
     x=x, y_image=d["image"], y_other=d["other"])
t = RF.RandFourierDiskMaskd("image", r=[2.0, 5.0], prob=0.7); t.set_random_state(seed=9)
outs = [t({"image": x})["image"] for _ in range(4)]
save("rng_randdisk", dict(kind="rng_randdisk", r=[2.0, 5.0], prob=0.7, seed=9, r_after=float(t.r)),
     x=x, **{f"y{i}": o for i, o in enumerate(outs)})
for cw in (True, False):
    t = RF.RandKSpaceSpikeNoise(prob=0.8, intensity_range=(5.0, 6.0), channel_wise=cw); t.set_random_state(seed=10)
    ys = [t(x) for _ in range(3)]
    save(f"rng_randkspike_cw{int(cw)}", dict(kind="rng_randkspike", prob=0.8, range=[5.0, 6.0], channel_wise=cw, seed=10,
                                             locs=[list(map(int, l)) for l in t.sampled_locs],
                                             ints=[float(v) for v in t.sampled_k_intensity]),
         x=x, y0=ys[0], y1=ys[1], y2=ys[2])
t = RF.RandKSpaceSpikeNoise(prob=1.0, intensity_range=None, channel_wise=True); t.set_random_state(seed=12)
y = t(x)
save("rng_randkspike_default", dict(kind="rng_randkspike_default", seed=12, locs=[list(map(int, l)) for l in t.sampled_locs],
                                    ints=[float(v) for v in t.sampled_k_intensity]), x=x, y=y)
t = RF.RandKSpaceSpikeNoised(["image", "label"], global_prob=1.0, prob=1.0, intensity_ranges={"image": (5., 6.), "label": (4., 5.)},
                             channel_wise=True, common_sampling=True, common_seed=42)
t.set_rand_state(seed=3)
d = t({"image": x, "label": x + 1})
save("rng_randkspiked", dict(kind="rng_randkspiked", seed=3), x=x, y_image=d["image"], y_label=d["label"])
t = RF.RandPlaneWaves_ellipsoid("image", 5., 4., 3., intensity_value=4.0, prob=0.6)
t.set_random_state(seed=21); t.ellipsoid.set_random_state(seed=22)
outs, idxs = [], []
for _ in range(4):
    outs.append(t({"image": x})["image"]); idxs.append(None if t.idx is None else [int(v) for v in t.idx])
save("rng_randplanes", dict(kind="rng_randplanes", seed=21, ell_seed=22, idxs=idxs), x=x, **{f"y{i}": o for i, o in enumerate(outs)})
t = RF.SaltAndPepper(0.2, prob=0.5); t.set_random_state(seed=31)
torch.manual_seed(5)
outs = [t({"image": x})["image"] for _ in range(4)]
save("rng_sap", dict(kind="rng_sap", p=0.2, prob=0.5, seed=31, torch_seed=5), x=x, **{f"y{i}": o for i, o in enumerate(outs)})

# ------------------------------------------------------------------ chain 127 (stage outputs)
for sname, shape, r, (a, b, c), es in [("s32x32x16", (1, 32, 32, 16), 12.5, (8., 8., 4.), 1),   # spike inside the ball
                                       ("s32x32x16", (1, 32, 32, 16), 4.0, (10., 10., 5.), 2),   # spike on a zeroed bin
                                       ("s12x10x31", (1, 12, 10, 31), 4.5, (3., 3., 3.), 3)]:
    x = vol(10, shape)
    t1 = RF.RandFourierDiskMaskd("image", r=r, inside_off=False, prob=1.)
    t2 = RF.RandPlaneWaves_ellipsoid("image", a, b, c, intensity_value=6.0, prob=1.); t2.ellipsoid.set_random_state(seed=es)
    t3 = RF.WrapArtifactd("image", 0.5)
    t4 = RF.SaltAndPepper(0.05)
    d1 = t1({"image": x}); d2 = t2(d1); d3 = t3(d2)
    torch.manual_seed(17); u = torch.rand(x.size()); torch.manual_seed(17)
    d4 = t4(d3)
    idx = tuple(int(v) for v in t2.idx)
    same(P.chain_127(x, r, idx, 6.0, 0.5, 0.05, u), d4["image"], "chain127")
    f = [i - n // 2 for i, n in zip(idx, shape[1:])]
    save(f"chain127_{sname}_r{r}", dict(kind="chain127", r=r, idx=list(idx), intensity=6.0, alpha=0.5, p=0.05,
                                       spike_in_ball=bool(sum(v * v for v in f) < r * r)),
         x=x, u=u, y1=d1["image"], y2=d2["image"], y3=d3["image"], y4=d4["image"])

# ------------------------------------------------------------------ FFT KATs from the reference notebook
# fourier_images_disk_masks.ipynb cells 7, 8, 12
k = torch.fft.fftn(torch.ones(3, 3))
assert abs(k[0, 0] - 9) < 1e-6 and k.abs().sum() - 9 < 1e-5
tile = torch.tensor([1., 1, 0, 0, 1, 1, 0, 0]).repeat(8, 1)
k = torch.fft.fftn(tile)
assert abs(k[0, 0] - 32) < 1e-4 and abs(k[0, 2] - (16 - 16j)) < 1e-4 and abs(k[0, 6] - (16 + 16j)) < 1e-4

with open(os.path.join(OUT, "INDEX.json"), "w") as f:
    json.dump(CASES, f, indent=0, sort_keys=True)
print(f"wrote {len(CASES)} cases to {OUT}")
