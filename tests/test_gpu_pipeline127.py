"""The training transform of the 127 scripts from the resampled volume on (127_...FLAIR.py:130-141), every stage a drop-in
running on the GPU, against the composition of the oracles with the same random states:

    RandSpatialCropd([128,128,64]) -> RandFlipd(0.5, axis 0) -> NormalizeIntensityd(nonzero, channel_wise)
    -> RandScaleIntensityd(0.1, 0.5) -> RandShiftIntensityd(0.1, 0.5) -> RandFourierDiskMaskd(12.5) -> WrapArtifactd(0.5)

(the plane-wave spike and SaltAndPepper stages have their own seeded tests in test_gpu_parity.py; here the chain stays
deterministic given the six states).  Image: rel-L2 <= 1e-5; label (crop + flip only): equal."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_train_transform_from_crop_to_wrap(cuda_device, seed):
    import filters_and_operators as F
    from mvtb import intensity as I, spatial as S
    from oracle import monai_intensity as MI, monai_spatial as MS, ref_port as P
    shape = (1, 176, 160, 90)
    base = P.synthetic_volume(40 + seed, shape).numpy()
    img = (np.abs(base) * 900.0 * (np.abs(base) > 0.2)).astype(np.float32)              # MR-like: zero background, positive tissue
    lab = (P.synthetic_volume(60 + seed, (3,) + shape[1:]).numpy() > 0.4).astype(np.float32)
    roi = [128, 128, 64]
    stages = [S.RandSpatialCropd(["image", "label"], roi_size=roi, random_size=False),
              S.RandFlipd(["image", "label"], prob=0.5, spatial_axis=0),
              I.NormalizeIntensityd("image", nonzero=True, channel_wise=True),
              I.RandScaleIntensityd("image", factors=0.1, prob=0.5),
              I.RandShiftIntensityd("image", offsets=0.1, prob=0.5),
              F.RandFourierDiskMaskd(keys="image", r=12.5, inside_off=False, prob=1.),
              F.WrapArtifactd("image", 0.5)]
    seeds = [1000 + 10 * seed + i for i in range(len(stages))]
    for tr, sd in zip(stages, seeds):
        if hasattr(tr, "set_random_state"):
            tr.set_random_state(seed=sd)
    d = {"image": torch.from_numpy(img).to(cuda_device), "label": torch.from_numpy(lab).to(cuda_device)}
    for tr in stages:
        d = tr(d)
    # the oracles with the same states
    Rc, Rf = np.random.RandomState(seeds[0]), np.random.RandomState(seeds[1])
    sl = MS.rand_spatial_crop_slices(shape[1:], roi, Rc)
    do_flip = Rf.rand() < 0.5
    xi = MS.crop_then_flip(img, sl, do_flip, 0)
    xl = MS.crop_then_flip(lab, sl, do_flip, 0)
    f, o = MI.draw_scale_shift(np.random.RandomState(seeds[3]), np.random.RandomState(seeds[4]))
    xi = MI.prologue(xi, f, o)
    y = P.wrap_artifact(P.fourier_disk_mask(torch.from_numpy(xi), 12.5, False), 0.5).numpy()
    assert tuple(d["image"].shape) == (1, 128, 128, 64) and d["image"].is_cuda
    assert torch.equal(d["label"].cpu(), torch.from_numpy(xl))
    assert rel_l2(d["image"].cpu().numpy(), y) <= TOL
