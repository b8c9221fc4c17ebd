"""Golden vectors for RandZF from the UNMODIFIED reference (50_reconstruction/reconGan/utils2.py, imported through
oracle/monai_shim).  Run in the build container (needs /root/reference):  python oracle/make_golden_recon.py
Writes tests/golden/randzf_*.npz: x, the uniform field the reference drew (torch.manual_seed before the call, then the
same draw repeated), p, and the reference's output.  Asserts the restatement oracle/recon_port.rand_zf is bit-identical."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "monai_shim"))
sys.path.insert(0, "/root/reference/50_reconstruction/reconGan")

import utils2 as REF  # noqa: E402
from oracle import recon_port as RP  # noqa: E402

CASES = [((3, 16, 12), 0.2, 1), ((1, 32, 32), 0.5, 2), ((2, 9, 15), 0.0, 3), ((2, 12, 10, 6), 0.35, 4), ((1, 8, 6, 31), 1.0, 5)]

for shape, p, seed in CASES:
    g = torch.Generator().manual_seed(100 + seed)
    x = torch.randn(*shape, generator=g)
    torch.manual_seed(seed)
    y = REF.RandZF(p)(x)
    torch.manual_seed(seed)
    u = torch.rand(x.size())
    assert torch.equal(RP.rand_zf(x, p, u), y), shape
    name = "randzf_s" + "x".join(str(s) for s in shape) + f"_p{p}"
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), x=x.numpy(), u=u.numpy(), y=y.numpy(),
                        meta=json.dumps({"p": p, "seed": seed, "shape": list(shape)}))
    print(name, float(y.abs().mean()))
# the frequency-consistency loss on a fixed pair
g = torch.Generator().manual_seed(77)
a, b = torch.randn(4, 1, 16, 12, generator=g), torch.randn(4, 1, 16, 12, generator=g)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "freqloss_s4x1x16x12.npz"), a=a.numpy(), b=b.numpy(),
                    loss=np.float64(RP.freq_consistency(a, b)), meta=json.dumps({"source": "reconGan_freq.py:134-140 restated"}))
