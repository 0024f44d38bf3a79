"""Writes tests/golden/formats/* with the UNMODIFIED reference (datasets/data_io.py) - run in the build container:
    python tests/golden/make_format_fixtures.py
The fixtures pin mvster_b200/formats.py against the reference's own writer / reader byte for byte."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, "/root/reference")
from datasets.data_io import read_pfm, save_pfm  # noqa: E402

out = Path(__file__).resolve().parent / "formats"
out.mkdir(exist_ok=True)
rng = np.random.RandomState(7)
gray = (rng.rand(6, 9).astype(np.float32) * 500 + 425)
gray[2, 3] = 0.0
color = rng.randn(4, 5, 3).astype(np.float32)
save_pfm(str(out / "depth_6x9.pfm"), gray)
save_pfm(str(out / "color_4x5.pfm"), color, scale=2)
np.save(out / "depth_6x9.npy", gray)
np.save(out / "color_4x5.npy", color)
# a big-endian file the reference reads (written by hand in the layout of data_io.py)
be = rng.rand(3, 4).astype(np.float32)
with open(out / "big_endian_3x4.pfm", "wb") as f:
    f.write(b"Pf\n4 3\n1.000000\n")
    f.write(np.flipud(be).astype(">f4").tobytes())
d, s = read_pfm(str(out / "big_endian_3x4.pfm"))
assert s == 1.0 and np.array_equal(np.asarray(d, np.float32), be)
np.save(out / "big_endian_3x4.npy", be)
(out / "00000000_cam.txt").write_text(
    "extrinsic\n0.970263 0.00747983 0.241939 -191.02\n-0.0147429 0.999493 0.0282234 3.28832\n"
    "-0.241605 -0.030951 0.969881 22.5401\n0.0 0.0 0.0 1.0\n\nintrinsic\n2892.33 0 823.205\n0 2883.18 619.071\n0 0 1\n\n425 2.5\n")
(out / "00000001_cam.txt").write_text(
    "extrinsic\n1 0 0 -10.5\n0 1 0 3\n0 0 1 0.25\n0.0 0.0 0.0 1.0\n\nintrinsic\n1000.5 0 400\n0 1000.25 300\n0 0 1\n\n2.5 0.01 256 7.62\n")
(out / "pair.txt").write_text("3\n0\n4 10 2346.41 1 2036.53 9 1243.89 12 1052.87\n1\n0\n2\n2 0 10.0 1 9.0\n")
print("written", sorted(p.name for p in out.iterdir()))

# camera files / pair list parsed by the reference's own reader (general_eval4.MVSDataset.read_cam_file needs only self.ndepths)
import json
import types

try:
    from datasets.general_eval4 import MVSDataset
    expected = {}
    for name, scale, nd in (("00000000_cam.txt", 1.06, 192), ("00000001_cam.txt", 1.0, 192), ("00000001_cam.txt", 0.8, 96)):
        me = types.SimpleNamespace(ndepths=nd)
        k, e, dmin, itv = MVSDataset.read_cam_file(me, str(out / name), scale)
        expected[f"{name}|{scale}|{nd}"] = {"K": k.tolist(), "E": e.tolist(), "depth_min": dmin, "depth_interval": itv}
    (out / "cam_expected.json").write_text(json.dumps(expected, indent=1))
    print("cam_expected.json written")
except Exception as ex:  # cv2 / PIL missing would land here
    print("reference cam reader unavailable:", ex)
