"""On-disk formats either side of the forward path (SURVEY.md 8f "next" #4): PFM depth / confidence maps, MVSNet camera
files, pair lists, and the per-stage projection stack the forward consumes.

Byte-compatible with the reference's readers and writers (``datasets/data_io.py:6-71``, ``datasets/general_eval4.py:25-79,
155-183``): a file written here is read back identically by the reference and vice versa (tests/test_formats.py pins both
directions against fixtures written by the unmodified reference).  Pure numpy, no GPU.
"""
from __future__ import annotations

import re
import sys
from typing import Dict, List, Sequence, Tuple

import numpy as np


def read_pfm(filename: str) -> Tuple[np.ndarray, float]:
    """``datasets/data_io.py:6-40``.  Returns (image [H,W] or [H,W,3] float32 in the FILE's byte order, top row first; scale).
    Header: ``PF``/``Pf``, ``width height``, signed scale (negative = little-endian); rows are stored bottom-up."""
    with open(filename, "rb") as f:
        header = f.readline().decode("utf-8").rstrip()
        if header == "PF":
            color = True
        elif header == "Pf":
            color = False
        else:
            raise Exception("Not a PFM file.")
        m = re.match(r"^(\d+)\s(\d+)\s$", f.readline().decode("utf-8"))
        if not m:
            raise Exception("Malformed PFM header.")
        width, height = map(int, m.groups())
        scale = float(f.readline().rstrip())
        endian = "<" if scale < 0 else ">"
        scale = abs(scale)
        data = np.frombuffer(f.read(), dtype=endian + "f4").copy()  # writable, like the reference's np.fromfile
    shape = (height, width, 3) if color else (height, width)
    if data.size != int(np.prod(shape)):
        raise Exception(f"PFM payload holds {data.size} floats, header says {shape}")
    return np.flipud(data.reshape(shape)), scale


def save_pfm(filename: str, image: np.ndarray, scale: float = 1) -> None:
    """``datasets/data_io.py:43-71``: float32 [H,W], [H,W,1] or [H,W,3]; rows written bottom-up, scale sign = byte order."""
    if image.dtype.name != "float32":
        raise Exception("Image dtype must be float32.")
    if image.ndim == 3 and image.shape[2] == 3:
        color = True
    elif image.ndim == 2 or (image.ndim == 3 and image.shape[2] == 1):
        color = False
    else:
        raise Exception("Image must have H x W x 3, H x W x 1 or H x W dimensions.")
    endian = image.dtype.byteorder
    if endian == "<" or (endian == "=" and sys.byteorder == "little"):
        scale = -scale
    with open(filename, "wb") as f:
        f.write(b"PF\n" if color else b"Pf\n")
        f.write(f"{image.shape[1]} {image.shape[0]}\n".encode("utf-8"))
        f.write(("%f\n" % scale).encode("utf-8"))
        f.write(np.ascontiguousarray(np.flipud(image)).tobytes())


def read_cam_file(filename: str, interval_scale: float = 1.0, ndepths: int = 192):
    """``datasets/general_eval4.py:59-79``: (intrinsics [3,3] already divided by 4 in its first two rows, extrinsics [4,4],
    depth_min, depth_interval).  With a third number on the depth line the interval is re-derived from it for ``ndepths``."""
    with open(filename) as f:
        lines = [line.rstrip() for line in f.readlines()]
    extrinsics = np.array(" ".join(lines[1:5]).split(), dtype=np.float32).reshape(4, 4)
    intrinsics = np.array(" ".join(lines[7:10]).split(), dtype=np.float32).reshape(3, 3)
    intrinsics[:2, :] /= 4.0
    tokens = lines[11].split()
    depth_min, depth_interval = float(tokens[0]), float(tokens[1])
    if len(tokens) >= 3:
        depth_max = depth_min + int(float(tokens[2])) * depth_interval
        depth_interval = (depth_max - depth_min) / ndepths
    depth_interval *= interval_scale
    return intrinsics, extrinsics, depth_min, depth_interval


def write_cam_file(filename: str, extrinsics: np.ndarray, intrinsics: np.ndarray, depth_min: float, depth_interval: float) -> None:
    """The MVSNet ``*_cam.txt`` layout ``read_cam_file`` parses (``intrinsics`` at FULL scale: the reader divides by 4)."""
    with open(filename, "w") as f:
        f.write("extrinsic\n")
        for r in np.asarray(extrinsics, dtype=np.float64).reshape(4, 4):
            f.write(" ".join(repr(float(x)) for x in r) + " \n")
        f.write("\nintrinsic\n")
        for r in np.asarray(intrinsics, dtype=np.float64).reshape(3, 3):
            f.write(" ".join(repr(float(x)) for x in r) + " \n")
        f.write(f"\n{float(depth_min)!r} {float(depth_interval)!r} \n")


def read_pair_file(filename: str, nviews: int) -> List[Tuple[int, List[int]]]:
    """``datasets/general_eval4.py:36-52``: [(reference view, source views)], views without sources dropped, short lists padded
    with their first source view up to ``nviews``."""
    metas = []
    with open(filename) as f:
        n = int(f.readline())
        for _ in range(n):
            ref = int(f.readline().rstrip())
            src = [int(x) for x in f.readline().rstrip().split()[1::2]]
            if src:
                if len(src) < nviews:
                    src = src + [src[0]] * (nviews - len(src))
                metas.append((ref, src))
    return metas


def stage_projections(extrinsics: Sequence[np.ndarray], intrinsics: Sequence[np.ndarray]) -> Dict[str, np.ndarray]:
    """``datasets/general_eval4.py:155-183``: per view ``[2,4,4]`` (slot 0 extrinsic, slot 1 intrinsic in the top-left 3x3) at
    the stage-2 (quarter) scale the cam reader delivers; stage 1 halves, stages 3 / 4 double / quadruple intrinsic rows 0-1."""
    mats = np.zeros((len(extrinsics), 2, 4, 4), np.float32)
    for v, (e, k) in enumerate(zip(extrinsics, intrinsics)):
        mats[v, 0, :4, :4] = e
        mats[v, 1, :3, :3] = k
    out = {}
    for name, s in (("stage1", 0.5), ("stage2", 1.0), ("stage3", 2.0), ("stage4", 4.0)):
        m = mats.copy()
        if s == 0.5:
            m[:, 1, :2, :] = mats[:, 1, :2, :] / 2.0
        elif s != 1.0:
            m[:, 1, :2, :] = mats[:, 1, :2, :] * s
        out[name] = m
    return out


def depth_values(depth_min: float, depth_interval: float, ndepths: int = 192) -> np.ndarray:
    """``datasets/general_eval4.py:164-165``: the evaluation loaders' hypothesis list (the forward reads its first and last entry)."""
    return np.arange(depth_min, depth_interval * (ndepths - 0.5) + depth_min, depth_interval, dtype=np.float32)


def read_img(filename: str) -> np.ndarray:
    """``datasets/general_eval4.py:81-86``: 8-bit image -> float32 in [0, 1], [H,W,3]."""
    from PIL import Image
    return np.array(Image.open(filename), dtype=np.float32) / 255.0


def scale_mvs_input(img: np.ndarray, intrinsics: np.ndarray, max_w: int, max_h: int, base: int = 64):
    """``datasets/general_eval4.py:92-109``: shrink to fit (max_h, max_w) if necessary, then cut both sides down to multiples of
    ``base`` by RESIZING (not cropping), scaling the intrinsic rows with it.  ``intrinsics`` is modified in place like the reference."""
    import cv2
    h, w = img.shape[:2]
    if h > max_h or w > max_w:
        scale = 1.0 * max_h / h
        if scale * w > max_w:
            scale = 1.0 * max_w / w
        new_w, new_h = scale * w // base * base, scale * h // base * base
    else:
        new_w, new_h = 1.0 * w // base * base, 1.0 * h // base * base
    intrinsics[0, :] *= 1.0 * new_w / w
    intrinsics[1, :] *= 1.0 * new_h / h
    return cv2.resize(img, (int(new_w), int(new_h))), intrinsics


def load_eval_sample(datapath: str, scan: str, ref_view: int, src_views: Sequence[int], nviews: int, interval_scale: float = 1.06,
                     max_h: int = 1200, max_w: int = 1600, ndepths: int = 192) -> Dict:
    """One evaluation sample as ``MVSDataset.__getitem__`` builds it (``datasets/general_eval4.py:111-188``, ``fix_res=False``):
    ``imgs`` list of nviews [3,H,W] float32, ``proj_matrices`` {stage1..4: [nviews,2,4,4]}, ``depth_values`` [ndepths],
    ``filename`` pattern.  Every view is resized to the (multiple-of-64) size of the reference view."""
    import os
    import cv2
    view_ids = [ref_view] + list(src_views)[:nviews - 1]
    imgs, ext, intr, dv = [], [], [], None
    s_h = s_w = 0
    for i, vid in enumerate(view_ids):
        name = os.path.join(datapath, "{}/images_post/{:0>8}.jpg".format(scan, vid))
        if not os.path.exists(name):
            name = os.path.join(datapath, "{}/images/{:0>8}.jpg".format(scan, vid))
        img = read_img(name)
        k, e, depth_min, depth_interval = read_cam_file(os.path.join(datapath, "{}/cams/{:0>8}_cam.txt".format(scan, vid)), interval_scale, ndepths)
        img, k = scale_mvs_input(img, k, max_w, max_h)
        if i == 0:
            s_h, s_w = img.shape[:2]
        c_h, c_w = img.shape[:2]
        if c_h != s_h or c_w != s_w:
            img = cv2.resize(img, (s_w, s_h))
            k[0, :] *= 1.0 * s_w / c_w
            k[1, :] *= 1.0 * s_h / c_h
        imgs.append(img.transpose(2, 0, 1))
        ext.append(e)
        intr.append(k)
        if i == 0:
            dv = depth_values(depth_min, depth_interval, ndepths)
    return {"imgs": imgs, "proj_matrices": stage_projections(ext, intr), "depth_values": dv,
            "filename": scan + "/{}/" + "{:0>8}".format(view_ids[0]) + "{}"}


def save_ply(filename: str, xyz: np.ndarray, rgb: np.ndarray = None) -> None:
    """Point cloud as the reference's ``filter_depth`` writes it through ``plyfile`` (``test_mvs4.py:402-414``): one ``vertex``
    element with float32 ``x y z`` and, with colours, uint8 ``red green blue``; binary, native (little-endian) byte order."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    fields = [("x", "<f4"), ("y", "<f4"), ("z", "<f4")]
    if rgb is not None:
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8).reshape(-1, 3)
        if len(rgb) != len(xyz):
            raise ValueError(f"{len(xyz)} points but {len(rgb)} colours")
        fields += [("red", "u1"), ("green", "u1"), ("blue", "u1")]
    rec = np.empty(len(xyz), dtype=fields)
    rec["x"], rec["y"], rec["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    if rgb is not None:
        rec["red"], rec["green"], rec["blue"] = rgb[:, 0], rgb[:, 1], rgb[:, 2]
    names = {"<f4": "float", "u1": "uchar"}
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % len(xyz)
    header += "".join("property %s %s\n" % (names[t], n) for n, t in fields) + "end_header\n"
    with open(filename, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(rec.tobytes())


def read_ply(filename: str):
    """Reads back what ``save_ply`` (or plyfile, for the same vertex layout) wrote: (xyz [N,3] float32, rgb [N,3] uint8 or None)."""
    with open(filename, "rb") as f:
        if f.readline().strip() != b"ply":
            raise Exception("Not a PLY file.")
        fmt, n, props = None, 0, []
        while True:
            line = f.readline().decode("ascii").strip()
            if line == "end_header":
                break
            tok = line.split()
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                if tok[1] != "vertex":
                    raise Exception("only a single vertex element is supported")
                n = int(tok[2])
            elif tok[0] == "property":
                props.append((tok[2], {"float": "f4", "float32": "f4", "uchar": "u1", "uint8": "u1"}[tok[1]]))
        if fmt not in ("binary_little_endian", "binary_big_endian"):
            raise Exception(f"unsupported PLY format {fmt}")
        order = "<" if fmt == "binary_little_endian" else ">"
        rec = np.frombuffer(f.read(), dtype=[(nm, order + t if t != "u1" else t) for nm, t in props], count=n)
    xyz = np.stack([rec["x"], rec["y"], rec["z"]], 1).astype(np.float32)
    rgb = np.stack([rec["red"], rec["green"], rec["blue"]], 1) if "red" in rec.dtype.names else None
    return xyz, rgb
