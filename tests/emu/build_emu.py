"""Builds tests/emu/_build/libmvster_emu.so: the SIMT sources of mvster_b200/csrc, rewritten by transform.py, compiled
with g++ against simt_emu.h.  The tensor-core (tcgen05 / TMA) entry points are stubs that fail."""
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REPO = HERE.parents[1]
CSRC = REPO / "mvster_b200" / "csrc"
sys.path.insert(0, str(HERE))
from transform import build_tree  # noqa: E402

SOURCES = ["common.cu", "hypo_head.cu", "fusion.cu", "et_fuse.cu", "et_fuse_bwd.cu", "conv_simt.cu", "conv_simt_px2.cu", "fpn.cu", "sinkhorn.cu"]
HEADERS = ["et_args.cuh", "et_fuse_tiled.cuh", "et_fuse_win.cuh", "et_fuse_tma.cuh", "conv_tc3_plan.h"]
FLAGS = ["-std=c++20", "-O1", "-ffp-contract=off", "-pthread", "-fPIC", "-w", "-x", "c++"]


def build(force: bool = False) -> Path:
    # MVSTER_EMU_ASAN=1: AddressSanitizer build in its own directory (run pytest with LD_PRELOAD=$(gcc -print-file-name=libasan.so)
    # and ASAN_OPTIONS=detect_leaks=0: tensors then come from the intercepted allocator and every out-of-bounds access of a
    # kernel is reported)
    asan = os.environ.get("MVSTER_EMU_ASAN", "0") == "1"
    tsan = os.environ.get("MVSTER_EMU_TSAN", "0") == "1"   # ThreadSanitizer: races between the threads of a block (LD_PRELOAD libtsan.so)
    out = HERE / ("_build_asan" if asan else "_build_tsan" if tsan else "_build")
    src = out / "src"
    lib = out / "libmvster_emu.so"
    deps = [CSRC / f for f in SOURCES + HEADERS] + [HERE / f for f in ("simt_emu.h", "common_emu.h", "transform.py", "build_emu.py", "emu_stubs.cpp")]
    if not force and lib.exists() and lib.stat().st_mtime > max(d.stat().st_mtime for d in deps):
        return lib
    build_tree(CSRC, src, SOURCES + HEADERS)
    for h in ("simt_emu.h", "common_emu.h"):
        (src / h).write_text((HERE / h).read_text())
    objs, procs = [], []
    for name in SOURCES + ["emu_stubs.cpp"]:
        path = src / name if name != "emu_stubs.cpp" else HERE / name
        obj = out / (Path(name).stem + ".o")
        procs.append((name, subprocess.Popen(["g++", *FLAGS, *(["-fsanitize=address", "-g", "-fno-omit-frame-pointer"] if asan else ["-fsanitize=thread", "-g"] if tsan else []),
                                              "-I", str(src), "-c", str(path), "-o", str(obj)], stdout=subprocess.PIPE,
                                             stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    bad = []
    for name, p in procs:
        log, _ = p.communicate()
        if p.returncode:
            bad.append(f"--- {name}\n{log[-6000:]}")
    if bad:
        raise RuntimeError("emulation build failed:\n" + "\n".join(bad))
    subprocess.check_call(["g++", "-shared", "-pthread", *(["-fsanitize=address"] if asan else ["-fsanitize=thread"] if tsan else []), "-o", str(lib), *map(str, objs)])
    return lib


if __name__ == "__main__":
    print(build(force=True))
