"""Source-to-source step of the CPU emulation (tests/emu): copies the SIMT sources of mvster_b200/csrc into a build
directory with the three CUDA-only constructs rewritten so that g++ accepts them -

    kernel<<<grid, block, smem, stream>>>(args);   ->  emu::launch(grid, block, smem, [&] { kernel(args); });
    extern __shared__ ... T name[];                ->  T* name = reinterpret_cast<T*>(emu::ctx.dyn_smem);
    #include "common.cuh"                          ->  #include "common_emu.h"   (CUDA runtime names as host stubs)

Everything else (kernel bodies, dispatch, argument checks, the extern "C" entry points) is compiled as it stands, so the
emulation library exports the same C ABI as libmvster_b200 for these files, taking host pointers."""
import re
from pathlib import Path

LAUNCH = re.compile(r"^(?P<lead>\s*(?:(?:else )?if \(.*?\) |else )?)(?P<kernel>[\w:]+(?:<[^;]*?>)?)<<<(?P<cfg>.*)>>>\((?P<args>.*)\);\s*$")
DYN_SMEM = re.compile(r"^(?P<lead>\s*)extern __shared__ (?:__align__\(\d+\) )?(?P<type>\w+) (?P<name>\w+)\[\];(?P<rest>.*)$")


def split_top(s: str):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "(<[":
            depth += 1
        elif ch in ")>]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def transform(text: str) -> str:
    lines = []
    for line in text.splitlines():
        m = LAUNCH.match(line)
        if m:
            cfg = split_top(m["cfg"])
            grid, block = cfg[0], cfg[1]
            smem = cfg[2] if len(cfg) > 2 else "0"
            line = f'{m["lead"]}emu::launch(dim3({grid}), dim3({block}), (size_t)({smem}), [&] {{ {m["kernel"]}({m["args"]}); }});'
        else:
            m = DYN_SMEM.match(line)
            if m:
                line = f'{m["lead"]}{m["type"]}* {m["name"]} = reinterpret_cast<{m["type"]}*>(emu::ctx.dyn_smem);{m["rest"]}'
        line = line.replace('#include "common.cuh"', '#include "common_emu.h"')
        assert "<<<" not in line, f"unhandled launch: {line}"
        lines.append(line)
    return "\n".join(lines) + "\n"


def build_tree(csrc: Path, out: Path, files):
    out.mkdir(parents=True, exist_ok=True)
    for name in files:
        (out / name).write_text(transform((csrc / name).read_text()))
