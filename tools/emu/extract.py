"""Pulls named definitions (kernels, device functions, structs, enums, device tables) out of a .cuh as text, rewritten for
tools/emu/cuda_emu.hpp: `extern __shared__ T name[];` -> EMU_DYN_SMEM(T, name).  Used by tests/test_kernel_emulation.py."""
import re


def _block_end(src, i):
    depth = 0
    while i < len(src):
        ch = src[i]
        if ch == "{":
            depth += 1
        elif ch == "}":
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced braces")


def extract(src, names):
    out = []
    for name in names:
        m = (re.search(r"^(?:template\s*<[^\n]*>\s*\n)?(?:__host__\s+)?(?:__global__|__device__)[^;{]*?\b%s\s*\(" % re.escape(name), src, re.M | re.S)
             or re.search(r"^(?:struct|enum)\s+%s\b[^;{]*\{" % re.escape(name), src, re.M)
             or re.search(r"^enum\s*\{[^}]*\b%s\b" % re.escape(name), src, re.M))
        if m is None:
            m = re.search(r"^(?:__device__|constexpr)[^;\n]*\b%s\b[^;\n]*;" % re.escape(name), src, re.M)       # device table / constant
            if m is None:
                raise KeyError(name)
            out.append(m.group(0))
            continue
        start = m.start()
        end = _block_end(src, src.index("{", m.start()))
        text = src[start:end]
        if text.lstrip().startswith(("struct", "enum")):
            text += ";"
        out.append(text)
    body = "#ifndef SSP\n#define SSP(...)      /* profiling build only (qrl_kernels.cuh, -DQRL_SS_PROF) */\n#endif\n\n" + "\n\n".join(out)
    body = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];", r"EMU_DYN_SMEM(\1, \2);", body)
    return body
