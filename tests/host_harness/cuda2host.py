"""TEST INFRASTRUCTURE -- turns the CUDA sources of warpx_b200/csrc into something g++ can compile against the SIMT
emulator (simt_host.h), WITHOUT touching the product files: copies of every .cu / .cuh go to a build directory with

  * `kernel<<<grid, block, smem, stream>>>(args)`  ->  `SIMT_LAUNCH(grid, block, smem, stream, kernel(args))`
  * `#include <cub/...>` dropped (simt_host.h carries a host stand-in for the one CUB call of the sort).

Everything else -- kernel bodies, host-side argument builders, the C++ step driver -- is compiled as written."""
import os
import re


def _match_back_angle(text, i):
    """text[i] == '>': index of the matching '<' (template arguments of the kernel name)."""
    depth = 0
    while i >= 0:
        if text[i] == '>':
            depth += 1
        elif text[i] == '<':
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced template arguments before <<<")


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def transform(text):
    text = re.sub(r'^[ \t]*#include\s*<cub/[^>]*>[^\n]*\n', '', text, flags=re.M)
    out, pos = "", 0
    while True:
        k = text.find("<<<", pos)
        if k < 0:
            return out + text[pos:]
        # kernel expression: identifier [ <template args> ] directly before <<<
        j = k - 1
        while text[j].isspace():
            j -= 1
        if text[j] == '>':
            j = _match_back_angle(text, j) - 1
        while j >= 0 and (text[j].isalnum() or text[j] in "_:"):
            j -= 1
        start = j + 1
        kernel = text[start:k].strip()
        end = text.index(">>>", k)
        cfg = _split_top(text[k + 3:end])
        cfg += ["0"] * (4 - len(cfg))
        a = end + 3
        while text[a].isspace() or text[a] == '\\':
            a += 1
        assert text[a] == '(', "launch without an argument list: " + text[k - 40:k + 80]
        depth, b = 0, a
        while True:
            if text[b] == '(':
                depth += 1
            elif text[b] == ')':
                depth -= 1
                if depth == 0:
                    break
            b += 1
        args = text[a + 1:b]
        out += text[pos:start] + "SIMT_LAUNCH(%s, %s, %s, %s, %s(%s))" % (cfg[0], cfg[1], cfg[2], cfg[3], kernel, args)
        pos = b + 1


def transform_tree(src_dir, dst_dir):
    """Transformed copies of src_dir/*.cu(h) in dst_dir; returns the .cpp paths.  The relative include of the C ABI
    header is made absolute (the copies live elsewhere)."""
    os.makedirs(dst_dir, exist_ok=True)
    header = os.path.abspath(os.path.join(src_dir, "..", "..", "include", "pic_b200.h"))
    names = sorted(f for f in os.listdir(src_dir) if f.endswith((".cu", ".cuh")))
    for f in names:
        with open(os.path.join(src_dir, f)) as fh:
            t = transform(fh.read()).replace('#include "../../include/pic_b200.h"', '#include "%s"' % header)
        path = os.path.join(dst_dir, f + (".cpp" if f.endswith(".cu") else ""))
        if not os.path.exists(path) or open(path).read() != t:
            with open(path, "w") as fh:
                fh.write(t)
    return [os.path.join(dst_dir, f + ".cpp") for f in names if f.endswith(".cu")]
