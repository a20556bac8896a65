"""cu2cpp.py -- TEST INFRASTRUCTURE.  Rewrites the CUDA launch syntax of a .cu file so that g++ can compile it against
tests/host_emul/fullhost/cuda_runtime.h:  kernel<targs><<<grid, block, smem, stream>>>(args)  ->
SHIM_LAUNCH(grid, block, kernel<targs>(args)),  `extern __shared__` -> `extern`, `__shared__` -> `static`."""
import re
import sys


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    out.append(cur)
    return out


def convert(src):
    out, pos = "", 0
    pat = re.compile(r"([A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*<<<")
    while True:
        m = pat.search(src, pos)
        if not m:
            out += src[pos:]
            break
        end_cfg = src.index(">>>", m.end())
        cfg = split_top(src[m.end():end_cfg])
        k = end_cfg + 3
        while src[k] in " \t\\\n":
            k += 1
        assert src[k] == "(", src[m.start():k + 20]
        depth, j = 0, k
        while True:
            if src[j] == "(":
                depth += 1
            elif src[j] == ")":
                depth -= 1
                if depth == 0:
                    break
            j += 1
        args = src[k:j + 1]
        out += src[pos:m.start()] + "SHIM_LAUNCH(%s, %s, %s%s)" % (cfg[0].strip(), cfg[1].strip(), m.group(1), args)
        pos = j + 1
    out = out.replace("extern __shared__", "extern")
    out = re.sub(r"\b__shared__\b", "static", out)
    return out


if __name__ == "__main__":
    sys.stdout.write(convert(open(sys.argv[1]).read()))
