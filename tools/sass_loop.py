#!/usr/bin/env python
"""Static look at a kernel's SASS (no GPU needed): instruction mix of an address range.

    python tools/sass_loop.py <file.sass> <mangled-name-substring> [--from 0x1170 --to 0x1a20] [--list]

`make -C simplemoc-kernel_b200 sass` writes build/smk_api.sass.  Used to count the instructions per
segment of the hot loops (FFMA2 / IMAD / SHFL / LDG ...) before spending GPU time.
"""
import argparse
import collections
import re


def kernel_body(path, needle):
    lines = open(path).read().split("\n")
    starts = [i for i, l in enumerate(lines) if "Function :" in l]
    for si, s in enumerate(starts):
        if needle in lines[s]:
            e = starts[si + 1] if si + 1 < len(starts) else len(lines)
            out = []
            for l in lines[s:e]:
                m = re.search(r"/\*([0-9a-f]{4})\*/\s+(.*?);", l)
                if m:
                    out.append((int(m.group(1), 16), m.group(2).strip()))
            return lines[s].split(":")[1].strip(), out
    raise SystemExit(f"no kernel matching {needle}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("sass")
    ap.add_argument("needle")
    ap.add_argument("--from", dest="lo", default="0")
    ap.add_argument("--to", dest="hi", default="0xffffff")
    ap.add_argument("--list", action="store_true")
    a = ap.parse_args()
    name, body = kernel_body(a.sass, a.needle)
    lo, hi = int(a.lo, 16), int(a.hi, 16)
    sel = [(ad, ins) for ad, ins in body if lo <= ad <= hi]
    print(f"{name}: {len(body)} instructions, {len(sel)} in [{lo:#x}, {hi:#x}]")
    if a.list:
        for ad, ins in sel:
            print(f"{ad:04x} {ins}")
        return
    mix = collections.Counter()
    for _, ins in sel:
        op = re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0].split(".")[0]
        mix[op] += 1
    for op, n in mix.most_common():
        print(f"{n:5d} {op}")


if __name__ == "__main__":
    main()
