#!/usr/bin/env python
"""Hashes of the device code of libhfg.so, per kernel, independent of cuobjdump's column padding and of the mangled name:

    python tools/sass_hash.py [flagger_b200/libhfg.so]

Used to state which kernels were on the GPU when a measurement was taken: host-only changes must
leave these hashes alone.  Each hash covers addresses, instructions and encodings of one function."""
import hashlib
import re
import subprocess
import sys


def functions(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    for part in re.split(r"\n\t\tFunction : ", out)[1:]:
        name, body = part.split("\n", 1)
        lines = [re.sub(r"\s+", " ", ln).strip() for ln in body.splitlines()]
        lines = [ln for ln in lines if ln and not ln.startswith("...") and not ln.startswith("Fatbin") and not ln.startswith("=")
                 and not ln.startswith("arch =") and not ln.startswith("code version") and not ln.startswith("host =")
                 and not ln.startswith("compile_size") and not ln.startswith("code for")]
        yield name.strip(), hashlib.md5("\n".join(lines).encode()).hexdigest(), len(lines)


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else "flagger_b200/libhfg.so"
    for name, digest, n in functions(path):
        if "cub" in name:
            continue
        print(f"{digest}  {n:6d} lines  {name}")


if __name__ == "__main__":
    main()
