"""Maps the local-memory (spill) instructions of one kernel to source lines: python tools/spill_lines.py <obj> <mangled-substring>."""
import re, subprocess, sys, tempfile, os, glob
obj, pat = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, stdout=subprocess.DEVNULL)
dis = subprocess.run(["nvdisasm", "-g", "-c"] + glob.glob(d + "/*.cubin"), capture_output=True, text=True).stdout
cur = fn = None
out, total = {}, 0
for line in dis.splitlines():
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*\.text\.(\S+):', line)
    if m:
        fn = m.group(1)
        continue
    if fn and pat in fn and not line.strip().startswith("//"):
        if re.search(r'\b[A-Z][A-Z0-9_.]+\b', line): total += 1
        if re.search(r'\b(STL|LDL)', line): out[cur] = out.get(cur, 0) + 1
print("instructions", total, "local-memory instructions", sum(out.values()))
for k, v in sorted(out.items()): print(k, v)
