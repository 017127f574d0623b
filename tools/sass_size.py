"""SASS instruction count of one kernel by source file / line bucket: python tools/sass_size.py <obj> <mangled-substring> [bucket]"""
import re, subprocess, sys, tempfile, os, glob, collections
obj, pat = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 25
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, stdout=subprocess.DEVNULL)
dis = subprocess.run(["nvdisasm", "-g", "-c"] + glob.glob(d + "/*.cubin"), capture_output=True, text=True).stdout
cur = fn = None
cnt = collections.Counter()
total = 0
for line in dis.splitlines():
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)) // bucket * bucket)
        continue
    m = re.match(r'\s*\.text\.(\S+):', line)
    if m:
        fn = m.group(1)
        continue
    if fn and pat in fn and re.match(r'\s+/\*[0-9a-f]{4,}\*/', line):
        cnt[cur] += 1
        total += 1
print("instructions", total)
for k, v in sorted(cnt.items(), key=lambda x: -x[1])[:45]:
    print(f"{v:6d}  {k}")
