"""Instruction mix of the kernels whose (mangled) name matches a regex: python tools/sass_stats.py <regex> [top]"""
import collections, re, subprocess, sys
from pathlib import Path

lib = Path(__file__).resolve().parent.parent / "beyond_deep_ensembles_b200" / "lib" / "libbde_b200.so"
pat = re.compile(sys.argv[1])
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
name, mix = None, None
def flush():
    if name and pat.search(name):
        tot = sum(mix.values())
        print(f"{name}: {tot} instructions;", ", ".join(f"{k} {v}" for k, v in mix.most_common(top)))
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        name, mix = m.group(1), collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and mix is not None:
        mix[m.group(1)] += 1
flush()
