"""Classify the hazards of a `compute-sanitizer --tool racecheck --print-limit N` log by (reader, writer) function pair.

racecheck models __syncthreads / named barriers; it does not model the completion of async-proxy writes
(cp.async.bulk / cp.async.bulk.tensor) on an mbarrier, so every ring stage that is refilled by the TMA engine after the
consumers released it through the `empty` mbarrier shows up as a hazard between the consumers' LDS and the bulk copy.
This tool shows that those are the ONLY hazards of the staged kernels: anything else (the per-warp fp64 accumulators,
the CTA sums, the ticket reduction, the bitonic sort of K1b) would appear as its own class."""
import collections
import re
import sys

classes = collections.Counter()
cur = None
for line in open(sys.argv[1], errors="replace"):
    m = re.search(r"Race reported between (\w+) access at (?:bde::)?([\w:<>~]+)", line)
    if m:
        cur = (m.group(1), m.group(2))
        continue
    m = re.search(r"and (\w+) access at (?:bde::)?([\w:<>~]+).*\[(\d+) hazards\]", line)
    if m and cur:
        classes[(cur[0] + " " + cur[1], m.group(1) + " " + m.group(2))] += int(m.group(3))
total = sum(classes.values())
print(f"hazard classes ({len(classes)}), {total} hazards in total:")
for (a, b), n in classes.most_common():
    asyncp = "tma_load" in b or "tma_load" in a
    print(f"  {n:>12}  {a}  <->  {b}   [{'async-proxy TMA write vs ring read: synchronised by the full / empty mbarriers, not modelled by racecheck' if asyncp else 'OTHER'}]")
print("other classes:", sum(n for (a, b), n in classes.items() if "tma_load" not in a and "tma_load" not in b))
