"""Static SASS size of every kernel in libhns_b200.so (no GPU needed): `python scripts/sass_counts.py [substring]`.
Used for the instruction-trimming notes in DESIGN.md / profiles/README.md (e.g. k_rbgs_split 232 -> 192 -> 176)."""
import collections, os, re, subprocess, sys

so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hnanosolver_b200", "libhns_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
name, counts, mix = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[name], mix[name] = 0, collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        counts[name] += 1
        mix[name][m.group(1).split(".")[0]] += 1
want = sys.argv[1] if len(sys.argv) > 1 else ""
for k, v in counts.items():
    if want in k:
        top = ", ".join(f"{op} {n}" for op, n in mix[k].most_common(6))
        print(f"{v:6d}  {k}   [{top}]")
