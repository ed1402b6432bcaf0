"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> markdown table of kernels by total time.
usage: python tools/summarize_launches.py launches.csv "title" "command" > profiles/<name>.md"""
import collections, csv, re, sys

path, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[ki].replace("rcn::<unnamed>::", "").replace("rcn::(anonymous namespace)::", ""))
    name = re.sub(r"^void ", "", name)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) / 1e6
tot = sum(a[1] for a in agg.values())
print(f"# {title}\n\nCommand (gpurun, 1x B200): `{cmd}`\n")
print(f"{len(rows)} captured launches, {tot:.2f} ms in total (cold-cache, serialised under the profiler: compare SHARES, not absolutes).\n")
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {a[0]} | {a[1]:.2f} | {100 * a[1] / tot:.1f}% |")
