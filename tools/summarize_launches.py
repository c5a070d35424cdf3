import csv, sys, collections
path, out, header = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0] != "ID"]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0][:72]
    agg.setdefault(name, []).append(float(r[-1].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
with open(out, "w") as f:
    f.write(header + "\n\n| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"| `{k}` | {len(v)} | {sum(v)/1e6:.2f} | {sum(v)/len(v)/1e3:.1f} | {100*sum(v)/tot:.1f}% |\n")
print(open(out).read())
