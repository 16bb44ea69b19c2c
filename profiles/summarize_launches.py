"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table."""
import collections
import csv
import sys


def main(path, title):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    # the FP64 peak probe (sympa_probe_fp64) is a diagnostic bench.py runs OUTSIDE its timed region
    probe = [r for r in rows if "fp64_probe_kernel" in r["Kernel Name"]]
    rows = [r for r in rows if "fp64_probe_kernel" not in r["Kernel Name"]]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r["Kernel Name"]
        if "pair_kernel" in name or "coop_kernel" in name or "scatter_kernel" in name or "scale_kernel" in name:
            name = name.split("(")[0]
        agg[name[:100]][0] += 1
        agg[name[:100]][1] += float(r["Metric Value"])
    tot = sum(v[1] for v in agg.values())
    print(f"# {title}")
    print("# per-launch times under ncu are cold-cache and serialised - compare shares, not absolutes\n")
    if probe:
        print(f"# excluded: {len(probe)} launches of fp64_probe_kernel (roofline denominator probe, outside the timed region)\n")
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        print(f"| `{k}` | {c} | {t / 1e3:.1f} | {100 * t / tot:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
