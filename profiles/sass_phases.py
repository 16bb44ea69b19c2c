"""Where a kernel's time goes, from the per-instruction warp-stall samples of an `ncu --set full --import-source on`
report: opcode mix (executed instructions and samples) and the kernel's SASS cut into equal slices of static
instructions, each with its executed instructions, FP64 share, stall samples and their reasons.

    python profiles/sass_phases.py gpurun_out/r02_pair_n4.ncu-rep [slices] > profiles/r02_phases_pair_n4.md

The slices follow program order, so for the straight-line pair kernels they map onto the phases of a pair
(gather / pack, forward, Jacobi loop, metric tail, backward, stores).  samples/kinst is the inverse efficiency of
a slice: at FP64-pipe saturation with two warps per scheduler a slice of pure FP64 code sits at ~0.115."""
import collections
import csv
import io
import subprocess
import sys

STALLS = ["stall_wait", "stall_math", "stall_short_sb", "stall_long_sb", "stall_no_inst", "stall_barrier", "stall_not_selected",
          "stall_selected", "stall_mio", "stall_lg", "stall_dispatch", "stall_branch_resolving"]


def main(path, nseg=40):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    name = rows[0][1] if rows and len(rows[0]) > 1 else "?"
    hdr, data = rows[1], rows[2:]
    ci = {h: i for i, h in enumerate(hdr)}
    recs, ops, samp = [], collections.Counter(), collections.Counter()
    for r in data:
        if len(r) < len(hdr):
            continue
        toks = r[ci["Source"]].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        ex, s = int(r[ci["Instructions Executed"]]), int(r[ci["# Samples"]])
        ops[op.split(".")[0]] += ex
        samp[op.split(".")[0]] += s
        recs.append((op, ex, s, r))
    tot_i, tot_s = sum(ops.values()), sum(samp.values())
    print(f"# {name}\n\n{tot_i} warp instructions executed, {tot_s} stall samples, {len(recs)} static instructions\n")
    print("| opcode | executed | share | samples | share |\n|---|---|---|---|---|")
    for k, v in ops.most_common(16):
        print(f"| {k} | {v} | {100 * v / tot_i:.2f}% | {samp[k]} | {100 * samp[k] / max(tot_s, 1):.2f}% |")
    print("\n| slice | static range | executed (k) | FP64 % | samples | samples/kinst | " + " | ".join(s[6:] for s in STALLS) + " |")
    print("|" + "---|" * (6 + len(STALLS)))
    n = len(recs)
    for g in range(nseg):
        a, b = g * n // nseg, (g + 1) * n // nseg
        ex = sum(x[1] for x in recs[a:b])
        if ex == 0:
            continue
        s = sum(x[2] for x in recs[a:b])
        fp = sum(x[1] for x in recs[a:b] if x[0][0] == "D" and not x[0].startswith("DEPBAR"))
        st = [sum(int(x[3][ci[k]]) for x in recs[a:b]) for k in STALLS]
        print(f"| {g} | {a}-{b} | {ex / 1e3:.0f} | {100 * fp / ex:.1f} | {s} | {1e3 * s / ex:.2f} | " + " | ".join(str(v) for v in st) + " |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
