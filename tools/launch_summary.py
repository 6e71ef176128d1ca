"""Per-kernel shares of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file x.csv ...`).
Usage: python tools/launch_summary.py x.csv [--first-of-step im2col_kernel] [--steps 2]
Takes the launches of the LAST `--steps` steps (a step starts at a launch of --first-of-step) and prints a markdown
table of launches, total ms and share per kernel."""
import argparse, collections, csv, re, sys


def short(name):
    m = re.search(r"(?:<unnamed>::)?(\w+)(<[^(]*>)?\(", name)
    if not m:
        return name[:60]
    base, targs = m.group(1), m.group(2) or ""
    if base.startswith("vectorized_elementwise_kernel") or "at::" in name:
        return "torch: " + base
    return base + targs.replace("(bool)", "").replace("(int)", "")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv"); ap.add_argument("--first-of-step", default="im2col_kernel"); ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--skip-steps", type=int, default=-1, help="steps (warm-up) before the ones to summarise; -1 = take the last ones")
    a = ap.parse_args()
    rows = []
    with open(a.csv, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((int(r["ID"]), r["Kernel Name"], float(r["Metric Value"].replace(",", "")) / 1e6))
    starts = [i for i, r in enumerate(rows) if a.first_of_step in r[1]]
    if a.skip_steps >= 0:
        lo, hi = starts[a.skip_steps], starts[a.skip_steps + a.steps]
    elif len(starts) < a.steps + 1:
        lo, hi = (starts[-a.steps] if len(starts) >= a.steps else 0), len(rows)
    else:
        lo, hi = starts[-a.steps - 1], starts[-1]   # whole steps only: up to the start of the last (possibly cut) one
    sel = rows[lo:hi]
    agg = collections.OrderedDict()
    for _, name, ms in sel:
        k = short(name)
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + ms)
    total = sum(t for _, t in agg.values())
    print(f"launches {rows[lo][0]}..{rows[hi - 1][0]} of the capture ({len(sel)} launches, {a.steps} steps)\n")
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {t:.3f} | {100 * t / total:.1f} % |")
    print(f"| total | {len(sel)} | {total:.3f} | |")


if __name__ == "__main__":
    main()
