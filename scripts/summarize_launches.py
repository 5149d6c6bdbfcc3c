"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total device time, share."""
import collections, csv, re, sys

def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("tclip::<unnamed>::", "").replace("<unnamed>::", "").replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print(f"# {path}: {sum(cnt.values())} launches, {total / 1e6:.2f} ms of kernel time (cold-cache, serialised: compare shares)")
    print(f"{'kernel':62s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{k[:62]:62s} {cnt[k]:8d} {v / 1e6:10.3f} {v / total:7.3f} {v / cnt[k] / 1e3:9.1f}")

if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
        print()
