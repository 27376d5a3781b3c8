"""Per-kernel totals and shares of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv).
usage: python tools/launch_shares.py <launches.csv> [out.json]"""
import collections, csv, json, sys


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    ms, n = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("unnamed>::", "")
        ms[name] += float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[ui]]
        n[name] += 1
    tot = sum(ms.values())
    out = [{"kernel": k, "launches": n[k], "ms": round(v, 3), "share": round(v / tot, 4)} for k, v in ms.most_common()]
    for o in out:
        print("%10.2f ms %6.2f%%  x%4d  %s" % (o["ms"], 100 * o["share"], o["launches"], o["kernel"]))
    if len(sys.argv) > 2:
        json.dump({"source": sys.argv[1].split("/")[-1], "total_ms": round(tot, 3), "kernels": out}, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
