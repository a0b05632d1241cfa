"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and launch count per kernel."""
import collections, csv, re, sys

def main(path, top=60):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0; n = 0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(row["Metric Unit"], 1.0)
        name = row["Kernel Name"]
        m = re.match(r"(?:void )?(?:sr::)?([\w:]+)", name)
        short = m.group(1) if m else name[:60]
        if short.startswith("at::") or "at::native" in name:
            f = re.search(r"at::native::(?:\(anonymous namespace\)::)?(\w+)", name)
            g = re.search(r"(\w+Functor|\w+_kernel_cuda|\w+Ops?)\b", name[name.find("<"):]) if "<" in name else None
            short = "torch:" + (f.group(1) if f else short) + ("<" + g.group(1) + ">" if g else "")
        agg[short][0] += 1; agg[short][1] += v; tot += v; n += 1
    print("launches %d, summed kernel time %.3f ms" % (n, tot / 1e6))
    ours = sum(v[1] for k, v in agg.items() if not k.startswith(("torch:", "cutlass", "gemm", "magma", "at::", "cub")))
    print("library (sr::) kernels: %.3f ms (%.1f%%)" % (ours / 1e6, 100 * ours / tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-78s %6d %9.3f ms %5.1f%%" % (k[:78], v[0], v[1] / 1e6, 100 * v[1] / tot))

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60)
