"""Per-kernel time shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python tools/launch_shares.py gpurun_out/launches.csv > profiles/r02_ncu_launch_shares_F30.txt
"""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    im = hdr.index("Metric Name")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        name = re.sub(r"^void |\(.*$", "", r[ik]).replace("sdw::", "")
        tot[name] += v
        cnt[name] += 1
    s = sum(tot.values())
    print(f"# total {s:.1f} ms over {sum(cnt.values())} launches\n")
    for k, v in tot.most_common():
        print(f"{100 * v / s:5.1f}%  {v:9.2f} ms  x{cnt[k]:4d}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
