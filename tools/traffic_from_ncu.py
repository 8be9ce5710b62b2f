"""Development aid (runs here, no GPU): profiles/roofline_traffic.json from an ncu CSV log of
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file <csv> python tools/msd_probe.py 3000000000 4 msd
The LAST launch of every kernel is taken (the second, warm build).  bench.py reads the file for `roofline.traffic`."""
import collections
import csv
import json
import sys

STAGES = {  # build stage (library timer name) -> kernel name prefix
    "msd_local_sort": "msd_local_sort_kernel",
    "msd_part": "void msd_partition_kernel",
    "msd_part1": "void msd_partition_text_kernel",
    "msd_hist": "msd_hist_elems_kernel",
    "pack_text": "void pack_kernel",
    "round_keys": "make_keys_round_kernel",
    "msd_hist1": "void msd_hist_text_kernel",
    "round_rank": "rank_kernel",
    "occ_build": "occ_dna_count_kernel",
}


def main(path, n, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        a = launches.setdefault(int(d["ID"]), {"kernel": d["Kernel Name"]})
        a[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
    ids = sorted(launches)
    # the second build starts at the second pack kernel
    packs = [i for i in ids if launches[i]["kernel"].startswith("void pack_kernel")]
    first = packs[-1]
    total = 0.0
    res = {"_source": f"{path}: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
                      f"--clock-control none, python tools/msd_probe.py {n} 4 msd (second, warm build); bytes read + "
                      "written per launch"}
    for i in ids:
        if i < first:
            continue
        a = launches[i]
        by = a.get("dram__bytes_read.sum", 0.0) + a.get("dram__bytes_write.sum", 0.0)
        total += by
        for stage, prefix in STAGES.items():
            if a["kernel"].startswith(prefix) and stage not in res:
                res[stage] = {"dram_bytes_per_launch": by, "n": n, "kernel": a["kernel"][:60],
                              "ms": a.get("gpu__time_duration.sum", 0.0) / 1e6}
    res["_total_dram_bytes_one_build"] = total
    old = {}
    try:
        old = json.load(open(out))
    except Exception:
        pass
    for k, v in old.items():  # keep entries this capture does not cover (e.g. the search kernel)
        if k not in res and not k.startswith("_"):
            res[k] = v
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(float(sys.argv[2])), sys.argv[3] if len(sys.argv) > 3 else "profiles/roofline_traffic.json")
