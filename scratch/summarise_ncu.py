"""Summarise ncu CSV captures (gpurun_out/) into the tracked files under profiles/.
  python scratch/summarise_ncu.py launches <launches.csv> <out.json> [note]
  python scratch/summarise_ncu.py raw <raw.csv> <out_summary.csv> [traffic.json]
"""
import collections, csv, json, sys

def read_rows(path):
    rows = list(csv.reader(open(path, newline="")))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            return r, rows[i + 1:]
    raise SystemExit("no header in " + path)

def short(name):
    for key in ("gemm_grouped_kernel", "flow_tc_kernel", "flow_tc_pack", "geom_lossgrad_kernel", "geom_forward_kernel",
                "geom_backward_angles", "elev_stats", "adam_kernel", "colsum_batched_zero", "colsum_batched", "cast_weight_batched",
                "pack_rows", "grad_compress", "pmpjpe_kernel", "mpjpe_kernel", "eval_lift_score", "occ_"):
        if key in name:
            if key in ("flow_tc_kernel", "geom_lossgrad_kernel") and "<" in name:
                return key + name[name.index("<"):name.index(">") + 1]
            return key
    return name.split("(")[0][:60]

if sys.argv[1] == "launches":
    H, rows = read_rows(sys.argv[2])
    ki, vi, mi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Name")
    tot = collections.OrderedDict()
    for r in rows:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        k = short(r[ki])
        t = tot.setdefault(k, [0, 0.0])
        t[0] += 1
        t[1] += float(r[vi].replace(",", "")) / 1e3
    total = sum(v[1] for v in tot.values())
    out = {"note": sys.argv[4] if len(sys.argv) > 4 else "", "total_us": total,
           "kernels": [{"kernel": k, "launches": v[0], "us": round(v[1], 1), "share": round(v[1] / total, 4)}
                       for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])]}
    json.dump(out, open(sys.argv[3], "w"), indent=1)
    print(json.dumps(out["kernels"][:8]))
else:
    H, rows = read_rows(sys.argv[2])
    want = ["launch__grid_size", "launch__block_size", "gpu__time_duration.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active",
            "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "local_load_bytes" if False else "smsp__inst_executed_op_local_ld.sum",
            "smsp__inst_executed_op_local_st.sum"]
    if "Metric Name" in H:                       # long format
        ki, vi, mi, ii = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Name"), H.index("ID")
        d = collections.OrderedDict()
        for r in rows:
            if len(r) > vi:
                d.setdefault((r[ii], short(r[ki])), {})[r[mi]] = r[vi]
        recs = [(k[1], m) for k, m in d.items()]
    else:                                        # --page raw: wide format, second row = units
        ki = H.index("Kernel Name")
        recs = [(short(r[ki]), dict(zip(H, r))) for r in rows[1:] if len(r) == len(H)]
    cols = [c for c in want if any(c in m for _, m in recs)]
    w = csv.writer(open(sys.argv[3], "w", newline=""))
    w.writerow(["Kernel Name"] + cols)
    dram = []
    for k, m in recs:
        w.writerow([k] + [m.get(c, "") for c in cols])
        try:
            dram.append(float(m["dram__bytes_read.sum"].replace(",", "")) + float(m["dram__bytes_write.sum"].replace(",", "")))
        except Exception:
            pass
    print(len(recs), "launches ->", sys.argv[3])
    if len(sys.argv) > 4 and dram:
        json.dump({"source": "ncu --set full, %d consecutive gemm_grouped_kernel launches of bench.py --no-graph --serial (%s)" %
                   (len(dram), sys.argv[3]), "dram_bytes_per_launch_mean": sum(dram) / len(dram), "launches": len(dram)},
                  open(sys.argv[4], "w"), indent=1)
