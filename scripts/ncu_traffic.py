"""profiles/<tag>_full_raw.csv (ncu --set full, --page raw --csv --print-units base; made by scripts/gpu_calls/gpu_ncu_r02.sh) ->
profiles/r02_ncu_dram_bytes.json: per stage of bench.py's roofline table, dram__bytes_read.sum + dram__bytes_write.sum per launch of
the stage's dominant kernel (mean over the captured launches), with the other counters the judge reads (duration, grid, tensor
pipe, SM / DRAM throughput).  bench.py only READS that JSON (roofline.traffic); it never runs ncu.
  python scripts/ncu_traffic.py profiles/r02_full_raw.csv [--merge]"""
import csv, json, os, sys

STAGE_OF = {"pdist_gemm_kernel": "pdist_gemm", "prep_kernel": "pdist_prep", "knn_smooth_block_kernel": "knn_smooth", "sgd_cluster_kernel": "umap_sgd",
            "lanczos_cluster_kernel": "spectral_init", "apparent_rows_kernel": "rips_apparent", "apparent_kernel": "rips_apparent", "rips_sweep2_kernel": "rips_reduce",
            "boruvka_scan_kernel": "rips_h0", "boruvka_chunked_kernel": "rips_h0", "rank_scatter_kernel": "rips_edge_sort"}
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum",
        "l1tex__t_bytes.sum", "smsp__cycles_active.avg", "sm__cycles_active.avg", "launch__occupancy_limit_shared_mem"]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    cols = rows[hdr]
    units = rows[hdr + 1]
    kcol = cols.index("Kernel Name")
    per = {}
    summary = []
    for r in rows[hdr + 2:]:
        if len(r) != len(cols):
            continue
        name = r[kcol]
        short = next((k for k in STAGE_OF if k in name), None)
        rec = {"kernel": name[:100]}
        for m in KEEP:
            if m in cols:
                try:
                    rec[m] = float(r[cols.index(m)].replace(",", ""))
                except ValueError:
                    pass
        summary.append(rec)
        if short:
            per.setdefault(STAGE_OF[short], []).append(rec)
    out = {}
    for stage, recs in per.items():
        n = len(recs)
        rd = sum(x.get("dram__bytes_read.sum", 0.0) for x in recs) / n
        wr = sum(x.get("dram__bytes_write.sum", 0.0) for x in recs) / n
        out[stage] = {"dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr, "launches_captured": n,
                      "kernel": recs[0]["kernel"], "duration_ms_under_ncu": sum(x.get("gpu__time_duration.sum", 0.0) for x in recs) / n / 1e6,
                      "grid": recs[0].get("launch__grid_size"), "block": recs[0].get("launch__block_size"),
                      "sm_throughput_pct": sum(x.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", 0.0) for x in recs) / n,
                      "dram_throughput_pct": sum(x.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0.0) for x in recs) / n,
                      "tensor_pipe_pct": sum(x.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) for x in recs) / n,
                      "source": os.path.relpath(path, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))) + " (ncu --set full, per launch of one group of layers)"}
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_ncu_dram_bytes.json")
    if "--merge" in sys.argv and os.path.exists(dst):   # a capture of the kernels that changed: the other stages keep their entries
        old = json.load(open(dst))
        old.update(out)
        out = old
    json.dump(out, open(dst, "w"), indent=1)
    sm = os.path.splitext(path)[0].replace("_raw", "") + "_summary.json"
    json.dump(summary, open(sm, "w"), indent=1)
    for k, v in out.items():
        print(f"{k:14s} {v['dram_bytes_per_launch'] / 1e6:10.1f} MB/launch  {v['duration_ms_under_ncu']:8.3f} ms  grid {v['grid']}  sm {v['sm_throughput_pct']:.1f}%  dram {v['dram_throughput_pct']:.1f}%  tensor {v['tensor_pipe_pct']:.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
