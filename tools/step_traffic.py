"""Condense an ncu launch list of ONE train step into the files kept under profiles/:
    python tools/step_traffic.py gpurun_out/r2_launches_raw.csv profiles/r2_launches_step.csv profiles/r2_step_traffic.json
input : ncu --csv --log-file output of
        B200FM_NCU_ONE_STEP=1 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
            --clock-control none --csv --log-file <raw.csv> python bench.py --steps 1 --warmup 3 --no-cpu-baseline
output: one row per launch (kernel, grid, block, ns, DRAM bytes read / written) and the per-family summary bench.py reads for
        `roofline.traffic` (DRAM bytes per launch of the tcgen05 GEMM family)."""
import csv, json, re, sys

raw, out_csv, out_json = sys.argv[1:4]
rows = [r for r in csv.reader(open(raw, errors="replace")) if r]
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
h = rows[hdr_i]
iid, ik, ig, ib, im, iv = (h.index(c) for c in ("ID", "Kernel Name", "Grid Size", "Block Size", "Metric Name", "Metric Value"))
launches = {}
for r in rows[hdr_i + 1:]:
    if len(r) <= iv or not r[iid].isdigit():
        continue
    e = launches.setdefault(int(r[iid]), dict(kernel=re.sub(r"\(.*", "", r[ik]).replace("void ", "").strip(), grid=r[ig], block=r[ib]))
    e[r[im]] = float(r[iv].replace(",", ""))
fam = lambda k: ("gemm" if "gemm_kernel" in k else "attention" if "attention" in k else "layernorm" if "layernorm" in k else
                 "adamw" if "adamw" in k else "allreduce" if "allreduce" in k else "other b200fm" if "b200fm" in k else "torch / other")
with open(out_csv, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["id", "kernel", "grid", "block", "gpu__time_duration.sum [ns]", "dram__bytes_read.sum [byte]", "dram__bytes_write.sum [byte]"])
    for i in sorted(launches):
        e = launches[i]
        w.writerow([i, e["kernel"], e["grid"], e["block"], int(e.get("gpu__time_duration.sum", 0)), int(e.get("dram__bytes_read.sum", 0)), int(e.get("dram__bytes_write.sum", 0))])
tot_ns = sum(e.get("gpu__time_duration.sum", 0) for e in launches.values())
summary = {"source": f"{out_csv} (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, B200FM_NCU_ONE_STEP=1 "
                     "python bench.py --steps 1 --warmup 3 --no-cpu-baseline; one steady-state 4M-B mod7 train step issued eagerly, 1x B200)",
           "launches_per_step": len(launches), "kernel_ms_per_step": tot_ns / 1e6}
for name in ("gemm", "attention", "layernorm", "adamw", "other b200fm", "torch / other"):
    es = [e for e in launches.values() if fam(e["kernel"]) == name]
    if not es:
        continue
    ns = sum(e.get("gpu__time_duration.sum", 0) for e in es)
    rd, wr = sum(e.get("dram__bytes_read.sum", 0) for e in es), sum(e.get("dram__bytes_write.sum", 0) for e in es)
    summary[name.replace(" / ", "_").replace(" ", "_")] = dict(launches=len(es), ms=ns / 1e6, share_of_kernel_time=ns / tot_ns, dram_bytes_read=rd, dram_bytes_write=wr,
                                                              dram_bytes_per_launch=(rd + wr) / len(es))
json.dump(summary, open(out_json, "w"), indent=1)
print(json.dumps(summary, indent=1))
