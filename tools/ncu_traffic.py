"""Turn an `ncu --set full` capture of a sweep kernel into profiles/ncu_traffic.json, the file bench.py reads for
`roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum per launch).

  python tools/ncu_traffic.py gpurun_out/prof_h.ncu-rep f16 [kernel-regex]

The path key is 'f16' (wtx_h_kernel), 'tcgen05' (wtx_tc_kernel) or 'simt' (wtx_partial_kernel).  The capture must be
of the default bench workload on one GPU (20 000 genes x 100 000 cells), one launch = one sweep over one resident
copy of X.  Runs on the build host (ncu -i reads the report without a GPU)."""
import csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, key = sys.argv[1], sys.argv[2]
    pattern = re.compile(sys.argv[3] if len(sys.argv) > 3 else {"f16": "wtx_h_kernel", "tcgen05": "wtx_tc_kernel",
                                                               "simt": "wtx_partial_kernel"}[key])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {name: i for i, name in enumerate(head)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    launches = []
    for r in body:
        if not pattern.search(r[col["Kernel Name"]]):
            continue
        tot = {}
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot[m] = float(r[col[m]].replace(",", "")) * scale[units[col[m]]]
        extra = {}
        for m in ("gpu__time_duration.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
                  "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size"):
            if m in col:
                extra[m] = r[col[m]] + " " + units[col[m]]
        launches.append({"kernel": r[col["Kernel Name"]], "read": tot["dram__bytes_read.sum"],
                         "write": tot["dram__bytes_write.sum"], **extra})
    if not launches:
        raise SystemExit("no launch matching %s in %s" % (pattern.pattern, rep))
    out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        data = json.load(open(out_path))
    except Exception:
        data = {}
    mean = sum(l["read"] + l["write"] for l in launches) / len(launches)
    data[key] = {"bytes": mean, "launches": launches, "source": os.path.basename(rep),
                 "what": "mean over the captured launches of dram__bytes_read.sum + dram__bytes_write.sum"}
    json.dump(data, open(out_path, "w"), indent=1)
    print(json.dumps(data[key], indent=1))


if __name__ == "__main__":
    main()
