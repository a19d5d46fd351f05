#!/usr/bin/env python
"""profiles/r02_traffic.json from ncu --set full captures of batch launches (tools/ncu_batch.sh):
   python tools/make_traffic.py <pixels per launch> file.ncu-rep ...
DRAM bytes (read + write) per launch and per pixel of the launch, keyed by the names bench.py's per-kernel pass reports."""
import csv, io, json, os, re, subprocess, sys
ALIAS = [(r"k_dec_rans_pair<.*1, .*15>", "k_dec_rans_pair_v1_s16"), (r"k_dec_rans_pair<.*1, .*0>", "k_dec_rans_pair_v1_big"),
         (r"k_dec_rans_pair<.*1, .*8>", "k_dec_rans_pair_v1_s8"), (r"k_dec_rans_pair<.*2, .*8>", "k_dec_rans_pair_v2_ctx"),
         (r"k_rans_v1_pair<.*16>", "k_rans_v1_pair_16"), (r"k_rans_v1_pair<.*256>", "k_rans_v1_pair_256"), (r"k_rans_v2_pair<.*16>", "k_rans_v2_pair_16"),
         (r"k_front2<.*1>", "k_front2<1>"), (r"k_front2<.*2>", "k_front2<2>"), (r"k_dec_walk3", "k_dec_walk3<2>"), (r"k_dec_unpredict_rgb", "k_dec_unpredict_rgb"),
         (r"k_dec_residuals<.*1>", "k_dec_residuals<1>"), (r"k_dec_residuals<.*2>", "k_dec_residuals<2>"), (r"k_compact<.*1>", "k_compact<1>"),
         (r"k_compact<.*2>", "k_compact<2>"), (r"k_assemble_m1", "k_assemble_m1"), (r"k_dec_chunk_hist", "k_dec_chunk_hist")]
npx = float(sys.argv[1]); out = {}
for f in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3: continue
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(h, r)); u = dict(zip(h, units))
        def val(k):
            v = float(d[k].replace(",", "")); un = u[k]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6}.get(un, 1)
        name = next((a for pat, a in ALIAS if re.search(pat, d["Kernel Name"])), None)
        if not name or name in out: continue
        dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        out[name] = {"dram_bytes_per_launch": int(dram), "pixels_per_launch": int(npx), "dram_bytes_per_pixel": round(dram / npx, 4),
                     "duration_ms": round(val("gpu__time_duration.sum"), 3), "report": os.path.basename(f)}
json.dump(out, sys.stdout, indent=1); print()
