"""profiles/roofline_traffic.json from ncu --set full reports: DRAM bytes (read + write) per launch of each captured kernel,
stamped with the hash of the kernel sources (bench.py refuses the numbers when the sources have changed since).
    python tools/make_traffic.py rep1.ncu-rep rep2.ncu-rep ...   > profiles/roofline_traffic.json"""
import csv, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
out = {}
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    scale = lambda u: {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
    acc = {}
    for r in rows[2:]:
        k = r[ik].split("(")[0].split("<")[0].replace("void ", "").replace("pcreg::", "").strip()
        b = float(r[ir].replace(",", "")) * scale(units[ir]) + float(r[iw].replace(",", "")) * scale(units[iw])
        acc.setdefault(k, []).append(b)
    for k, v in acc.items():
        out[k] = sum(v) / len(v)
print(json.dumps(dict(csrc_sha16=bench.kernel_source_sha(), how="ncu --set full --clock-control none: dram__bytes_read.sum + dram__bytes_write.sum, mean per launch of the captured launches",
                      reports=[os.path.basename(r) for r in sys.argv[1:]], kernels=out), indent=1))
