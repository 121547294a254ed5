"""Top stalled instructions of an ncu report with source (python tools/ncu_stalls.py rep [N])."""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
for part in txt.split('"Kernel Name",')[1:]:
    lines = part.split("\n")
    print("=====", lines[0][:110])
    rows = list(csv.reader(lines[1:]))
    hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[1:] if len(r) == len(hdr)]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
    agg = {}
    for r in data:
        for k in hdr:
            if k.startswith("stall_") and "Not Issued" not in k:
                agg[k] = agg.get(k, 0) + int(r[ix[k]] or 0)
    print("total samples", tot, {k[6:]: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]})
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:N]:
        st = {k[6:]: int(r[ix[k]] or 0) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
        st = {k: v for k, v in st.items() if v > 0.1 * int(r[ix["# Samples"]])}
        print(r[ix["Address"]][-5:], r[ix["# Samples"]], r[ix["Source"]][:64], st)
