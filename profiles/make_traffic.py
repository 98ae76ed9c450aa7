"""Write profiles/traffic.json from an `ncu --set full` capture (run here, no GPU needed):
    python profiles/make_traffic.py <rep> <refs> <length>
Per kernel: dram__bytes_read.sum + dram__bytes_write.sum of its first profiled launch, in bytes.
bench.py reports it as roofline.traffic when it runs the same workload."""
import csv, json, os, subprocess, sys

rep, refs, length = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
kn, ir, iw = hdr.index('Kernel Name'), hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
out = {}
for vals in rows[2:]:
    name = vals[kn].split('(')[0].split('::')[-1].split('<')[0]
    if name in out:
        continue
    out[name] = float(vals[ir].replace(',', '')) * scale[units[ir]] + float(vals[iw].replace(',', '')) * scale[units[iw]]
json.dump({"refs": refs, "length": length, "source": os.path.basename(rep), "kernels": out},
          open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json"), "w"), indent=1)
print(out)
