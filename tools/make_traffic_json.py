"""Write profiles/megakernel_traffic.json from an `ncu --set full` capture of tools/mega_ncu.py (raw page CSV), stamped with the
sha256 of the kernel source it was captured from: bench.py reports roofline.traffic from this file and null when the kernel has
changed since.  usage: python tools/make_traffic_json.py RAW.csv TOKENS_PER_LAUNCH SOURCE_NAME"""
import csv, hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
hdr, vals = rows[0], rows[-1]
units = rows[1] if len(rows) > 2 else [""] * len(hdr)
get = lambda name: next((float(v.replace(",", "")), u) for h, v, u in zip(hdr, vals, units) if h == name)
rd, ru = get("dram__bytes_read.sum")
wr, wu = get("dram__bytes_write.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "": 1.0}
total = rd * scale.get(ru, 1.0) + wr * scale.get(wu, 1.0)
tokens = int(sys.argv[2])
sha = hashlib.sha256(open(os.path.join(ROOT, "gabby_b200", "csrc", "mega_decode.cuh"), "rb").read()).hexdigest()
commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
out = {"dram_bytes_per_token": total / tokens, "dram_bytes_per_launch": total, "tokens_per_launch": tokens, "source": sys.argv[3],
       "mega_decode_cuh_sha256": sha, "commit": commit, "read_bytes": rd * scale.get(ru, 1.0), "write_bytes": wr * scale.get(wu, 1.0)}
json.dump(out, open(os.path.join(ROOT, "profiles", "megakernel_traffic.json"), "w"), indent=1)
print(out)
