"""ncu csv (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per launch) -> JSON for bench.py's
roofline.traffic: DRAM bytes per launch of the dominant kernel, averaged over the launches of one denoiser call.

    python tools/summarize_dram.py profiles/r2_igemm_dram_unet_b256.csv profiles/igemm_dram_traffic.json \
        profiles/r2_ops_unet_b256.names.txt        # op names of the captured plan (tools/profile_call.py writes them)
"""
import collections
import csv
import io
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0,
        "msecond": 1e3, "ms": 1e3}


def main(src, dst, names_file):
    lines = [l for l in open(src) if not l.startswith("==")]
    per = collections.defaultdict(dict)
    for r in csv.DictReader(io.StringIO("".join(lines))):
        v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
        per[r["ID"]][r["Metric Name"]] = v
        per[r["ID"]]["name"] = r["Kernel Name"]
    n = len(per)
    rd = sum(p.get("dram__bytes_read.sum", 0.0) for p in per.values())
    wr = sum(p.get("dram__bytes_write.sum", 0.0) for p in per.values())
    us = sum(p.get("gpu__time_duration.sum", 0.0) for p in per.values())
    from bench import igemm_fingerprint

    # the hash ties the capture to the kernel sources it was taken from: bench.py quotes `traffic` only while it matches
    out = {"kernel": "igemm_sm100_kernel (all launches of one latent-UNet denoiser call, batch 256, bf16)", "launches": n,
           "source_sha16": igemm_fingerprint(Path(names_file).read_text().split("\n")), "capture": Path(src).name,
           "dram_read_bytes": rd, "dram_write_bytes": wr, "bytes_per_launch": (rd + wr) / max(n, 1),
           "ncu_time_us": us, "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum ({src}); cold cache per launch"}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
