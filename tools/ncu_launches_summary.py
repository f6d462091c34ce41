"""Summarise an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv) of
tools/run_forward.py --n 2 into per-kernel-family totals of the SECOND forward (warm library state):
    python tools/ncu_launches_summary.py gpurun_out/launches.csv profiles/r1_launches_vNN_summary.json
bench.py reads the newest round's profiles/r<N>_*_summary.json for roofline.traffic (DRAM bytes per launch of the dominant kernel)."""
import collections
import csv
import json
import sys

FAMILY = [("decoder_mega", "decoder_mega"), ("gemm_fused2", "gemm_fused2_tcgen05"), ("gemm2_bf16x3", "gemm2_bf16x3_pair"), ("Cfg<128, 2, 3", "gemm_bf16x3_wide"), ("Cfg<128, 3, 1", "gemm_bf16x3_deep"),
          ("Cfg<64, 3, 2", "gemm_bf16x3_n64"), ("Cfg<256, 2, 1", "gemm_bf16x3_big"), ("stem_tc", "stem_conv"), ("dwconv", "dwconv3x3x3"), ("attn_tc", "attention_tc"),
          ("attn", "attention"),
          ("layernorm", "layernorm"), ("gather_rows", "gather_rows"), ("sgemm", "sgemm_fp32"), ("head_gemm", "sgemm_fp32"), ("pool_mix", "pool_mix"), ("maxpool", "maxpool"),
          ("tpool", "tpool"), ("posenc", "posenc"), ("mask_resize", "mask_resize"), ("to_split", "to_split")]


def main(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    by_id = collections.OrderedDict()
    for x in csv.DictReader(lines):
        d = by_id.setdefault(x["ID"], {"name": x["Kernel Name"]})
        v, u = float(x["Metric Value"].replace(",", "")), x["Metric Unit"]
        if x["Metric Name"] == "gpu__time_duration.sum":
            d["us"] = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[u] * v
        else:
            d[x["Metric Name"]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    ids = list(by_id.values())
    half = ids[len(ids) // 2:]
    fam = collections.OrderedDict()
    for d in half:
        name = next((f for k, f in FAMILY if k in d["name"]), d["name"].split("(")[0])
        a = fam.setdefault(name, {"launches": 0, "us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        a["launches"] += 1
        a["us"] += d["us"]
        a["dram_read_bytes"] += d.get("dram__bytes_read.sum", 0.0)
        a["dram_write_bytes"] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a["us"] for a in fam.values())
    for a in fam.values():
        a["share"] = round(a["us"] / tot, 4)
        a["dram_bytes_per_launch"] = (a["dram_read_bytes"] + a["dram_write_bytes"]) / a["launches"]
        a["us"] = round(a["us"], 1)
    out = {"source": src, "forward_launches": len(half), "sum_us": round(tot, 1), "kernels": fam,
           "note": "ncu serialises launches and measures each cold: shares are comparable with bench.py's, absolute times are not"}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps({k: (v["launches"], v["us"], v["share"]) for k, v in fam.items()}))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
