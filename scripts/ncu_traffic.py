"""profiles/kernel_traffic.json from an `ncu --set full` report of one bench step; usage: ncu_traffic.py rep out.json "capture text"

bench.py reads the dominant kernel's dram bytes per launch from that file (roofline.traffic)."""
import csv, io, json, subprocess, sys

NAMES = {"orbits_kernel": "orbits_kernel", "sparse_decode_kernel": "sparse_decode_kernel", "stream_kernel<512, 1>": "stream_kernel<512,or>", "stream_kernel<512, 2>": "stream_kernel<512,dnf>", "stream_kernel<512, 0>": "stream_kernel<512>",
         "stream_kernel<256, 0>": "stream_kernel<256>", "and_kernel<0>": "and_kernel", "and_kernel<1>": "and_kernel<hits>",
         "eval_kernel<1>": "eval_kernel<hits>", "hot_decode_kernel": "hot_decode_kernel", "merge_kernel": "merge_kernel"}


def main():
    rep, dst, capture = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]

    def val(r, k, scale=None):
        if k not in h:
            return None
        i = h.index(k)
        try:
            v = float(r[i])
        except ValueError:
            return None
        u = units[i]
        mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u)
        return v * mult if mult else v

    res = {}
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        key = next((v for k, v in NAMES.items() if k in name), None)
        if not key:
            continue
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        res[key] = {"dram_bytes_per_launch": (rd + wr) if rd is not None and wr is not None else None, "dram_read": rd, "dram_write": wr,
                    "gpu_time_ms": val(r, "gpu__time_duration.sum"), "warp_instructions": val(r, "smsp__inst_executed.sum"),
                    "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                    "registers": val(r, "launch__registers_per_thread"), "l2_hit_pct": val(r, "lts__t_sector_hit_rate.pct"),
                    "capture": capture}
    import math
    for v in res.values():
        for a, b in list(v.items()):
            if isinstance(b, float) and math.isnan(b):
                v[a] = None      # ncu could not collect this counter for the launch
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    res["kernels_sha"] = bench.kernels_sha()     # bench.py voids these figures as soon as a kernel source changes
    json.dump(res, open(dst, "w"), indent=1)
    for k, v in res.items():
        print(k, v)



if __name__ == "__main__":
    main()
