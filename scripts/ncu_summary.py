#!/usr/bin/env python
"""Summarise an .ncu-rep (captured with `ncu --set full`) into profiles/<tag>_ncu_summary.txt and
update profiles/traffic.json (DRAM bytes per launch of each op, read by bench.py's roofline).

    python scripts/ncu_summary.py gpurun_out/prof_r01a.ncu-rep r01a [outdir]
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dsmem"),
    ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("smsp__inst_executed_op_shared_ld.sum", "lds"),
    ("smsp__inst_executed_op_global_red.sum", "red_inst"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe%"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
]
# kernel-name fragment -> op of the warp microbench; an op launched as several kernels (resample2d_bwd =
# scatter + flow gradient) gets the SUM of their DRAM bytes (one launch of each)
KERNEL_OPS = [
    ("scatter_rows_kernel<Resample2dScatterGeo", "resample2d_bwd"), ("scatter_tiled_kernel<Resample2dScatterGeo", "resample2d_bwd"),
    ("gather_quad_kernel<RsQuadPolicy", "resample2d_bwd"), ("resample2d_gflow", "resample2d_bwd"), ("resample2d_bwd", "resample2d_bwd"),
    ("resample2d_fwd", "resample2d_fwd"),
    ("scatter_rows_kernel<GridWarpScatterGeo", "grid_warp_bwd"), ("scatter_tiled_kernel<GridWarpScatterGeo", "grid_warp_bwd"),
    ("gather_quad_kernel<GwQuadPolicy", "grid_warp_bwd"), ("grid_warp_tiled_kernel<1>", "grid_warp_bwd"), ("grid_warp_roll_kernel<1>", "grid_warp_bwd"), ("grid_warp_bwd", "grid_warp_bwd"),
    ("grid_warp_fwd", "grid_warp_fwd"), ("grid_warp_tiled_kernel<0>", "grid_warp_fwd"), ("grid_warp_roll_kernel<0>", "grid_warp_fwd"),
    ("block_extractor_fwd", "block_extractor_fwd"), ("block_extractor_bwd", "block_extractor_bwd"),
    ("lar_tiled_kernel<float, 3, 1>", "local_attn_reshape_fwd"), ("lar_tiled_kernel<float, 3, 0>", "local_attn_reshape_bwd"),
]


def to_bytes(v, unit):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    outdir = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = ["# ncu --set full --clock-control none summary of %s (per launch; cold cache, serialised)" % os.path.basename(rep),
           "# columns: " + " ".join(n for _, n in METRICS)]
    traffic_path = os.path.join(outdir, "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    per_kernel = {}
    for r in body:
        name = r[col["Kernel Name"]]
        short = name.split("(")[0].replace("void ", "").replace("ffwm::", "")
        vals = []
        for m, n in METRICS:
            if m in col:
                vals.append("%s=%s%s" % (n, r[col[m]], units[col[m]] if units[col[m]] not in ("%", "") else ""))
        out.append("%-48s grid=%s block=%s  %s" % (short[:48], r[col["Grid Size"]], r[col["Block Size"]], "  ".join(vals)))
        for frag, key in KERNEL_OPS:
            if frag in short.replace("(int)", ""):
                rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
                wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
                per_kernel.setdefault(key, {})[short[:48]] = rd + wr      # last launch of each kernel wins
                break
    for key, d in per_kernel.items():
        traffic[key] = sum(d.values())
    traffic["_source"] = "profiles/%s_ncu_summary.txt" % tag
    open(os.path.join(outdir, "%s_ncu_summary.txt" % tag), "w").write("\n".join(out) + "\n")
    json.dump(traffic, open(traffic_path, "w"), indent=1)
    print("\n".join(out))


if __name__ == "__main__":
    main()
