"""Summarises gpurun_out/<tag>_*.ncu-rep (ncu --set full captures made by tools/ncu_kernels.sh) into
profiles/<tag>_ncu_kernels.json: per kernel the metrics the roofline discussion uses.  Runs in the build container
(ncu reads reports without a GPU).

    python tools/ncu_summary.py r02
"""
import csv
import glob
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "gpc__cycles_elapsed.avg.per_second"]
SHAPES = {"attn_S4680": "Lq 4680, S 4680, H 12", "attn_S18720": "Lq 4680, S 18720, H 12", "attn_S32760": "Lq 4680, S 32760, H 12",
          "attn_cross": "Lq 4680, S 512, H 12", "gemm_qkv": "4680 x 4608 x 1536 bias", "gemm_o_proj": "4680 x 1536 x 1536 gate+residual",
          "gemm_ffn1": "4680 x 8960 x 1536 GELU", "gemm_ffn2": "4680 x 1536 x 8960 gate+residual",
          "gemm_o_proj_stats": "4680 x 1536 x 1536 gate+residual + row statistics of the output",
          "gemm_cross_q_fold": "4680 x 1536 x 1536, LayerNorm folded into the epilogue + row statistics of the output",
          "qk_rope_stream": "4680 x 1536 q,k -> q, cache (statistics from the QKV epilogue)",
          "ln_modulate": "4680 x 1536", "ln_affine": "4680 x 1536", "rmsnorm": "4680 x 1536", "qk_norm_rope": "4680 x 1536 q,k -> q, cache"}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    out = {"_how": "ncu --set full --clock-control none --import-source on, one launch per kernel after 3 warm-up launches "
                   "(tools/ncu_kernels.sh -> tools/gpu_microbench.py, L2 flushed before the captured launch)", "kernels": {}}
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"{tag}_*.ncu-rep"))):
        name = os.path.basename(rep)[len(tag) + 1:-len(".ncu-rep")]
        r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(io.StringIO(r.stdout)))
        if len(rows) < 3:
            print("skip", rep, r.stderr[-200:])
            continue
        header, units, vals = rows[0], rows[1], rows[2]
        rec = {"capture": name, "shape": SHAPES.get(name, "")}
        kname = vals[header.index("Kernel Name")] if "Kernel Name" in header else name
        for k in KEEP:
            if k in header:
                i = header.index(k)
                rec[k] = f"{vals[i]} {units[i]}".strip()
        out["kernels"][f"{kname.split('(')[0].replace('void ', '').replace('sfb::', '')} [{name}]"] = rec
    path = os.path.join(ROOT, "profiles", f"{tag}_ncu_kernels.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, len(out["kernels"]), "kernels")


if __name__ == "__main__":
    main()
