"""Summarise an `ncu --set full` report (read here, without a GPU: `ncu -i <rep> --page raw|source --csv`) into a JSON / markdown row
per captured launch: duration, tensor-pipe activity, DRAM and L2 traffic, occupancy limits and the top warp-stall reasons.

    python scripts/ncu_summary.py gpurun_out/prof_r02_gemm.ncu-rep [--md]
"""
import csv
import io
import json
import subprocess
import sys


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    return rows[hdr], rows[hdr + 1], rows[hdr + 2:]


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


WANT = {
    "gpu__time_duration.sum": "duration",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "lts__t_bytes.sum": "l2_bytes",
    "lts__t_sectors_srcunit_tex_op_read.sum": "l2_to_sm_sectors",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic",
    "launch__shared_mem_per_block_static": "smem_static",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "smsp__cycles_active.avg": "smsp_cycles_active",
    "sm__cycles_elapsed.max": "cycles_elapsed",
}


def stalls(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    res, cur, hdr = [], None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {}
            res.append(cur)
            hdr = None
        elif r and r[0] == "Address":
            hdr = r
        elif hdr and cur is not None and len(r) == len(hdr):
            for h, v in zip(hdr, r):
                if h.startswith("stall_") and "Not Issued" not in h:
                    try:
                        cur[h] = cur.get(h, 0) + int(v or 0)
                    except ValueError:
                        pass
    top = []
    for d in res:
        tot = sum(d.values()) or 1
        top.append([(k, round(100.0 * v / tot, 1)) for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:4]])
    return top


def main():
    rep = sys.argv[1]
    hdr, units, rows = raw_rows(rep)
    ci = {h: i for i, h in enumerate(hdr)}
    st = stalls(rep)
    out = []
    for k, r in enumerate(rows):
        d = {"kernel": r[ci["Kernel Name"]].split("(")[0][-60:], "grid": r[ci.get("Grid Size", 0)] if "Grid Size" in ci else None}
        for m, name in WANT.items():
            if m in ci:
                d[name] = num(r[ci[m]])
                d[name + "_unit"] = units[ci[m]]
        if k < len(st):
            d["top_stalls_pct"] = st[k]
        out.append(d)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
