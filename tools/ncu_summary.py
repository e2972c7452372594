"""Text summary of an .ncu-rep (one block per kernel launch): duration, DRAM bytes and % of peak, tensor-pipe %, issue
activity, shared-memory wavefronts / bank conflicts, registers, the top stall lines of the SASS view.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [max_launches] > profiles/x.txt"""
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration [us]"),
    ("dram__bytes_read.sum", "DRAM read [MB]"),
    ("dram__bytes_write.sum", "DRAM write [MB]"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput [% of peak]"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active [%]"),
    ("smsp__issue_active.avg.pct", "issue slots busy [%]"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active [% of max]"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__cycles_elapsed.max", "SM cycles"),
]


def main(path, max_launches):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    seen = {}
    for r in rows[2:]:
        key = r[name_i]
        seen[key] = seen.get(key, 0) + 1
        if seen[key] > 1 or len(seen) > max_launches:
            continue
        print("=" * 110)
        print(key[:200])
        for m, label in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {label:34s} {r[i]:>16s} {units[i]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    kern, cur = [], None
    for r in csv.reader(src.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            kern.append(cur)
        elif cur is not None:
            if cur["hdr"] is None:
                cur["hdr"] = r
            else:
                cur["rows"].append(r)
    done = set()
    for k in kern:
        if k["name"] in done or not k["rows"] or "Warp Stall Sampling (All Samples)" not in k["hdr"]:
            continue
        done.add(k["name"])
        h = k["hdr"]
        ss, s_i = h.index("Warp Stall Sampling (All Samples)"), h.index("Source")
        tot = sum(int(r[ss]) for r in k["rows"]) or 1
        print("-" * 110)
        print("top stall samples:", k["name"][:150])
        for r in sorted(k["rows"], key=lambda r: -int(r[ss]))[:8]:
            print(f"   {100 * int(r[ss]) / tot:5.1f}%  {r[s_i][:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
